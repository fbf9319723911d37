// Stand-in for GR/plugin.hpp (test infrastructure, like the rest of oracle/ref_stub): just enough of the plugin API —
// GR_PLUGIN -> grPluginInstance(), gr::plugin<>::addBlockType<TBlock>(type, params) (GR/plugin.hpp:40-63, 77-98) — for
// gr4_packet_modem_b200/blocks/b200_plugin.cpp's GR4 branch to be compiled here.
#pragma once
#include <string>
#include <vector>

namespace gr {
template <int ABI_VERSION = 1>
class plugin
{
public:
    std::vector<std::string> provided;
    template <typename TBlock>
    void addBlockType(std::string blockType = {}, std::string /*blockParams*/ = {})
    {
        static_assert(sizeof(TBlock) > 0, "complete block type required");
        provided.push_back(std::move(blockType));
    }
};
} // namespace gr

#define GR_PLUGIN(Name, Author, License, Version)                \
    inline gr::plugin<>& grPluginInstance()                      \
    {                                                            \
        static gr::plugin<> instance;                            \
        return instance;                                         \
    }

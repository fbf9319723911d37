// b200_shell_common.hpp — what the block shells share: the GR4 / compat switch, and the translation
// between a GR4 tag (a property_map) and the plain-C b200sync_stream_tag that crosses the C ABI.
//
// A stream tag carries the syncword_* keys (PM/syncword_detection.hpp:106-114) by value, because the
// hot path reads them (SymbolFilter: amplitude, time_est, phase, freq — PM/symbol_filter.hpp:141-156),
// and everything else as an opaque id: the shell keeps the original property_map under that id and
// re-attaches it when the library hands the tag back (delayed / re-indexed).
#pragma once
#include <complex>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../../include/b200sync.h"

#if __has_include(<gnuradio-4.0/Block.hpp>)
#include <gnuradio-4.0/Block.hpp>
#include <gnuradio-4.0/reflection.hpp>
#define B200SYNC_HAVE_GR4 1
#else
#include "gr4_compat.hpp"
#define B200SYNC_HAVE_GR4 0
#endif

namespace b200sync_shell {

#if B200SYNC_HAVE_GR4
template <typename T, typename V>
inline T pmt_cast(const V& v) { return pmtv::cast<T>(v); }
template <typename T>
inline auto pmt_make(T x) { return pmtv::pmt(x); }
#else
template <typename T, typename V>
inline T pmt_cast(const V& v)
{
    return std::visit(
        [](auto&& a) -> T {
            using A = std::decay_t<decltype(a)>;
            if constexpr (std::is_same_v<A, std::string>) throw gr::exception("pmt: string where a number was expected");
            else return static_cast<T>(a);
        },
        v);
}
template <typename T>
inline gr::pmt_value pmt_make(T x) { return gr::pmt_value(x); }
#endif

inline bool is_syncword_key(const std::string& k) { return k.rfind("syncword_", 0) == 0; }

// id -> original property_map of tags that are in flight inside the library
class TagStore
{
    std::map<uint32_t, gr::property_map> _maps;
    uint32_t _next = 0;

public:
    void clear() { _maps.clear(); _next = 0; }

    // property_map -> stream tag at chunk offset 0 (the runtime cuts chunks at tags, GR/Block.hpp:1501-1506)
    b200sync_stream_tag to_abi(const gr::property_map& m)
    {
        b200sync_stream_tag t{};
        t.index = 0;
        t.has_syncword = m.contains("syncword_amplitude") ? 1u : 0u;  // PM/symbol_filter.hpp:130
        if (t.has_syncword) {
            auto f = [&](const char* k) { return m.contains(k) ? pmt_cast<double>(m.at(k)) : 0.0; };
            t.sw.syncword_amplitude = static_cast<float>(f("syncword_amplitude"));
            t.sw.syncword_phase = static_cast<float>(f("syncword_phase"));
            t.sw.syncword_freq = f("syncword_freq");
            t.sw.syncword_freq_bin = static_cast<int32_t>(f("syncword_freq_bin"));
            t.sw.syncword_noise_power = static_cast<float>(f("syncword_noise_power"));
            t.sw.syncword_esn0_db = static_cast<float>(f("syncword_esn0_db"));
            t.sw.syncword_time_est = static_cast<float>(f("syncword_time_est"));
        }
        if (++_next == 0) ++_next;
        t.other = _next;
        _maps[_next] = m;
        return t;
    }

    // stream tag handed back by the library -> the property_map to publish
    gr::property_map from_abi(const b200sync_stream_tag& t)
    {
        gr::property_map m;
        if (auto it = _maps.find(t.other); it != _maps.end()) {
            m = std::move(it->second);
            _maps.erase(it);
        }
        // the only key the hot path rewrites (PM/symbol_filter.hpp:150-155)
        if (t.has_syncword && m.contains("syncword_phase")) m["syncword_phase"] = pmt_make(t.sw.syncword_phase);
        return m;
    }
};

}  // namespace b200sync_shell

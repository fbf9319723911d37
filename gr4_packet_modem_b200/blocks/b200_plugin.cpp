// b200_plugin.cpp — the B200 shells packaged as a GNU Radio 4.0 plugin (a shared object the runtime's PluginLoader
// picks up from its plugin directories, GR/PluginLoader.hpp), next to — not instead of — the reference's blocks.
//
// Follows the reference's own packaging: GR/plugin.hpp:77-98 (GR_PLUGIN defines grPluginInstance(), gr_plugin_make and
// gr_plugin_free) and the non-template registration form of python/bindings/register_syncword_detection.cpp:5-10
// (`reg.addBlockType<SyncwordDetection>("gr::packet_modem::SyncwordDetection", "")`).  The shells take the reference
// blocks' reflected settings, so a flowgraph description (GRC/YAML) switches block by block by appending "B200" to the
// type name.
//
// Build inside a GR4 tree, where <gnuradio-4.0/Block.hpp> exists (b200_shell_common.hpp then sets B200SYNC_HAVE_GR4):
//   c++ -std=c++23 -shared -fPIC b200_plugin.cpp -I<gr4>/core/include -I<repo>/include -L<repo>/gr4_packet_modem_b200 \
//       -lb200sync -o libb200sync_gr4_plugin.so
// NOT compiled in this repository's container (GNU Radio 4.0's dependencies are fetched from the network at
// configure time): unlike the shells themselves — whose GR4 branch is compiled and run against the
// stand-in runtime of oracle/ref_stub — this file has only been checked against the headers it cites.
#include "b200_shell_common.hpp"

#if B200SYNC_HAVE_GR4
#include <gnuradio-4.0/plugin.hpp>

#include "coarse_frequency_correction_b200.hpp"
#include "costas_loop_b200.hpp"
#include "pfb_arb_resampler_b200.hpp"
#include "symbol_filter_b200.hpp"
#include "syncword_detection_b200.hpp"
#include "syncword_detection_filter_b200.hpp"
#include "syncword_wipeoff_b200.hpp"

GR_PLUGIN("gr4-packet-modem RX synchronisation on B200 (libb200sync)", "b200-packet-sync", "see repository", "r2")

namespace {
using namespace gr::packet_modem;

template <typename TBlock>
bool add(const char* type_name)
{
    grPluginInstance().template addBlockType<TBlock>(type_name, "");
    return true;
}

const bool registered[] = {
    add<SyncwordDetectionB200>("gr::packet_modem::SyncwordDetectionB200"),
    add<SyncwordDetectionFilterB200>("gr::packet_modem::SyncwordDetectionFilterB200"),
    add<SymbolFilterB200>("gr::packet_modem::SymbolFilterB200"),
    add<CoarseFrequencyCorrectionB200>("gr::packet_modem::CoarseFrequencyCorrectionB200"),
    add<SyncwordWipeoffB200>("gr::packet_modem::SyncwordWipeoffB200"),
    add<CostasLoopB200>("gr::packet_modem::CostasLoopB200"),
    add<PfbArbResamplerB200T<float>>("gr::packet_modem::PfbArbResamplerB200"),
    add<PfbArbResamplerB200T<double>>("gr::packet_modem::PfbArbResamplerB200<double>"),
    add<RotatorB200>("gr::packet_modem::RotatorB200"),
    add<RxFrontEndB200>("gr::packet_modem::RxFrontEndB200"),
};
}  // namespace
#endif  // B200SYNC_HAVE_GR4

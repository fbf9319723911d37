// Drives the C++ block shells of the whole RX synchronisation chain the way the GR4 runtime does:
//   PfbArbResamplerB200 -> RotatorB200 -> SyncwordDetectionB200 -> SyncwordDetectionFilterB200
//     -> CoarseFrequencyCorrectionB200 -> SymbolFilterB200 -> SyncwordWipeoffB200 -> CostasLoopB200
// Each stage is offered bounded chunks that start at tags (GR/Block.hpp:1501-1508), its consume()/publish()
// calls are honoured, and the tags it publishes are carried to the next stage with absolute indices.
// A stand-in for the header parser answers every forwarded syncword with a parsed_header message.
// Output: stage counts and tags as text lines, every stage's stream as a raw cf32 file, compared by
// tests/test_gpu_cpp_shell.py with the Python mirrors (bit-exact) and the oracle.
//   usage: test_rx_chain_shells <raw.cf32> <rrc.f32> <fe_taps.f32> <sf_taps.f32> <rate> <phase_incr>
//                               <payload_bytes> <chunk> <out_prefix>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <deque>
#include <fstream>
#include <span>
#include <vector>

#include "../../gr4_packet_modem_b200/blocks/coarse_frequency_correction_b200.hpp"
#include "../../gr4_packet_modem_b200/blocks/costas_loop_b200.hpp"
#include "../../gr4_packet_modem_b200/blocks/syncword_wipeoff_b200.hpp"
#include "../../gr4_packet_modem_b200/blocks/pfb_arb_resampler_b200.hpp"
#include "../../gr4_packet_modem_b200/blocks/symbol_filter_b200.hpp"
#include "../../gr4_packet_modem_b200/blocks/syncword_detection_b200.hpp"
#include "../../gr4_packet_modem_b200/blocks/syncword_detection_filter_b200.hpp"

using c64 = std::complex<float>;
using TagList = std::vector<gr::Tag>;  // absolute indices, sorted

template <typename T>
static std::vector<T> slurp(const char* path)
{
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", path); std::exit(2); }
    const std::streamsize n = f.tellg();
    f.seekg(0);
    std::vector<T> v(static_cast<size_t>(n) / sizeof(T));
    f.read(reinterpret_cast<char*>(v.data()), n);
    return v;
}

static void dump(const std::string& path, const std::vector<c64>& v)
{
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char*>(v.data()), static_cast<std::streamsize>(v.size() * sizeof(c64)));
}

static void print_tags(const char* stage, const TagList& tags)
{
    for (const auto& t : tags) {
        std::printf("tag %s %lld", stage, static_cast<long long>(t.index));
        for (const auto& [k, v] : t.map) {
            std::printf(" %s=", k.c_str());
            std::visit([](auto&& a) {
                using A = std::decay_t<decltype(a)>;
                if constexpr (std::is_same_v<A, std::string>) std::printf("%s", a.c_str());
                else if constexpr (std::is_same_v<A, int>) std::printf("%d", a);
                else if constexpr (std::is_same_v<A, std::uint64_t>) std::printf("%llu", (unsigned long long)a);
                else std::printf("%.17g", static_cast<double>(a));
            }, v);
        }
        std::printf("\n");
    }
}

// One block with one stream input and one stream output, fed like workInternal does.
// call(in_span, out_span) -> Status; blk.out.published_tags are chunk relative.
template <typename Block, typename Call>
static void run_stage(const char* name, Block& blk, const std::vector<c64>& x, const TagList& in_tags, size_t chunk,
                      size_t min_in, std::vector<c64>& y, TagList& out_tags, Call&& call)
{
    std::vector<c64> obuf(chunk + 64);
    size_t pos = 0, ti = 0, stalls = 0;
    y.clear();
    out_tags.clear();
    while (x.size() - pos >= min_in && x.size() > pos) {
        size_t n = std::min(chunk, x.size() - pos);
        // a chunk never extends past the next tagged sample
        size_t tj = ti;
        while (tj < in_tags.size() && static_cast<size_t>(in_tags[tj].index) <= pos) ++tj;
        if (tj < in_tags.size()) n = std::max<size_t>(min_in, std::min(n, static_cast<size_t>(in_tags[tj].index) - pos));
        n = std::min(n, x.size() - pos);
        if constexpr (requires { blk.offer_input_tag(gr::property_map{}); }) {
            blk.clear_input_tag();
            gr::property_map merged;
            for (; ti < in_tags.size() && static_cast<size_t>(in_tags[ti].index) == pos; ++ti)
                for (const auto& kv : in_tags[ti].map) merged.insert_or_assign(kv.first, kv.second);
            if (!merged.empty()) blk.offer_input_tag(merged);
        }
        gr::ConsumableSpanShim<c64> in_span(std::span<const c64>(x.data() + pos, n));
        gr::PublishableSpanShim<c64> out_span(std::span<c64>(obuf.data(), obuf.size()));
        blk.out.published_tags.clear();
        const auto st = call(in_span, out_span);
        const size_t c = in_span.consumed(), p = out_span.published();
        for (const auto& t : blk.out.published_tags)
            out_tags.push_back(gr::Tag{ static_cast<ssize_t>(y.size()) + t.index, t.map });
        y.insert(y.end(), obuf.begin(), obuf.begin() + static_cast<std::ptrdiff_t>(p));
        pos += c;
        if (st != gr::work::Status::OK) break;
        if (c == 0 && p == 0 && ++stalls > 4) break;
        if (c != 0) stalls = 0;
    }
    std::printf("stage %s consumed %zu produced %zu tags %zu\n", name, pos, y.size(), out_tags.size());
}

int main(int argc, char** argv)
{
    if (argc != 10) return 2;
    const auto raw = slurp<c64>(argv[1]);
    const auto rrc = slurp<float>(argv[2]);
    const auto fe_taps = slurp<float>(argv[3]);
    const auto sf_taps = slurp<float>(argv[4]);
    const float rate = static_cast<float>(std::atof(argv[5]));
    const float phase_incr = static_cast<float>(std::atof(argv[6]));
    const uint64_t payload_bytes = static_cast<uint64_t>(std::atoll(argv[7]));
    const size_t chunk = static_cast<size_t>(std::atoll(argv[8]));
    const std::string prefix = argv[9];
    const gr::property_map none;
    TagList no_tags, t_sd, t_sdf, t_cfc, t_sf, t_unused;
    std::vector<c64> y, z, d, f, g, sym;

    // error text of the reference comes through (PM/pfb_arb_resampler.hpp:70-72)
    {
        gr::packet_modem::PfbArbResamplerB200 bad;
        bad.taps = fe_taps;
        bad.filter_size = 0;
        try { bad.settingsChanged(none, none); std::printf("error none\n"); }
        catch (const gr::exception& e) { std::printf("error %s\n", e.what()); }
    }

    gr::packet_modem::PfbArbResamplerB200 resampler;
    resampler.rate = rate;
    resampler.taps = fe_taps;
    resampler.filter_size = 32;
    resampler.settingsChanged(none, none);
    run_stage("resampler", resampler, raw, no_tags, chunk, 1, y, t_unused,
              [&](auto& i, auto& o) { return resampler.processBulk(i, o); });

    gr::packet_modem::RotatorB200 rotator;
    rotator.phase_incr = phase_incr;
    rotator.settingsChanged(none, none);
    rotator.start();
    run_stage("rotator", rotator, y, no_tags, chunk, 1, z, t_unused,
              [&](auto& i, auto& o) { return rotator.processBulk(i, o); });

    // the fused pair must give the rotator's stream exactly
    {
        gr::packet_modem::RxFrontEndB200 fused;
        fused.rate = rate;
        fused.taps = fe_taps;
        fused.phase_incr = phase_incr;
        fused.settingsChanged(none, none);
        std::vector<c64> zf;
        run_stage("fused_front_end", fused, raw, no_tags, chunk, 1, zf, t_unused,
                  [&](auto& i, auto& o) { return fused.processBulk(i, o); });
        std::printf("fused_equals_pair %d\n",
                    zf.size() == z.size() && std::memcmp(zf.data(), z.data(), z.size() * sizeof(c64)) == 0 ? 1 : 0);
    }

    gr::packet_modem::SyncwordDetectionB200 detection;
    detection.rrc_taps = rrc;
    static const uint8_t sw[64] = { 0,0,0,0,0,0,1,1,0,1,0,0,0,1,1,1,0,1,1,1,0,1,1,0,1,1,0,0,0,1,1,1,
                                    0,0,1,0,0,1,1,1,0,0,1,0,1,0,0,0,1,0,0,1,0,1,0,1,1,0,1,1,0,0,0,0 };
    detection.syncword.assign(sw, sw + 64);
    detection.constellation = { { 1.0f, 0.0f }, { -1.0f, 0.0f } };
    detection.min_freq_bin = -4;
    detection.max_freq_bin = 4;
    detection.start();
    run_stage("syncword_detection", detection, z, no_tags, std::max<size_t>(chunk, 2048), 2048, d, t_sd,
              [&](auto& i, auto& o) { return detection.processBulk(i, o); });
    print_tags("syncword_detection", t_sd);

    gr::packet_modem::SyncwordDetectionFilterB200 filter;
    filter.settingsChanged(none, none);
    filter.start();
    std::deque<gr::Message> headers;  // what HeaderParser would have sent back
    const std::vector<gr::Message> no_messages;
    run_stage("syncword_detection_filter", filter, d, t_sd, chunk, 1, f, t_sdf, [&](auto& i, auto& o) {
        std::vector<gr::Message> pending(headers.begin(), headers.end());
        gr::ConsumableSpanShim<gr::Message> hspan{ std::span<const gr::Message>(pending) };
        gr::ConsumableSpanShim<gr::Message> ispan{ std::span<const gr::Message>(no_messages) };
        const size_t before = filter.out.published_tags.size();
        const auto st = filter.processBulk(hspan, ispan, i, o);
        for (size_t k = 0; k < hspan.consumed() && !headers.empty() && !pending.empty(); ++k) headers.pop_front();
        // every syncword that got through starts a packet whose header decodes fine
        for (size_t k = before; k < filter.out.published_tags.size(); ++k)
            if (filter.out.published_tags[k].map.contains("syncword_amplitude"))
                headers.push_back(gr::Message{ gr::property_map{ { "packet_length", payload_bytes } } });
        return st;
    });
    print_tags("syncword_detection_filter", t_sdf);

    // PM/packet_receiver.hpp:94-95: delay = (rrc_taps.size() - 1) / 2 + samples_per_symbol
    gr::packet_modem::CoarseFrequencyCorrectionB200 freq_correction;
    freq_correction.delay = (rrc.size() - 1) / 2 + 4;
    freq_correction.settingsChanged(none, none);
    freq_correction.start();
    run_stage("coarse_frequency_correction", freq_correction, f, t_sdf, chunk, 1, g, t_cfc,
              [&](auto& i, auto& o) { return freq_correction.processBulk(i, o); });
    print_tags("coarse_frequency_correction", t_cfc);

    gr::packet_modem::SymbolFilterB200 symbol_filter;
    symbol_filter.taps = sf_taps;
    symbol_filter.num_arms = 32;
    symbol_filter.samples_per_symbol = 4;
    symbol_filter.delay = rrc.size() - 1;  // PM/packet_receiver.hpp:115
    symbol_filter.settingsChanged(none, none);
    symbol_filter.start();
    run_stage("symbol_filter", symbol_filter, g, t_cfc, chunk, 1, sym, t_sf,
              [&](auto& i, auto& o) { return symbol_filter.processBulk(i, o); });
    print_tags("symbol_filter", t_sf);

    // SymbolFilter with the frequency correction fused into its load stage, fed the filter's output
    // directly, must give the pair's symbols exactly
    {
        gr::packet_modem::SymbolFilterB200 fused;
        fused.taps = sf_taps;
        fused.num_arms = 32;
        fused.samples_per_symbol = 4;
        fused.delay = rrc.size() - 1;
        fused.fused_cfc_delay = static_cast<int>((rrc.size() - 1) / 2 + 4);
        fused.settingsChanged(none, none);
        fused.start();
        std::vector<c64> symf;
        TagList t_f;
        run_stage("fused_cfc_symbol_filter", fused, f, t_sdf, chunk, 1, symf, t_f,
                  [&](auto& i, auto& o) { return fused.processBulk(i, o); });
        std::printf("fused_cfc_equals_pair %d\n",
                    symf.size() == sym.size() && t_f.size() == t_sf.size() &&
                            std::memcmp(symf.data(), sym.data(), sym.size() * sizeof(c64)) == 0 ? 1 : 0);
    }

    // PM/packet_receiver.hpp:117-125, 203-214: SymbolFilter -> SyncwordWipeoff(bipolar syncword) ->
    // [PayloadMetadataInsert: host control logic, not on the hot path; it forwards packet symbols and the syncword
    //  tag and drops the inter-packet remainder — left out here] -> CostasLoop (defaults)
    std::vector<float> syncword_bipolar;
    for (auto bit : detection.syncword) syncword_bipolar.push_back(bit ? -1.0f : 1.0f);
    std::vector<c64> wiped, locked;
    TagList t_wo, t_cl;
    gr::packet_modem::SyncwordWipeoffB200 wipeoff;
    wipeoff.syncword = syncword_bipolar;
    wipeoff.settingsChanged(none, none);
    wipeoff.start();
    run_stage("syncword_wipeoff", wipeoff, sym, t_sf, chunk, 1, wiped, t_wo,
              [&](auto& i, auto& o) { return wipeoff.processBulk(i, o); });
    gr::packet_modem::CostasLoopB200 costas;
    costas.settingsChanged(none, none);
    costas.start();
    run_stage("costas_loop", costas, wiped, t_wo, chunk, 1, locked, t_cl,
              [&](auto& i, auto& o) { return costas.processBulk(i, o); });
    {
        gr::packet_modem::CostasLoopB200 fused;
        fused.fused_wipeoff_syncword = syncword_bipolar;
        fused.settingsChanged(none, none);
        fused.start();
        std::vector<c64> lf;
        TagList t_f;
        run_stage("fused_wipeoff_costas_loop", fused, sym, t_sf, chunk, 1, lf, t_f,
                  [&](auto& i, auto& o) { return fused.processBulk(i, o); });
        std::printf("fused_wipeoff_equals_pair %d\n",
                    lf.size() == locked.size() && t_f.size() == t_cl.size() &&
                            std::memcmp(lf.data(), locked.data(), locked.size() * sizeof(c64)) == 0 ? 1 : 0);
        try {
            gr::packet_modem::CostasLoopB200 bad;
            bad.constellation = "8PSK";
            bad.settingsChanged(none, none);
            std::printf("costas_error none\n");
        } catch (const gr::exception& e) { std::printf("costas_error %s\n", e.what()); }
    }
    dump(prefix + "_wiped.cf32", wiped);
    dump(prefix + "_locked.cf32", locked);

    dump(prefix + "_resampled.cf32", y);
    dump(prefix + "_rotated.cf32", z);
    dump(prefix + "_delayed.cf32", d);
    dump(prefix + "_filtered.cf32", f);
    dump(prefix + "_corrected.cf32", g);
    dump(prefix + "_symbols.cf32", sym);
    return 0;
}

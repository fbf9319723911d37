// Drop-in test, literally: ONE templated driver configures and runs a block class through the GR4 block
// contract (same settings members, start()/settingsChanged(), processBulk with consumable / publishable
// spans, merged input tag, out.publishTag) — instantiated once with the REFERENCE's class, compiled unmodified
// from /root/reference, and once with this repository's B200 shell of the same block.  Both build against the
// stand-in GR4 runtime of oracle/ref_stub/ (which also makes this the build that exercises the shells'
// B200SYNC_HAVE_GR4 branch).  The reference's FFT is the oracle's radix-2 (FFTW is not in the tree).
// Built HERE by `make -C oracle ref` into oracle/_ref/dropin_test (the reference tree does not exist on the
// GPU box; the binary travels like the built libraries); run by tests/test_gpu_dropin.py.
//   usage: dropin_test <capture.cf32> <rrc.f32> <sf_taps.f32> <fe_taps.f32>
// Output: one "name key=value ..." line per comparison.
#include <gnuradio-4.0/packet-modem/constellation.hpp>
#include <magic_enum.hpp>
namespace magic_enum {
template <>
struct names<gr::packet_modem::Constellation> {
    static constexpr std::array<std::string_view, 3> value{ "PILOT", "BPSK", "QPSK" };
};
}  // namespace magic_enum
#include <gnuradio-4.0/packet-modem/coarse_frequency_correction.hpp>
#include <gnuradio-4.0/packet-modem/costas_loop.hpp>
#include <gnuradio-4.0/packet-modem/pfb_arb_resampler.hpp>
#include <gnuradio-4.0/packet-modem/symbol_filter.hpp>
#include <gnuradio-4.0/packet-modem/syncword_detection.hpp>
#include <gnuradio-4.0/packet-modem/syncword_wipeoff.hpp>

#include "../../gr4_packet_modem_b200/blocks/coarse_frequency_correction_b200.hpp"
#include "../../gr4_packet_modem_b200/blocks/costas_loop_b200.hpp"
#include "../../gr4_packet_modem_b200/blocks/pfb_arb_resampler_b200.hpp"
#include "../../gr4_packet_modem_b200/blocks/symbol_filter_b200.hpp"
#include "../../gr4_packet_modem_b200/blocks/syncword_detection_b200.hpp"
#include "../../gr4_packet_modem_b200/blocks/syncword_wipeoff_b200.hpp"

#include <cstdio>
#include <cstring>
#include <fstream>

static_assert(B200SYNC_HAVE_GR4 == 1, "the shells must take their GR4 branch here");
using c64 = std::complex<float>;
namespace pm = gr::packet_modem;

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

struct Run {
    std::vector<c64> y;
    std::vector<gr::Tag> tags;  // absolute output indices
    size_t consumed = 0;
};

// the scheduler: chunks cut at input tags, consume()/publish() honoured, tags re-based (GR/Block.hpp:1501-1651)
template <typename Blk>
static Run drive(Blk& b, const std::vector<c64>& x, const std::vector<gr::Tag>& in_tags, size_t chunk, size_t min_in,
                 double out_per_in = 1.0)
{
    Run r;
    size_t pos = 0, ti = 0, stalls = 0;
    std::vector<c64> obuf;
    while (x.size() - pos >= min_in && pos < x.size()) {
        size_t n = std::min(chunk, x.size() - pos);
        size_t tj = ti;
        while (tj < in_tags.size() && static_cast<size_t>(in_tags[tj].index) <= pos) ++tj;
        if (tj < in_tags.size()) n = std::max(min_in, std::min(n, static_cast<size_t>(in_tags[tj].index) - pos));
        n = std::min(n, x.size() - pos);
        b.clear_input_tag();
        gr::property_map merged;
        for (; ti < in_tags.size() && static_cast<size_t>(in_tags[ti].index) == pos; ++ti)
            for (const auto& kv : in_tags[ti].map) merged.insert_or_assign(kv.first, kv.second);
        if (!merged.empty()) b.offer_input_tag(merged);
        obuf.resize(static_cast<size_t>(static_cast<double>(n) * out_per_in) + 8);
        if (out_per_in == 1.0) obuf.resize(n);  // one-to-one blocks assert equal span sizes
        gr::InSpan<c64> is{ std::span<const c64>(x.data() + pos, n) };
        gr::OutSpan<c64> os{ std::span<c64>(obuf) };
        b.out.published_tags.clear();
        const auto st = b.processBulk(is, os);
        if (st != gr::work::Status::OK) break;
        const size_t c = is.consumed(), p = os.published();
        for (const auto& t : b.out.published_tags)
            r.tags.push_back(gr::Tag{ static_cast<ssize_t>(r.y.size()) + t.index, t.map });
        r.y.insert(r.y.end(), obuf.begin(), obuf.begin() + static_cast<std::ptrdiff_t>(p));
        pos += c;
        if (c == 0 && p == 0 && ++stalls > 2) break;
    }
    r.consumed = pos;
    return r;
}

static double rel_l2(const std::vector<c64>& a, const std::vector<c64>& b)
{
    double num = 0, den = 0;
    for (size_t i = 0; i < std::min(a.size(), b.size()); ++i) {
        num += std::norm(std::complex<double>(a[i]) - std::complex<double>(b[i]));
        den += std::norm(std::complex<double>(b[i]));
    }
    return den > 0 ? std::sqrt(num / den) : 0.0;
}
static int same_bits(const std::vector<c64>& a, const std::vector<c64>& b)
{
    return a.size() == b.size() && std::memcmp(a.data(), b.data(), a.size() * sizeof(c64)) == 0;
}
template <typename T>
static T tag_get(const gr::property_map& m, const char* k) { return pmtv::cast<T>(m.at(k)); }

// ---- identical configuration code for the reference class and the B200 shell ----
template <typename Blk>
static void configure_detection(Blk& b, const std::vector<float>& rrc)
{
    static const uint8_t sw[64] = { 0,0,0,0,0,0,1,1,0,1,0,0,0,1,1,1,0,1,1,1,0,1,1,0,1,1,0,0,0,1,1,1,
                                    0,0,1,0,0,1,1,1,0,0,1,0,1,0,0,0,1,0,0,1,0,1,0,1,1,0,1,1,0,0,0,0 };
    b.rrc_taps = rrc;
    b.syncword.assign(sw, sw + 64);
    b.constellation = { { 1.0f, 0.0f }, { -1.0f, 0.0f } };
    b.min_freq_bin = -4;
    b.max_freq_bin = 4;
    b.start();
}
template <typename Blk>
static void configure_symbol_filter(Blk& b, const std::vector<float>& taps, size_t delay)
{
    b.taps = taps;
    b.num_arms = 32;
    b.samples_per_symbol = 4;
    b.delay = delay;
    b.settingsChanged({}, {});
    b.start();
}

int main(int argc, char** argv)
{
    if (argc != 5) return 2;
    const auto x = slurp<c64>(argv[1]);
    const auto rrc = slurp<float>(argv[2]);
    const auto sf_taps = slurp<float>(argv[3]);
    const auto fe_taps = slurp<float>(argv[4]);
    const std::vector<gr::Tag> none;

    // SyncwordDetection
    pm::SyncwordDetection ref_sd;
    pm::SyncwordDetectionB200 gpu_sd;
    configure_detection(ref_sd, rrc);
    configure_detection(gpu_sd, rrc);
    const Run a = drive(ref_sd, x, none, 65536, 2048), g = drive(gpu_sd, x, none, 65536, 2048);
    int idx_equal = a.tags.size() == g.tags.size();
    double dfreq = 0, dphase = 0, dtime = 0, damp = 0;
    int bins_equal = 1;
    for (size_t i = 0; idx_equal && i < a.tags.size(); ++i) {
        idx_equal = a.tags[i].index == g.tags[i].index;
        bins_equal &= tag_get<int>(a.tags[i].map, "syncword_freq_bin") == tag_get<int>(g.tags[i].map, "syncword_freq_bin");
        dfreq = std::max(dfreq, std::abs(tag_get<double>(a.tags[i].map, "syncword_freq") - tag_get<double>(g.tags[i].map, "syncword_freq")));
        const double dp = tag_get<double>(a.tags[i].map, "syncword_phase") - tag_get<double>(g.tags[i].map, "syncword_phase");
        dphase = std::max(dphase, std::abs(std::remainder(dp, 2.0 * M_PI)));
        dtime = std::max(dtime, std::abs(tag_get<double>(a.tags[i].map, "syncword_time_est") - tag_get<double>(g.tags[i].map, "syncword_time_est")));
        damp = std::max(damp, std::abs(tag_get<double>(a.tags[i].map, "syncword_amplitude") / tag_get<double>(g.tags[i].map, "syncword_amplitude") - 1.0));
    }
    std::printf("syncword_detection tags=%zu consumed_equal=%d delayed_bits_equal=%d indices_equal=%d bins_equal=%d "
                "keys=%zu dfreq=%.3g dphase=%.3g dtime=%.3g damp=%.3g\n",
                a.tags.size(), a.consumed == g.consumed, same_bits(a.y, g.y), idx_equal, bins_equal,
                g.tags.empty() ? 0 : g.tags[0].map.size(), dfreq, dphase, dtime, damp);

    // CoarseFrequencyCorrection, fed the reference's detection output and tags
    pm::CoarseFrequencyCorrection<> ref_cfc;
    pm::CoarseFrequencyCorrectionB200 gpu_cfc;
    ref_cfc.delay = gpu_cfc.delay = (rrc.size() - 1) / 2 + 4;
    gpu_cfc.settingsChanged({}, {});
    gpu_cfc.start();
    const Run ca = drive(ref_cfc, a.y, a.tags, 50000, 1), cg = drive(gpu_cfc, a.y, a.tags, 50000, 1);
    std::printf("coarse_frequency_correction n=%zu size_equal=%d rel_l2=%.3g\n", ca.y.size(), ca.y.size() == cg.y.size(),
                rel_l2(cg.y, ca.y));

    // SymbolFilter (exact arithmetic: bit for bit), fed the reference's corrected stream and tags
    pm::SymbolFilter<c64, c64, float> ref_sf;
    pm::SymbolFilterB200 gpu_sf;
    configure_symbol_filter(ref_sf, sf_taps, rrc.size() - 1);
    configure_symbol_filter(gpu_sf, sf_taps, rrc.size() - 1);
    const Run sa = drive(ref_sf, ca.y, a.tags, 50000, 1, 0.25 + 1e-3), sg = drive(gpu_sf, ca.y, a.tags, 50000, 1, 0.25 + 1e-3);
    int sf_tags_equal = sa.tags.size() == sg.tags.size();
    for (size_t i = 0; sf_tags_equal && i < sa.tags.size(); ++i)
        sf_tags_equal = sa.tags[i].index == sg.tags[i].index &&
                        tag_get<float>(sa.tags[i].map, "syncword_phase") == tag_get<float>(sg.tags[i].map, "syncword_phase");
    std::printf("symbol_filter symbols=%zu bits_equal=%d tags=%zu tags_equal=%d\n", sa.y.size(), same_bits(sa.y, sg.y),
                sa.tags.size(), sf_tags_equal);

    // SyncwordWipeoff (exact) and CostasLoop (libm vs the kernel's sincos: tolerance)
    std::vector<float> bipolar;
    for (auto bit : ref_sd.syncword) bipolar.push_back(bit ? -1.0f : 1.0f);
    pm::SyncwordWipeoff<> ref_wo;
    pm::SyncwordWipeoffB200 gpu_wo;
    ref_wo.syncword = gpu_wo.syncword = bipolar;
    gpu_wo.settingsChanged({}, {});
    gpu_wo.start();
    const Run wa = drive(ref_wo, sa.y, sa.tags, 50000, 1), wg = drive(gpu_wo, sa.y, sa.tags, 50000, 1);
    std::printf("syncword_wipeoff n=%zu bits_equal=%d\n", wa.y.size(), same_bits(wa.y, wg.y));
    pm::CostasLoop<> ref_cl;
    pm::CostasLoopB200 gpu_cl;
    ref_cl.settingsChanged({}, {});
    gpu_cl.settingsChanged({}, {});
    gpu_cl.start();
    const Run la = drive(ref_cl, wa.y, sa.tags, 50000, 1), lg = drive(gpu_cl, wa.y, sa.tags, 50000, 1);
    std::printf("costas_loop n=%zu size_equal=%d rel_l2=%.3g\n", la.y.size(), la.y.size() == lg.y.size(), rel_l2(lg.y, la.y));

    // PfbArbResampler (exact arithmetic: bit for bit)
    pm::PfbArbResampler<c64, c64, float, float> ref_rs;
    pm::PfbArbResamplerB200 gpu_rs;
    ref_rs.rate = gpu_rs.rate = 1.0f + 1e-6f * 1.2f;
    ref_rs.taps = gpu_rs.taps = fe_taps;
    ref_rs.filter_size = gpu_rs.filter_size = 32;
    ref_rs.settingsChanged({}, {});
    gpu_rs.settingsChanged({}, {});
    const Run ra = drive(ref_rs, x, none, 40000, 1, 1.01), rg = drive(gpu_rs, x, none, 40000, 1, 1.01);
    std::printf("pfb_arb_resampler outputs=%zu consumed_equal=%d bits_equal=%d\n", ra.y.size(), ra.consumed == rg.consumed,
                same_bits(ra.y, rg.y));
    return 0;
}

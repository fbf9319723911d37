// =============================================================================
// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// oracle/_ref/librefblocks.so: the REFERENCE's own RX-synchronisation block headers, compiled UNMODIFIED from
// where they lie under /root/reference (blocks/include/gnuradio-4.0/packet-modem/*.hpp plus the real
// gnuradio-4.0/HistoryBuffer.hpp) against the stand-in runtime of oracle/ref_stub/ (the real GR4 runtime
// needs network-fetched dependencies).  This file plays the scheduler: it offers chunks, hands
// over the merged input tag of a chunk's first item and honours consume()/publish().  The one piece that is
// not the reference's is the FFT (FFTW 3.3.10 is not in the tree): the stub FFTw uses the oracle's radix-2
// arithmetic, so the reference block and the oracle's restatement must agree BIT FOR BIT — which is what
// tests/test_oracle_vs_reference_blocks.py checks, and what tests/golden/ref_blocks_golden.npz carries to
// the GPU box.  No reference source is copied into this repository.
// Build: make -C oracle refblocks   (g++ -std=c++23 -ffp-contract=off; only when /root/reference exists)
// =============================================================================
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
#include <gnuradio-4.0/packet-modem/interpolating_fir_filter.hpp>
#include <gnuradio-4.0/packet-modem/pfb_arb_resampler.hpp>
#include <gnuradio-4.0/packet-modem/rotator.hpp>
#include <gnuradio-4.0/packet-modem/symbol_filter.hpp>
#include <gnuradio-4.0/packet-modem/syncword_detection.hpp>
#include <gnuradio-4.0/packet-modem/syncword_detection_filter.hpp>
#include <gnuradio-4.0/packet-modem/syncword_wipeoff.hpp>

#include <cstring>
#include <memory>

using c64 = std::complex<float>;
namespace pm = gr::packet_modem;

namespace {
struct RefTag {  // what crosses the C boundary for a syncword tag
    int64_t index;
    double freq;
    float amplitude, phase, noise_power, esn0_db, time_est;
    int32_t freq_bin;
    int32_t no_syncword;  // != 0: the tag carries NO syncword_* keys
    int32_t other;        // != 0: the tag carries a non-syncword key ("other_key" = other)
};
template <typename T>
T get(const gr::property_map& m, const char* k, T dflt = T{})
{
    auto it = m.find(k);
    return it == m.end() ? dflt : pmtv::cast<T>(it->second);
}
RefTag to_ref_tag(int64_t index, const gr::property_map& m)
{
    RefTag t{};
    t.index = index;
    t.freq = get<double>(m, "syncword_freq");
    t.amplitude = get<float>(m, "syncword_amplitude");
    t.phase = get<float>(m, "syncword_phase");
    t.noise_power = get<float>(m, "syncword_noise_power");
    t.esn0_db = get<float>(m, "syncword_esn0_db");
    t.time_est = get<float>(m, "syncword_time_est");
    t.freq_bin = get<int>(m, "syncword_freq_bin");
    t.no_syncword = m.contains("syncword_amplitude") ? 0 : 1;
    t.other = get<int>(m, "other_key");
    return t;
}
gr::property_map from_ref_tag(const RefTag& t)
{
    gr::property_map m;
    if (!t.no_syncword)
        m = gr::property_map{ { "syncword_amplitude", t.amplitude }, { "syncword_phase", t.phase },
                              { "syncword_freq", t.freq },           { "syncword_freq_bin", t.freq_bin },
                              { "syncword_noise_power", t.noise_power }, { "syncword_esn0_db", t.esn0_db },
                              { "syncword_time_est", t.time_est } };
    if (t.other) m["other_key"] = t.other;
    return m;
}
// one processBulk call: chunk [in, in + n_in) -> out (capacity max_out); tag = merged tag of the first item
template <typename Blk>
int call(Blk& b, const c64* in, size_t n_in, c64* out, size_t max_out, const RefTag* tag, size_t* consumed,
         size_t* produced, RefTag* otags, size_t max_otags, size_t* n_otags)
{
    b.clear_input_tag();
    if (tag) b.offer_input_tag(from_ref_tag(*tag));
    gr::InSpan<c64> is{ std::span<const c64>(in, n_in) };
    gr::OutSpan<c64> os{ std::span<c64>(out, max_out) };
    b.out.published_tags.clear();
    try {
        const auto st = b.processBulk(is, os);
        *consumed = st == gr::work::Status::OK ? is.consumed() : 0;   // GR/Block.hpp:1620-1624
        *produced = st == gr::work::Status::OK ? os.published() : 0;
        size_t nt = 0;
        for (const auto& t : b.out.published_tags)
            if (nt < max_otags) otags[nt++] = to_ref_tag(t.index, t.map);
        if (n_otags) *n_otags = nt;
        return static_cast<int>(st);
    } catch (const std::exception&) {
        return -1000;
    }
}
}  // namespace

extern "C" {

size_t refblk_sizeof_tag() { return sizeof(RefTag); }

// ---- SyncwordDetection (PM/syncword_detection.hpp) ----
void* refblk_sd_create(const float* rrc, size_t n_rrc, const uint8_t* sw, size_t n_sw, const float* constellation,
                       size_t n_const, int min_bin, int max_bin, size_t time_threshold, float power_threshold,
                       size_t fft_size)
{
    auto b = std::make_unique<pm::SyncwordDetection>();
    if (fft_size != 0) b->fft_size = fft_size;   // PM/syncword_detection.hpp:133 (default 2048)
    b->rrc_taps.assign(rrc, rrc + n_rrc);
    b->syncword.assign(sw, sw + n_sw);
    b->constellation.clear();
    for (size_t i = 0; i < n_const; ++i) b->constellation.emplace_back(constellation[2 * i], constellation[2 * i + 1]);
    b->min_freq_bin = min_bin;
    b->max_freq_bin = max_bin;
    b->time_threshold = time_threshold;
    b->power_threshold = power_threshold;
    try {
        b->start();
    } catch (const std::exception&) {
        return nullptr;
    }
    return b.release();
}
void refblk_sd_destroy(void* h) { delete static_cast<pm::SyncwordDetection*>(h); }
int refblk_sd_process(void* h, const float* in, size_t n_in, float* out, size_t* consumed, RefTag* tags,
                      size_t max_tags, size_t* n_tags)
{
    size_t produced = 0;
    return call(*static_cast<pm::SyncwordDetection*>(h), reinterpret_cast<const c64*>(in), n_in,
                reinterpret_cast<c64*>(out), n_in, nullptr, consumed, &produced, tags, max_tags, n_tags);
}

// ---- Rotator (PM/rotator.hpp) ----
void refblk_rotator(float phase_incr, const float* in, size_t n, float* out)
{
    pm::Rotator<> r;
    r.phase_incr = phase_incr;
    r.settingsChanged({}, {});
    const c64* i = reinterpret_cast<const c64*>(in);
    c64* o = reinterpret_cast<c64*>(out);
    for (size_t k = 0; k < n; ++k) o[k] = r.processOne(i[k]);
}

// ---- CoarseFrequencyCorrection / SyncwordWipeoff / CostasLoop: chunk in, chunk out, optional tag ----
void* refblk_cfc_create(size_t delay)
{
    auto b = std::make_unique<pm::CoarseFrequencyCorrection<>>();
    b->delay = delay;
    return b.release();
}
void refblk_cfc_destroy(void* h) { delete static_cast<pm::CoarseFrequencyCorrection<>*>(h); }
int refblk_cfc_process(void* h, const float* in, size_t n, float* out, const RefTag* tag)
{
    size_t c = 0, p = 0;
    return call(*static_cast<pm::CoarseFrequencyCorrection<>*>(h), reinterpret_cast<const c64*>(in), n,
                reinterpret_cast<c64*>(out), n, tag, &c, &p, nullptr, 0, nullptr);
}

void* refblk_wo_create(const float* syncword, size_t n)
{
    auto b = std::make_unique<pm::SyncwordWipeoff<>>();
    b->syncword.assign(syncword, syncword + n);
    return b.release();
}
void refblk_wo_destroy(void* h) { delete static_cast<pm::SyncwordWipeoff<>*>(h); }
int refblk_wo_process(void* h, const float* in, size_t n, float* out, const RefTag* tag)
{
    size_t c = 0, p = 0;
    return call(*static_cast<pm::SyncwordWipeoff<>*>(h), reinterpret_cast<const c64*>(in), n,
                reinterpret_cast<c64*>(out), n, tag, &c, &p, nullptr, 0, nullptr);
}

void* refblk_cl_create(double loop_bandwidth, const char* constellation)
{
    auto b = std::make_unique<pm::CostasLoop<>>();
    b->loop_bandwidth = loop_bandwidth;
    b->constellation = constellation;
    try {
        b->settingsChanged({}, {});
    } catch (const std::exception&) {
        return nullptr;
    }
    return b.release();
}
void refblk_cl_destroy(void* h) { delete static_cast<pm::CostasLoop<>*>(h); }
int refblk_cl_process(void* h, const float* in, size_t n, float* out, const RefTag* tag)
{
    size_t c = 0, p = 0;
    return call(*static_cast<pm::CostasLoop<>*>(h), reinterpret_cast<const c64*>(in), n,
                reinterpret_cast<c64*>(out), n, tag, &c, &p, nullptr, 0, nullptr);
}
void refblk_cl_state(void* h, float* phase, float* freq, float* k1, float* k2)
{
    auto* b = static_cast<pm::CostasLoop<>*>(h);
    *phase = b->_phase;
    *freq = b->_freq;
    *k1 = b->_k1;
    *k2 = b->_k2;
}

// ---- SyncwordDetectionFilter (PM/syncword_detection_filter.hpp): two message inputs, one stream ----
void* refblk_sdf_create(size_t sps, size_t syncword_size, size_t header_size)
{
    auto b = std::make_unique<pm::SyncwordDetectionFilter<>>();
    b->samples_per_symbol = sps;
    b->syncword_size = syncword_size;
    b->header_size = header_size;
    b->start();
    return b.release();
}
void refblk_sdf_destroy(void* h) { delete static_cast<pm::SyncwordDetectionFilter<>*>(h); }
// header_kind: 0 none, 1 parsed header with packet_length, 2 invalid_header; returns items consumed or -1
long long refblk_sdf_process(void* h, int header_kind, uint64_t packet_length, size_t n_ignored, const float* in,
                             size_t n_in, float* out, size_t n_out, const RefTag* tag_in, size_t* hdr_used,
                             size_t* ign_used, RefTag* tag_out, int* tag_forwarded, int* in_packet)
{
    auto& b = *static_cast<pm::SyncwordDetectionFilter<>*>(h);
    std::vector<gr::Message> hdr, ign(n_ignored);
    if (header_kind == 1) hdr.push_back(gr::Message{ gr::property_map{ { "packet_length", packet_length } } });
    if (header_kind == 2) hdr.push_back(gr::Message{ gr::property_map{ { "invalid_header", true } } });
    b.clear_input_tag();
    if (tag_in) b.offer_input_tag(from_ref_tag(*tag_in));
    gr::InSpan<gr::Message> hs{ std::span<const gr::Message>(hdr) }, gs{ std::span<const gr::Message>(ign) };
    gr::InSpan<c64> is{ std::span<const c64>(reinterpret_cast<const c64*>(in), n_in) };
    gr::OutSpan<c64> os{ std::span<c64>(reinterpret_cast<c64*>(out), n_out) };
    b.out.published_tags.clear();
    try {
        if (b.processBulk(hs, gs, is, os) != gr::work::Status::OK) return -1;
    } catch (const std::exception&) {
        return -1;
    }
    *hdr_used = hs.consumed();
    *ign_used = gs.consumed();
    *tag_forwarded = b.out.published_tags.empty() ? 0 : 1;
    if (*tag_forwarded) *tag_out = to_ref_tag(b.out.published_tags[0].index, b.out.published_tags[0].map);
    *in_packet = b._in_packet ? 1 : 0;
    return static_cast<long long>(is.consumed());
}

// ---- SymbolFilter (PM/symbol_filter.hpp) ----
using RefSymbolFilter = pm::SymbolFilter<c64, c64, float>;
void* refblk_sf_create(const float* taps, size_t n_taps, size_t num_arms, size_t sps, size_t delay)
{
    auto b = std::make_unique<RefSymbolFilter>();
    b->taps.assign(taps, taps + n_taps);
    b->num_arms = num_arms;
    b->samples_per_symbol = sps;
    b->delay = delay;
    try {
        b->settingsChanged({}, {});
        b->start();
    } catch (const std::exception&) {
        return nullptr;
    }
    return b.release();
}
void refblk_sf_destroy(void* h) { delete static_cast<RefSymbolFilter*>(h); }
int refblk_sf_process(void* h, const float* in, size_t n_in, float* out, size_t max_out, const RefTag* tag,
                      size_t* consumed, size_t* produced, RefTag* otags, size_t max_otags, size_t* n_otags)
{
    return call(*static_cast<RefSymbolFilter*>(h), reinterpret_cast<const c64*>(in), n_in,
                reinterpret_cast<c64*>(out), max_out, tag, consumed, produced, otags, max_otags, n_otags);
}

// ---- PfbArbResampler<c64, c64, float, float> (PM/pfb_arb_resampler.hpp) ----
using RefResampler = pm::PfbArbResampler<c64, c64, float, float>;
void* refblk_rs_create(float rate, const float* taps, size_t n_taps, size_t filter_size)
{
    auto b = std::make_unique<RefResampler>();
    b->rate = rate;
    b->taps.assign(taps, taps + n_taps);
    b->filter_size = filter_size;
    try {
        b->settingsChanged({}, {});
    } catch (const std::exception&) {
        return nullptr;
    }
    return b.release();
}
void refblk_rs_destroy(void* h) { delete static_cast<RefResampler*>(h); }
int refblk_rs_process(void* h, const float* in, size_t n_in, float* out, size_t max_out, size_t* consumed,
                      size_t* produced)
{
    return call(*static_cast<RefResampler*>(h), reinterpret_cast<const c64*>(in), n_in, reinterpret_cast<c64*>(out),
                max_out, nullptr, consumed, produced, nullptr, 0, nullptr);
}

// ---- InterpolatingFirFilter<c64, c64, float> (PM/interpolating_fir_filter.hpp), fresh block ----
int refblk_interp_fir(const float* taps, size_t n_taps, size_t interpolation, const float* in, size_t n_in, float* out)
{
    pm::InterpolatingFirFilter<c64, c64, float> b;
    b.interpolation = interpolation;
    b.taps.assign(taps, taps + n_taps);
    try {
        b.settingsChanged({}, {});
    } catch (const std::exception&) {
        return -1000;
    }
    size_t c = 0, p = 0;
    return call(b, reinterpret_cast<const c64*>(in), n_in, reinterpret_cast<c64*>(out), n_in * interpolation, nullptr,
                &c, &p, nullptr, 0, nullptr);
}

}  // extern "C"

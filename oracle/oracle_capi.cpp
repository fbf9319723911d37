// =============================================================================
// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle.hpp).
// extern "C" surface so tests/, smoke() and bench.py's CPU legs can drive the
// restated reference blocks through ctypes.
// =============================================================================
#include "oracle.hpp"

#include <cstring>
#include <memory>

using namespace orc;

extern "C" {

int orc_rrc(double gain, double fs, double symrate, double alpha, size_t ntaps, float* out, size_t max_out)
{
    const auto t = root_raised_cosine(gain, fs, symrate, alpha, ntaps);
    if (t.size() > max_out) return -1;
    std::memcpy(out, t.data(), t.size() * sizeof(float));
    return static_cast<int>(t.size());
}

// ---- FFT (2048 only for Mirror) ----
int orc_fft(int kind, int which, size_t n, const float* in, float* out)
{
    try {
        Fft f(n, static_cast<FftKind>(kind));
        if (which == 0) f.forward(reinterpret_cast<const c64*>(in), reinterpret_cast<c64*>(out));
        else f.second(reinterpret_cast<const c64*>(in), reinterpret_cast<c64*>(out));
        return 0;
    } catch (...) {
        return -1;
    }
}

// ---- SyncwordDetection ----
void* orc_sd_create(size_t fft_size, size_t sps, const float* rrc, size_t nrrc, const uint8_t* sw, size_t nsw,
                    const float* constel, size_t nconst, int min_bin, int max_bin, uint64_t time_threshold,
                    float power_threshold, int fft_kind, int record_metric)
{
    try {
        auto sd = std::make_unique<SyncwordDetection>();
        sd->fft_size = fft_size;
        sd->samples_per_symbol = sps;
        sd->rrc_taps.assign(rrc, rrc + nrrc);
        sd->syncword.assign(sw, sw + nsw);
        sd->constellation.resize(nconst);
        for (size_t i = 0; i < nconst; ++i) sd->constellation[i] = c64(constel[2 * i], constel[2 * i + 1]);
        sd->min_freq_bin = min_bin;
        sd->max_freq_bin = max_bin;
        sd->time_threshold = time_threshold;
        sd->power_threshold = power_threshold;
        sd->fft_kind = static_cast<FftKind>(fft_kind);
        sd->record_metric = record_metric != 0;
        sd->start();
        return sd.release();
    } catch (...) {
        return nullptr;
    }
}
void orc_sd_destroy(void* h) { delete static_cast<SyncwordDetection*>(h); }
void orc_sd_info(void* h, uint32_t* L, float* self_corr)
{
    auto* sd = static_cast<SyncwordDetection*>(h);
    *L = static_cast<uint32_t>(sd->_syncword_samples_size);
    *self_corr = sd->_syncword_self_corr;
}
// returns items consumed (== published), -1 on error / tag overflow
long long orc_sd_process(void* h, const float* in, size_t n, float* out, SyncwordTag* tags, size_t max_tags,
                         size_t* n_tags)
{
    auto* sd = static_cast<SyncwordDetection*>(h);
    std::vector<SyncwordTag> t;
    std::vector<c64> scratch;
    c64* o = reinterpret_cast<c64*>(out);
    if (!o) {
        scratch.resize(n);
        o = scratch.data();
    }
    const size_t c = sd->processBulk(reinterpret_cast<const c64*>(in), n, o, t);
    if (t.size() > max_tags) return -1;
    if (!t.empty()) std::memcpy(tags, t.data(), t.size() * sizeof(SyncwordTag));
    *n_tags = t.size();
    return static_cast<long long>(c);
}
size_t orc_sd_metric(void* h, float* pow, int8_t* bin, size_t max_n)
{
    auto* sd = static_cast<SyncwordDetection*>(h);
    const size_t n = std::min(max_n, sd->metric_pow.size());
    if (pow) std::memcpy(pow, sd->metric_pow.data(), n * sizeof(float));
    if (bin) std::memcpy(bin, sd->metric_bin.data(), n * sizeof(int8_t));
    return n;
}
int orc_sd_template(void* h, size_t k, float* out)
{
    auto* sd = static_cast<SyncwordDetection*>(h);
    if (k >= sd->_syncword_fft_conj.size()) return -1;
    std::memcpy(out, sd->_syncword_fft_conj[k].data(), sd->fft_size * sizeof(c64));
    return 0;
}

// ---- Rotator (fresh block, n samples) ----
void orc_rotator(float phase_incr, const float* in, size_t n, float* out)
{
    Rotator r;
    r.phase_incr = phase_incr;
    r.settingsChanged();
    r.start();
    const c64* i = reinterpret_cast<const c64*>(in);
    c64* o = reinterpret_cast<c64*>(out);
    for (size_t k = 0; k < n; ++k) o[k] = r.processOne(i[k]);
}

// ---- CoarseFrequencyCorrection ----
void* orc_cfc_create(size_t delay)
{
    auto c = std::make_unique<CoarseFrequencyCorrection>();
    c->delay = delay;
    return c.release();
}
void orc_cfc_destroy(void* h) { delete static_cast<CoarseFrequencyCorrection*>(h); }
// one chunk; has_freq != 0: the chunk's first sample carries a tag with syncword_freq = freq
void orc_cfc_process(void* h, const float* in, size_t n, float* out, int has_freq, double freq)
{
    static_cast<CoarseFrequencyCorrection*>(h)->processBulk(reinterpret_cast<const c64*>(in), n,
                                                            reinterpret_cast<c64*>(out), has_freq != 0, freq);
}

// ---- SyncwordWipeoff ----
void* orc_wo_create(const float* syncword, size_t n)
{
    auto w = std::make_unique<SyncwordWipeoff>();
    w->syncword.assign(syncword, syncword + n);
    return w.release();
}
void orc_wo_destroy(void* h) { delete static_cast<SyncwordWipeoff*>(h); }
// one chunk; has_tag != 0: the chunk's first sample carries a tag with a syncword_amplitude key
void orc_wo_process(void* h, const float* in, size_t n, float* out, int has_tag)
{
    static_cast<SyncwordWipeoff*>(h)->processBulk(reinterpret_cast<const c64*>(in), n, reinterpret_cast<c64*>(out),
                                                  has_tag != 0);
}

// ---- CostasLoop ----
void* orc_cl_create(double loop_bandwidth, int constellation, int trig_kind)
{
    auto c = std::make_unique<CostasLoop>();
    c->loop_bandwidth = loop_bandwidth;
    c->constellation = static_cast<CostasLoop::Constellation>(constellation);
    c->trig = static_cast<TrigKind>(trig_kind);
    c->settingsChanged();
    return c.release();
}
void orc_cl_destroy(void* h) { delete static_cast<CostasLoop*>(h); }
// one chunk; has_phase != 0: the chunk's first sample carries a tag with syncword_phase = phase
void orc_cl_process(void* h, const float* in, size_t n, float* out, int has_phase, float phase)
{
    static_cast<CostasLoop*>(h)->processBulk(reinterpret_cast<const c64*>(in), n, reinterpret_cast<c64*>(out),
                                             has_phase != 0, phase);
}
void orc_cl_state(void* h, float* phase, float* freq, float* k1, float* k2)
{
    auto* c = static_cast<CostasLoop*>(h);
    *phase = c->_phase;
    *freq = c->_freq;
    *k1 = c->_k1;
    *k2 = c->_k2;
}
void orc_mirror_sincosf(const float* x, size_t n, float* s, float* c)
{
    for (size_t i = 0; i < n; ++i) mirror_sincosf(x[i], s[i], c[i]);
}

// ---- PfbArbResampler ----
struct ResamplerBox {
    bool dbl;
    PfbArbResampler<float> f;
    PfbArbResampler<double> d;
    uint64_t in_total = 0;
};
void* orc_resampler_create(double rate, int use_double, const float* taps, size_t ntaps, size_t filter_size)
{
    try {
        auto b = std::make_unique<ResamplerBox>();
        b->dbl = use_double != 0;
        if (b->dbl) {
            b->d.rate = rate;
            b->d.taps.assign(taps, taps + ntaps);
            b->d.filter_size = filter_size;
            b->d.settingsChanged();
        } else {
            b->f.rate = static_cast<float>(rate);
            b->f.taps.assign(taps, taps + ntaps);
            b->f.filter_size = filter_size;
            b->f.settingsChanged();
        }
        return b.release();
    } catch (...) {
        return nullptr;
    }
}
void orc_resampler_destroy(void* h) { delete static_cast<ResamplerBox*>(h); }
// arms / in_counts / accs may be NULL; otherwise they must hold n_out entries
void orc_resampler_process(void* h, const float* in, size_t n_in, float* out, size_t n_out, size_t* consumed,
                           size_t* produced, uint32_t* arms, uint64_t* in_counts, double* accs)
{
    auto* b = static_cast<ResamplerBox*>(h);
    std::vector<uint32_t> va;
    std::vector<uint64_t> vc;
    std::vector<double> vp;
    if (b->dbl)
        b->d.processBulk(reinterpret_cast<const c64*>(in), n_in, reinterpret_cast<c64*>(out), n_out, *consumed,
                         *produced, arms ? &va : nullptr, in_counts ? &vc : nullptr, accs ? &vp : nullptr,
                         b->in_total);
    else
        b->f.processBulk(reinterpret_cast<const c64*>(in), n_in, reinterpret_cast<c64*>(out), n_out, *consumed,
                         *produced, arms ? &va : nullptr, in_counts ? &vc : nullptr, accs ? &vp : nullptr,
                         b->in_total);
    b->in_total += *consumed;
    if (arms) std::memcpy(arms, va.data(), va.size() * sizeof(uint32_t));
    if (in_counts) std::memcpy(in_counts, vc.data(), vc.size() * sizeof(uint64_t));
    if (accs) std::memcpy(accs, vp.data(), vp.size() * sizeof(double));
}

// ---- SymbolFilter ----
void* orc_symfilt_create(size_t sps, const float* taps, size_t ntaps, size_t num_arms, size_t delay)
{
    try {
        auto s = std::make_unique<SymbolFilter>();
        s->samples_per_symbol = sps;
        s->taps.assign(taps, taps + ntaps);
        s->num_arms = num_arms;
        s->delay = delay;
        s->settingsChanged();
        s->start();
        return s.release();
    } catch (...) {
        return nullptr;
    }
}
void orc_symfilt_destroy(void* h) { delete static_cast<SymbolFilter*>(h); }
// tag_in may be NULL.  Returns number of output tags written, -1 on overflow.
int orc_symfilt_process(void* h, const float* in, size_t n_in, float* out, size_t n_out, const StreamTag* tag_in,
                        size_t* consumed, size_t* produced, StreamTag* out_tags, size_t max_tags)
{
    auto* s = static_cast<SymbolFilter*>(h);
    std::vector<StreamTag> t;
    s->processBulk(reinterpret_cast<const c64*>(in), n_in, reinterpret_cast<c64*>(out), n_out, tag_in, *consumed,
                   *produced, t);
    if (t.size() > max_tags) return -1;
    if (!t.empty()) std::memcpy(out_tags, t.data(), t.size() * sizeof(StreamTag));
    return static_cast<int>(t.size());
}

// ---- SyncwordDetectionFilter ----
void* orc_sdf_create(size_t sps, size_t syncword_size, size_t header_size)
{
    auto s = std::make_unique<SyncwordDetectionFilter>();
    s->samples_per_symbol = sps;
    s->syncword_size = syncword_size;
    s->header_size = header_size;
    s->start();
    return s.release();
}
void orc_sdf_destroy(void* h) { delete static_cast<SyncwordDetectionFilter*>(h); }
// header_kind: 0 none, 1 parsed header with packet_length, 2 invalid_header
long long orc_sdf_process(void* h, int header_kind, uint64_t packet_length, size_t n_ignored, const float* in,
                          size_t n_in, float* out, size_t n_out, const StreamTag* tag_in, size_t* hdr_used,
                          size_t* ign_used, StreamTag* tag_out, int* tag_forwarded, int* in_packet)
{
    auto* s = static_cast<SyncwordDetectionFilter*>(h);
    HeaderMsg m;
    m.invalid_header = header_kind == 2;
    m.packet_length = packet_length;
    bool fwd = false;
    try {
        const size_t c = s->processBulk(&m, header_kind ? 1 : 0, n_ignored, reinterpret_cast<const c64*>(in), n_in,
                                        reinterpret_cast<c64*>(out), n_out, tag_in, *hdr_used, *ign_used, tag_out,
                                        fwd);
        *tag_forwarded = fwd ? 1 : 0;
        *in_packet = s->_in_packet ? 1 : 0;
        return static_cast<long long>(c);
    } catch (...) {
        return -1;
    }
}

// ---- InterpolatingFirFilter (stimulus) ----
void orc_interp_fir(const float* taps, size_t ntaps, size_t interpolation, const float* in, size_t n_in, float* out)
{
    interpolating_fir(std::vector<float>(taps, taps + ntaps), interpolation, reinterpret_cast<const c64*>(in), n_in,
                      reinterpret_cast<c64*>(out));
}

size_t orc_sizeof_syncword_tag() { return sizeof(SyncwordTag); }
size_t orc_sizeof_stream_tag() { return sizeof(StreamTag); }

} // extern "C"

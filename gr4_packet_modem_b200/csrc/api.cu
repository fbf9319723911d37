// api.cu — host side of libb200sync.so: the C ABI of include/b200sync.h.
//
// Host responsibilities (everything else runs on the GPU):
//   * SyncwordDetection::start()  — settings validation and construction of the modulated,
//     frequency-shifted syncword in the reference's own float/double arithmetic
//     (PM/syncword_detection.hpp:143-182); the spectra themselves are computed by the GPU FFT;
//   * the delay line of the block contract when the spans live in host memory
//     (PM/syncword_detection.hpp:318-319) — a memcpy, the samples never needed the GPU;
//   * SyncwordDetection::output_tag() (PM/syncword_detection.hpp:56-115) on the raw
//     per-detection records the GPU returns;
//   * buffer management and the streaming bookkeeping (_items_consumed, pending tags).
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <complex>
#include <cstdio>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <new>
#include <numbers>
#include <string>
#include <thread>
#include <vector>

#include "b200sync_internal.h"

using namespace b200sync;

namespace {

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{ 0 };

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
#define CU(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(B200SYNC_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));       \
    } while (0)

using c64 = std::complex<float>;

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
};

}  // namespace

struct b200sync_sd {
    // settings (PM/syncword_detection.hpp:131-141)
    uint32_t fft_size = 2048, sps = 4;
    std::vector<float> rrc_taps;
    std::vector<uint8_t> syncword;
    std::vector<c64> constellation;
    int min_bin = 0, max_bin = 0;
    uint64_t time_threshold = 768;
    float power_threshold = 9.5f;
    int device = 0, num_sms = 148;
    // derived (:148, :161-164, :236)
    uint32_t L = 0, S = 0, K = 0;
    bool generic = false;   // correlator_generic.cu: every fft_size but 2048 (and 2048 under B200SYNC_FORCE_GENERIC)
    int fft_arg = kFft;     // what the launchers get: fft_size, | kGenericFlag for a forced 2048
    float self_corr = 0.0f;
    int T = 0;
    uint64_t delay = 0;
    // device constants
    DevBuf<float2> d_tw, d_hperm;
    // peak state + detection lists
    DevBuf<PeakState> d_state;
    DevBuf<unsigned long long> d_det_idx;
    DevBuf<DetectionRecord> d_recs;
    DevBuf<unsigned char> d_ws;
    std::vector<DetectionRecord> h_recs;
    PeakState* h_state = nullptr;  // pinned
    // streaming (process): staging windows with absolute origins
    cudaStream_t stream = nullptr;
    DevBuf<float2> d_x, d_xtmp;
    long long x_base = 0, x_end = 0;
    DevBuf<float> d_z, d_ztmp;
    long long z_base = 0, z_end = 0;
    uint64_t consumed = 0;  // _items_consumed
    long long lo_next = 0;  // first undecided sample
    unsigned long long r_abs_host = 0;  // host copy of the search position (PeakState::r_abs), streaming path
    // streaming fast path: pinned staging for pageable input spans, pinned landing buffer for the records
    unsigned char* h_in_stage = nullptr;
    void* d_in_stage_view = nullptr;
    // auto-registration of pageable input spans (b200sync_sd_set_auto_register): disjoint page-aligned ranges this context
    // has page-locked, sorted by address; ranges whose registration failed once are not tried again
    int auto_register = -1;            // -1: take B200SYNC_AUTO_REGISTER from the environment at first use
    std::vector<std::pair<uintptr_t, uintptr_t>> reg_ranges, reg_failed;   // device-visible address of h_in_stage (zero-copy input of the streaming path)
    DetectionRecord* h_recs_pin = nullptr;
    size_t h_recs_pin_cap = 0;
    DevBuf<unsigned int> d_done;       // CTA completion counter of the streaming refine launch (StreamWalk::done)
    unsigned int stream_seq = 0;       // sequence number of the last streaming step (the completion flag's value)
    std::vector<c64> carry; // last `delay` input samples (delay line)
    std::deque<b200sync_sd_tag> pending;
    // offline: metric, and the correlator's group extrema of it for the peak stage (gm_* in b200sync_internal.h)
    DevBuf<float2> d_gm;
    long long gm_b0 = 0, gm_blocks = 0;   // rows of d_gm valid for the current offline / shard call (0: none)
    DevBuf<float2> d_chan_gm;
    DevBuf<float> d_zoff;
    DevBuf<float2> d_xoff;
    size_t metric_n = 0;
    const float* metric_ptr = nullptr;
    long long metric_base = 0;
    // device-side stage timing of the last offline call (CUDA events on the caller's stream)
    cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
    bool ev_valid = false;
    // batched channel mode: metric, workspace and detection lists of all channels of a group side by side
    DevBuf<float> d_chan_z;
    DevBuf<unsigned char> d_chan_ws;
    DevBuf<unsigned long long> d_chan_det;
    DevBuf<PeakState> d_chan_state;
    DevBuf<DetectionRecord> d_chan_recs;
    PeakState* h_chan_state = nullptr;  // pinned
    size_t h_chan_cap = 0;
    // raw capture ingestion (b200sync_sd_detect_file): pinned staging ring + copy stream
    float2* h_stage = nullptr;
    cudaEvent_t ev_stage[3] = {nullptr, nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> ev_pieces;  // b200sync_sd_shard_phase1_host: one event per H2D piece
    // host output span of the bulk host-span calls when it is pinned: the correlator writes the delayed stream into
    // d_outoff and a second copy stream sends finished pieces back while later pieces are still coming up (PCIe is
    // full duplex; a host-side memcpy would cost the host memory system twice the bytes)
    DevBuf<float2> d_outoff;
    cudaStream_t d2h_stream = nullptr;
    std::vector<cudaEvent_t> ev_out;
    // shard
    DevBuf<uint16_t> d_table;
    struct {
        const float2* d_in = nullptr;
        long long in_base = 0, z_base = 0, lo = 0, hi = 0, P_total = 0;
        cudaStream_t st = nullptr;
        bool valid = false;
        float2* d_out = nullptr;      // optional delayed output of the NEXT phase 1 (b200sync_sd_shard_output)
        long long out_first = 0, out_len = 0;
        float* h_out = nullptr;       // ... or into HOST memory, for phase1_host (b200sync_sd_shard_output_host)
        long long h_out_first = 0, h_out_len = 0;
    } shard;
};

namespace {

constexpr int kTwTotalHost = 256 + 2048;  // == kTwTotal of fft2048.cuh
constexpr long long kSmallRange = 1LL << 19;      // peak ranges up to this size take the two-launch in-order walk (bitmaps in shared memory)
constexpr size_t kStageBytes = 4u << 20;          // pinned staging for pageable input spans (streaming)
constexpr size_t kStagePiece = 128u << 10;        // ... copied and sent piece by piece so memcpy and DMA overlap

int reset_state(b200sync_sd* sd, cudaStream_t st) {
    CU(sd->d_state.ensure(1));
    CU(cudaMemsetAsync(sd->d_state.p, 0, sizeof(PeakState), st));
    return 0;
}

int ensure_det(b200sync_sd* sd, size_t cap) {
    if (cap < 64) cap = 64;
    CU(sd->d_det_idx.ensure(cap));
    CU(sd->d_recs.ensure(cap));
    return 0;
}

// correlate blocks [b0, b0+nb), then decide peaks on [lo, hi)
int run_chunk(b200sync_sd* sd, const float2* d_in, long long in_base, float* d_z, long long z_base,
              long long b0, long long nb, long long lo, long long hi, float2* d_out_delayed,
              long long out_end, cudaStream_t st) {
    float2* gm = (d_z == sd->d_zoff.p && sd->gm_blocks > 0) ? sd->d_gm.p : nullptr;   // offline metric only
    CU(launch_correlate(d_in, in_base, d_z, z_base, sd->d_hperm.p, (int)sd->K, (int)sd->S, sd->fft_arg, b0, nb,
                        sd->d_tw.p, d_out_delayed, 0, 0, out_end, (int)sd->delay, sd->num_sms, st, 0, 0, 0, gm,
                        sd->gm_b0, 0));
    if (hi > lo) {
        const long long z_end = (b0 + nb) * (long long)sd->S;
        if (hi - lo <= kSmallRange || sd->T > kMaxTimeThreshold) {
            CU(launch_peak_stream(d_z, z_base, z_end, lo, hi, sd->T, sd->power_threshold, sd->d_ws.p, sd->d_ws.cap,
                                  sd->d_state.p, sd->d_det_idx.p, (unsigned)sd->d_det_idx.cap, sd->num_sms, st));
        } else {
            CU(launch_peak_phase1(d_z, z_base, z_end, lo, hi, sd->T, sd->power_threshold, sd->d_ws.p,
                                  sd->d_ws.cap, nullptr, sd->num_sms, st));
            CU(launch_peak_phase2(lo, hi, sd->T, sd->d_ws.p, sd->d_ws.cap, -1, sd->d_state.p,
                                  sd->d_det_idx.p, (unsigned)sd->d_det_idx.cap, sd->num_sms, st));
        }
    }
    return 0;
}


// Page-lock the pages of a pageable span (b200sync_sd_set_auto_register); true when the whole span lies inside ONE
// registration of this context afterwards.
bool auto_register_span(b200sync_sd* sd, const void* p, size_t bytes) {
    if (sd->auto_register < 0) {
        const char* v = getenv("B200SYNC_AUTO_REGISTER");
        sd->auto_register = (v && v[0] == '1') ? 1 : 0;
    }
    if (sd->auto_register != 1 || bytes == 0) return false;
    const uintptr_t page = 4096;
    uintptr_t lo = reinterpret_cast<uintptr_t>(p) & ~(page - 1);
    const uintptr_t hi = (reinterpret_cast<uintptr_t>(p) + bytes + page - 1) & ~(page - 1);
    for (const auto& f : sd->reg_failed)
        if (lo < f.second && f.first < hi) return false;
    auto& rr = sd->reg_ranges;
    // One registration must cover the whole span (a copy may not straddle two registrations), so every range of this
    // context that overlaps or touches [lo, hi) is released and the union registered as ONE range: a ring converges to
    // a single registration after one pass, a capture walked front to back to one that grows with it.
    uintptr_t ulo = lo, uhi = hi;
    std::vector<std::pair<uintptr_t, uintptr_t>> keep, merge;
    for (const auto& r : rr) {
        if (r.second < lo || r.first > hi) keep.push_back(r);
        else merge.push_back(r);
    }
    if (merge.size() == 1 && merge[0].first <= lo && merge[0].second >= hi) return true;   // already inside one range
    for (const auto& r : merge) {
        ulo = std::min(ulo, r.first);
        uhi = std::max(uhi, r.second);
        if (cudaHostUnregister(reinterpret_cast<void*>(r.first)) != cudaSuccess) cudaGetLastError();
    }
    rr = keep;
    if (cudaHostRegister(reinterpret_cast<void*>(ulo), uhi - ulo, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        sd->reg_failed.emplace_back(ulo, uhi);
        return false;
    }
    rr.emplace_back(ulo, uhi);
    std::sort(rr.begin(), rr.end());
    return true;
}

// H2D of a host span for the streaming path.  Pinned / registered memory goes straight to the copy engine; a
// pageable span of ordinary ring-chunk size is staged through the context's own pinned buffer piece by piece (the
// memcpy of piece i+1 overlaps the DMA of piece i) instead of through the driver's internal staging.
int stream_h2d(b200sync_sd* sd, float2* d_dst, const float2* h_src, size_t count, cudaStream_t st) {
    const size_t bytes = count * sizeof(float2);
    bool pinned = false;
    if (bytes <= kStageBytes) {
        if (auto_register_span(sd, h_src, bytes)) {
            pinned = true;   // inside ONE registration of this context
        } else {
            cudaPointerAttributes at{};
            if (cudaPointerGetAttributes(&at, h_src) == cudaSuccess) pinned = (at.type == cudaMemoryTypeHost);
            else cudaGetLastError();
            // the first byte may lie in a range this context registered while the span runs past its end (a later
            // registration failed): such a span takes the staged copy
            const uintptr_t a = reinterpret_cast<uintptr_t>(h_src), b = a + bytes;
            for (const auto& r : sd->reg_ranges)
                if (a < r.second && r.first < b && !(r.first <= a && b <= r.second)) pinned = false;
        }
    }
    if (bytes > kStageBytes) {
        CU(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
        return 0;
    }
    if (pinned) {
        // (a span that runs from page-locked into pageable memory, or across two registrations — e.g. a ring whose
        //  mirror mapping was registered separately — may be refused: such a span takes the staged copy below)
        if (cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st) == cudaSuccess) return 0;
        cudaGetLastError();
    }
    if (!sd->h_in_stage) CU(cudaMallocHost(&sd->h_in_stage, kStageBytes));
    const unsigned char* src = reinterpret_cast<const unsigned char*>(h_src);
    unsigned char* dst = reinterpret_cast<unsigned char*>(d_dst);
    for (size_t off = 0; off < bytes; off += kStagePiece) {
        const size_t len = std::min(kStagePiece, bytes - off);
        std::memcpy(sd->h_in_stage + off, src + off, len);   // every call ends with a stream sync: the buffer is free
        CU(cudaMemcpyAsync(dst + off, sd->h_in_stage + off, len, cudaMemcpyHostToDevice, st));
    }
    return 0;
}

// Zero-copy input of the streaming path (the correlator pulls the span out of mapped host memory itself instead of
// waiting for an H2D copy): MEASURED SLOWER on B200 — 89 vs 84 us per 65536-item pinned span; SM-issued PCIe reads of
// 512 KiB take longer than the copy engine's DMA plus its fixed latency — so it is off unless B200SYNC_ZERO_COPY=1.
bool zero_copy_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* v = getenv("B200SYNC_ZERO_COPY");
        on = (v && v[0] == '1') ? 1 : 0;
    }
    return on == 1;
}

// Device-visible address of a host span for the zero-copy correlator: pinned / registered-and-mapped memory as it is;
// a pageable span of ordinary ring-chunk size through the context's pinned staging buffer.  *view stays null when
// neither applies (the caller then takes the H2D copy).
int stream_host_view(b200sync_sd* sd, const float2* h_src, size_t count, const float2** view) {
    *view = nullptr;
    const size_t bytes = count * sizeof(float2);
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, h_src) == cudaSuccess) {
        if (at.type == cudaMemoryTypeHost && at.devicePointer != nullptr) {
            *view = static_cast<const float2*>(at.devicePointer);
            return 0;
        }
    } else {
        cudaGetLastError();
    }
    if (bytes > kStageBytes) return 0;
    if (!sd->h_in_stage) CU(cudaMallocHost(&sd->h_in_stage, kStageBytes));
    if (!sd->d_in_stage_view) CU(cudaHostGetDevicePointer(&sd->d_in_stage_view, sd->h_in_stage, 0));
    std::memcpy(sd->h_in_stage, h_src, bytes);   // the previous step has completed: the buffer is free
    *view = static_cast<const float2*>(sd->d_in_stage_view);
    return 0;
}

// refine + result copies, all asynchronous: at most `nmax` records can exist for the decided range, so the state and
// that many records are copied blind into pinned memory and ONE synchronisation (records_finish) suffices
int records_enqueue(b200sync_sd* sd, const float2* d_in, long long in_base, const float* d_z, long long z_base,
                    size_t nmax, cudaStream_t st) {
    nmax = std::min(nmax, sd->d_det_idx.cap);
    if (sd->h_recs_pin_cap < nmax) {
        if (sd->h_recs_pin) cudaFreeHost(sd->h_recs_pin);
        sd->h_recs_pin = nullptr;
        sd->h_recs_pin_cap = 0;
        CU(cudaMallocHost(&sd->h_recs_pin, sizeof(DetectionRecord) * (nmax + 64)));
        sd->h_recs_pin_cap = nmax + 64;
    }
    CU(launch_refine(d_in, in_base, d_z, z_base, sd->d_hperm.p, (int)sd->K, (int)sd->S, sd->fft_arg, sd->min_bin, sd->d_tw.p,
                     sd->d_det_idx.p, &sd->d_state.p->det_count, (unsigned)std::max<size_t>(nmax, 1), sd->d_recs.p,
                     sd->num_sms, st));
    CU(cudaMemcpyAsync(sd->h_state, sd->d_state.p, sizeof(PeakState), cudaMemcpyDeviceToHost, st));
    if (nmax > 0)
        CU(cudaMemcpyAsync(sd->h_recs_pin, sd->d_recs.p, sizeof(DetectionRecord) * nmax, cudaMemcpyDeviceToHost, st));
    CU(cudaMemsetAsync(&sd->d_state.p->det_count, 0, sizeof(unsigned int), st));  // the list has been drained
    return 0;
}
int records_finish(b200sync_sd* sd, size_t nmax, cudaStream_t st, const DetectionRecord** recs, size_t* n) {
    CU(cudaStreamSynchronize(st));
    const size_t cnt = sd->h_state->det_count;
    if (cnt > std::min(nmax, sd->d_det_idx.cap)) return fail(B200SYNC_ENOMEM, "internal detection list overflow");
    *recs = sd->h_recs_pin;
    *n = cnt;
    return 0;
}

// refine the device-side detection list (already sorted by index, peaks.cu det_gather_kernel)
// into host records.  The refine launch sizes itself from the device-side count, so the host
// synchronises once to learn the count and once for the records.
int collect_records(b200sync_sd* sd, const float2* d_in, long long in_base, const float* d_z,
                    long long z_base, cudaStream_t st, std::vector<DetectionRecord>& out) {
    CU(launch_refine(d_in, in_base, d_z, z_base, sd->d_hperm.p, (int)sd->K, (int)sd->S, sd->fft_arg, sd->min_bin,
                     sd->d_tw.p, sd->d_det_idx.p, &sd->d_state.p->det_count, (unsigned)sd->d_det_idx.cap,
                     sd->d_recs.p, sd->num_sms, st));
    CU(cudaMemcpyAsync(sd->h_state, sd->d_state.p, sizeof(PeakState), cudaMemcpyDeviceToHost, st));
    // the list has been drained
    CU(cudaMemsetAsync(&sd->d_state.p->det_count, 0, sizeof(unsigned int), st));
    CU(cudaStreamSynchronize(st));
    out.clear();
    const unsigned n = sd->h_state->det_count;
    if (n == 0) return 0;
    if (n > sd->d_det_idx.cap) return fail(B200SYNC_ENOMEM, "internal detection list overflow");
    // through the context's pinned landing buffer: a D2H copy into pageable memory (the vector) runs at a fraction of
    // the link rate and blocks in the driver (2 MB of records: 0.25 ms against 0.05)
    if (sd->h_recs_pin_cap < n) {
        if (sd->h_recs_pin) cudaFreeHost(sd->h_recs_pin);
        sd->h_recs_pin = nullptr;
        sd->h_recs_pin_cap = 0;
        CU(cudaMallocHost(&sd->h_recs_pin, sizeof(DetectionRecord) * (static_cast<size_t>(n) + n / 4 + 64)));
        sd->h_recs_pin_cap = static_cast<size_t>(n) + n / 4 + 64;
    }
    CU(cudaMemcpyAsync(sd->h_recs_pin, sd->d_recs.p, sizeof(DetectionRecord) * n, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    out.assign(sd->h_recs_pin, sd->h_recs_pin + n);
    return 0;
}

// SyncwordDetection::output_tag — PM/syncword_detection.hpp:56-115
b200sync_sd_tag make_tag(const b200sync_sd* sd, const DetectionRecord& r) {
    const double Ld = static_cast<double>(sd->L);
    const double bin_spacing = std::numbers::pi / Ld;
    double freq = static_cast<double>(r.freq_bin) * bin_spacing;
    float phase = std::arg(c64(r.corr_re, r.corr_im));
    float cpow;
    if (r.freq_bin > sd->min_bin && r.freq_bin < sd->max_bin) {
        const double a = r.pow_left, b = r.pow, c = r.pow_right;
        const double quad = std::clamp((c - a) / (2.0 * (2.0 * b - (a + c))), -0.5, 0.5);
        const double dfreq = quad * bin_spacing;
        freq += dfreq;
        phase -= static_cast<float>(dfreq * 0.5 * Ld);
        if (phase >= std::numbers::pi_v<float>) phase -= 2.0f * std::numbers::pi_v<float>;
        else if (phase < -std::numbers::pi_v<float>) phase += 2.0f * std::numbers::pi_v<float>;
        cpow = static_cast<float>(b + (c - a) * (c - a) / (16.0 * (b - 0.5 * (a + c))));
    } else {
        cpow = r.pow;
    }
    const float amp = std::sqrt(cpow) / (static_cast<float>(sd->fft_size) * sd->self_corr);
    const float spow = amp * amp * sd->self_corr;
    const float esn0 = 10.0f * std::log10((spow * static_cast<float>(sd->sps)) /
                                          (r.noise_power * static_cast<float>(sd->L)));
    const double a = r.pow_prev, b = r.pow, c = r.pow_next;
    const float time_est =
        static_cast<float>(std::clamp((c - a) / (2.0 * (2.0 * b - (a + c))), -0.5, 0.5));
    b200sync_sd_tag t{};
    t.index = r.index + sd->delay;
    t.syncword_freq = freq;
    t.syncword_amplitude = amp;
    t.syncword_phase = phase;
    t.syncword_freq_bin = r.freq_bin;
    t.syncword_noise_power = r.noise_power;
    t.syncword_esn0_db = esn0;
    t.syncword_time_est = time_est;
    return t;
}

// start(): PM/syncword_detection.hpp:143-202
int do_start(b200sync_sd* sd) {
    if (sd->min_bin > sd->max_bin) return fail(B200SYNC_EINVAL, "min_freq_bin is greater than max_freq_bin");
    if (sd->syncword.empty() || sd->rrc_taps.empty() || sd->constellation.empty() || sd->sps == 0)
        return fail(B200SYNC_EINVAL, "syncword, rrc_taps and constellation must be non-empty");
    for (uint8_t s : sd->syncword)
        if (s >= sd->constellation.size()) return fail(B200SYNC_EINVAL, "syncword symbol outside constellation");
    sd->L = static_cast<uint32_t>((sd->syncword.size() - 1) * sd->sps + sd->rrc_taps.size());
    if (sd->L > sd->fft_size) return fail(B200SYNC_EINVAL, "fft_size too small");
    if ((sd->fft_size & (sd->fft_size - 1)) != 0)
        return fail(B200SYNC_EINVAL, "Input data must have 2^N samples, input size: ");   // the text of ALG/fourier/fftw.hpp:182-184 (sic)
    {
        const char* fg = getenv("B200SYNC_FORCE_GENERIC");     // tests / A-B runs: fft_size 2048 on the generic path too
        sd->generic = sd->fft_size != (uint32_t)kFft || (fg && fg[0] == '1');
        sd->fft_arg = (int)sd->fft_size | (sd->generic && sd->fft_size == (uint32_t)kFft ? kGenericFlag : 0);
    }
    if (sd->generic && !generic_fft_supported(sd->fft_size))
        return fail(B200SYNC_EUNSUPPORTED, "fft_size outside [64, 8192] is not implemented on the GPU path");
    sd->K = static_cast<uint32_t>(sd->max_bin - sd->min_bin + 1);
    if (sd->K > (uint32_t)kMaxHyp) return fail(B200SYNC_EUNSUPPORTED, "too many frequency hypotheses (max 601)");
    if (sd->time_threshold > (uint64_t)kMaxTimeThresholdSeq)
        return fail(B200SYNC_EUNSUPPORTED, "time_threshold > 4095 is not implemented on the GPU path");
    if (!(sd->power_threshold > 0.0f)) return fail(B200SYNC_EINVAL, "power_threshold must be positive");
    sd->S = sd->fft_size - sd->L + 1;
    sd->T = static_cast<int>(sd->time_threshold);
    sd->delay = 2 * sd->time_threshold + 1;

    std::vector<c64> sw(sd->L);
    for (size_t j = 0; j < sd->syncword.size(); ++j)
        for (size_t k = 0; k < sd->rrc_taps.size(); ++k) {
            const c64 cst = sd->constellation[sd->syncword[j]];
            const float t = sd->rrc_taps[k];
            c64& d = sw[j * sd->sps + k];
            d = c64(d.real() + cst.real() * t, d.imag() + cst.imag() * t);
        }
    sd->self_corr = 0.0f;
    for (auto x : sw) sd->self_corr += x.real() * x.real() + x.imag() * x.imag();

    const size_t F = sd->fft_size;
    std::vector<c64> td(static_cast<size_t>(sd->K) * F, c64(0.0f, 0.0f));
    for (int bin = sd->min_bin; bin <= sd->max_bin; ++bin) {
        double phase = 0.0;
        const double incr = static_cast<double>(bin) * std::numbers::pi / static_cast<double>(sd->L);
        c64* dst = td.data() + static_cast<size_t>(bin - sd->min_bin) * F;
        for (uint32_t n = 0; n < sd->L; ++n) {
            const float er = static_cast<float>(std::cos(phase)), ei = static_cast<float>(std::sin(phase));
            const c64 x = sw[n];
            dst[n] = c64(x.real() * er - x.imag() * ei, x.real() * ei + x.imag() * er);
            phase += incr;
            if (phase >= std::numbers::pi) phase -= 2.0 * std::numbers::pi;
            else if (phase < std::numbers::pi) phase += 2.0 * std::numbers::pi;  // sic, :179-181
        }
    }
    // twiddle table Wt[j] = exp(-2 pi i j / 2048) in float (fft2048.cuh arithmetic contract),
    // re-ordered into the two conflict-free tables the kernels index
    std::vector<float2> wt(kFft);
    for (int j = 0; j < kFft; ++j) {
        const double a = 2.0 * std::numbers::pi * static_cast<double>(j) / static_cast<double>(kFft);
        wt[j] = make_float2(static_cast<float>(std::cos(a)), static_cast<float>(-std::sin(a)));
    }
    std::vector<float2> tw(kTwTotalHost);
    for (int m2 = 0; m2 < 16; ++m2)
        for (int f1 = 0; f1 < 16; ++f1) tw[m2 * 16 + f1] = wt[8 * f1 * m2];
    for (int m3 = 0; m3 < 8; ++m3)
        for (int p = 0; p < 256; ++p) tw[256 + m3 * 256 + p] = wt[p * m3];
    if (sd->generic) {
        // correlator_generic.cu: per-stage radix-2 factors, stage of length s at offset s/2 - 1:
        // (float cos a, float sin a), a = -2 pi j / s in double (the oracle's independent arithmetic)
        tw.assign(F, make_float2(0.0f, 0.0f));
        size_t o = 0;
        for (size_t s = 2; s <= F; s *= 2)
            for (size_t j = 0; j < s / 2; ++j) {
                const double a = -2.0 * std::numbers::pi * static_cast<double>(j) / static_cast<double>(s);
                tw[o++] = make_float2(static_cast<float>(std::cos(a)), static_cast<float>(std::sin(a)));
            }
    }
    CU(cudaSetDevice(sd->device));
    cudaDeviceProp prop{};
    CU(cudaGetDeviceProperties(&prop, sd->device));
    if (prop.major < 10)
        return fail(B200SYNC_EUNSUPPORTED, "libb200sync is built for sm_100a; no Blackwell device found");
    sd->num_sms = prop.multiProcessorCount;
    if (!sd->stream) CU(cudaStreamCreateWithFlags(&sd->stream, cudaStreamNonBlocking));
    for (auto& e : sd->ev)
        if (!e) CU(cudaEventCreate(&e));
    if (!sd->h_state) CU(cudaMallocHost(&sd->h_state, sizeof(PeakState)));
    CU(sd->d_tw.ensure(tw.size()));
    CU(sd->d_hperm.ensure(static_cast<size_t>(sd->K) * F));
    DevBuf<float2> d_td;
    CU(d_td.ensure(td.size()));
    CU(cudaMemcpyAsync(d_td.p, td.data(), td.size() * sizeof(float2), cudaMemcpyHostToDevice, sd->stream));
    CU(cudaMemcpyAsync(sd->d_tw.p, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, sd->stream));
    CU(launch_template_spectra(d_td.p, sd->d_hperm.p, (int)sd->K, sd->d_tw.p, sd->stream, sd->fft_arg));
    // streaming state (:191-201)
    if (int rc = reset_state(sd, sd->stream)) return rc;
    CU(cudaStreamSynchronize(sd->stream));
    sd->consumed = 0;
    sd->lo_next = 0;
    sd->r_abs_host = 0;
    sd->x_base = sd->x_end = 0;
    sd->z_base = sd->z_end = 0;
    sd->carry.assign(sd->delay, c64(0.0f, 0.0f));
    sd->pending.clear();
    sd->shard.valid = false;
    return 0;
}

constexpr long long kOfflineChunkBlocks = 8192;  // ~14.4 M samples per correlator launch when chasing H2D copies
constexpr long long kStreamStepBlocks = 4096;    // max blocks per streaming step

// slide a device window down so that it starts at absolute index `keep_from`
template <typename T>
int slide_window(DevBuf<T>& buf, DevBuf<T>& tmp, long long& base, long long end, long long keep_from,
                 cudaStream_t st) {
    if (keep_from <= base) return 0;
    if (keep_from > end) keep_from = end;
    const size_t n = static_cast<size_t>(end - keep_from);
    if (n > 0) {
        CU(tmp.ensure(n));
        CU(cudaMemcpyAsync(tmp.p, buf.p + (keep_from - base), n * sizeof(T), cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(buf.p, tmp.p, n * sizeof(T), cudaMemcpyDeviceToDevice, st));
    }
    base = keep_from;
    return 0;
}

}  // namespace

namespace b200sync {
void count_launch(int n) { g_launches += static_cast<uint64_t>(n); }
int set_last_error(int code, const std::string& msg) { return fail(code, msg); }
}  // namespace b200sync

extern "C" {

const char* b200sync_last_error(void) { return g_last_error.c_str(); }
int b200sync_abi_version(void) { return 1; }
uint64_t b200sync_launch_count(void) { return g_launches.load(); }

int b200sync_host_register(const void* ptr, size_t bytes) {
    if (!ptr || !bytes) return fail(B200SYNC_EINVAL, "null buffer");
    CU(cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return 0;
}
int b200sync_sd_set_auto_register(b200sync_sd* sd, int on) {
    if (!sd) return fail(B200SYNC_EINVAL, "null context");
    sd->auto_register = on ? 1 : 0;
    return 0;
}
int b200sync_host_unregister(const void* ptr) {
    if (!ptr) return fail(B200SYNC_EINVAL, "null buffer");
    CU(cudaHostUnregister(const_cast<void*>(ptr)));
    return 0;
}

int b200sync_sd_create(const b200sync_sd_config* cfg, b200sync_sd** out) {
    if (!cfg || !out) return fail(B200SYNC_EINVAL, "null argument");
    *out = nullptr;
    if (!cfg->rrc_taps || !cfg->syncword || !cfg->constellation)
        return fail(B200SYNC_EINVAL, "null settings array");
    b200sync_sd* sd = new (std::nothrow) b200sync_sd();
    if (!sd) return fail(B200SYNC_ENOMEM, "out of memory");
    sd->fft_size = cfg->fft_size ? cfg->fft_size : 2048;
    sd->sps = cfg->samples_per_symbol ? cfg->samples_per_symbol : 4;
    sd->rrc_taps.assign(cfg->rrc_taps, cfg->rrc_taps + cfg->n_rrc_taps);
    sd->syncword.assign(cfg->syncword, cfg->syncword + cfg->n_syncword);
    sd->constellation.resize(cfg->n_constellation);
    for (uint32_t i = 0; i < cfg->n_constellation; ++i)
        sd->constellation[i] = c64(cfg->constellation[2 * i], cfg->constellation[2 * i + 1]);
    sd->min_bin = cfg->min_freq_bin;
    sd->max_bin = cfg->max_freq_bin;
    sd->time_threshold = cfg->time_threshold;
    sd->power_threshold = cfg->power_threshold;
    sd->device = cfg->device;
    const int rc = do_start(sd);
    if (rc != 0) {
        const std::string keep = g_last_error;
        b200sync_sd_destroy(sd);
        g_last_error = keep;
        return rc;
    }
    *out = sd;
    return 0;
}

void b200sync_sd_destroy(b200sync_sd* sd) {
    if (!sd) return;
    cudaSetDevice(sd->device);
    if (sd->stream) {
        cudaStreamSynchronize(sd->stream);
        cudaStreamDestroy(sd->stream);
    }
    for (auto& e : sd->ev)
        if (e) cudaEventDestroy(e);
    for (const auto& r : sd->reg_ranges)
        if (cudaHostUnregister(reinterpret_cast<void*>(r.first)) != cudaSuccess) cudaGetLastError();
    if (sd->h_state) cudaFreeHost(sd->h_state);
    if (sd->h_in_stage) cudaFreeHost(sd->h_in_stage);
    if (sd->h_recs_pin) cudaFreeHost(sd->h_recs_pin);
    if (sd->h_chan_state) cudaFreeHost(sd->h_chan_state);
    if (sd->h_stage) cudaFreeHost(sd->h_stage);
    for (auto& e : sd->ev_stage)
        if (e) cudaEventDestroy(e);
    for (auto& e : sd->ev_pieces) cudaEventDestroy(e);
    for (auto& e : sd->ev_out) cudaEventDestroy(e);
    if (sd->d2h_stream) cudaStreamDestroy(sd->d2h_stream);
    if (sd->copy_stream) cudaStreamDestroy(sd->copy_stream);
    delete sd;
}

int b200sync_sd_start(b200sync_sd* sd) {
    if (!sd) return fail(B200SYNC_EINVAL, "null context");
    return do_start(sd);
}

int b200sync_sd_info(const b200sync_sd* sd, uint32_t* syncword_samples, uint32_t* stride, float* self_corr,
                     uint32_t* num_hypotheses, uint64_t* delay) {
    if (!sd) return fail(B200SYNC_EINVAL, "null context");
    if (syncword_samples) *syncword_samples = sd->L;
    if (stride) *stride = sd->S;
    if (self_corr) *self_corr = sd->self_corr;
    if (num_hypotheses) *num_hypotheses = sd->K;
    if (delay) *delay = sd->delay;
    return 0;
}

int b200sync_sd_records_to_tags(const b200sync_sd* sd, const b200sync_detection_record* recs, size_t n,
                                b200sync_sd_tag* tags) {
    if (!sd || (!recs && n) || (!tags && n)) return fail(B200SYNC_EINVAL, "null argument");
    // output_tag() is per-record host arithmetic (atan2f, log10f, double divides): spread large
    // batches over the host cores
    const size_t kPerThread = 4096;
    size_t nthreads = std::min<size_t>((n + kPerThread - 1) / kPerThread,
                                       std::max(1u, std::thread::hardware_concurrency()));
    if (nthreads <= 1) {
        for (size_t i = 0; i < n; ++i) tags[i] = make_tag(sd, recs[i]);
        return 0;
    }
    std::vector<std::thread> pool;
    pool.reserve(nthreads);
    for (size_t t = 0; t < nthreads; ++t) {
        const size_t i0 = n * t / nthreads, i1 = n * (t + 1) / nthreads;
        pool.emplace_back([=] { for (size_t i = i0; i < i1; ++i) tags[i] = make_tag(sd, recs[i]); });
    }
    for (auto& th : pool) th.join();
    return 0;
}

// Development aid (B200SYNC_TRACE=1): where the host time of a streaming call goes; printed to stderr every 256 calls
namespace {
struct StreamTrace {
    bool on = false, init = false, dev = false;   // dev (B200SYNC_TRACE=2): device-side stage times too (events between the launches)
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    double dacc[4] = {0, 0, 0, 0};
    double acc[6] = {0, 0, 0, 0, 0, 0};
    long calls = 0;
    static double now() {
        return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }
    void lap(int i, double& t) {
        if (!on) return;
        const double n = now();
        acc[i] += n - t;
        t = n;
    }
    void end() {
        if (!on || ++calls % 256) return;
        std::fprintf(stderr, "[b200sync trace] per call, us: setup %.1f | h2d enqueue (+staging) %.1f | kernels + d2h enqueue %.1f | "
                             "host delay line %.1f | wait %.1f | tags %.1f\n",
                     acc[0] / 256, acc[1] / 256, acc[2] / 256, acc[3] / 256, acc[4] / 256, acc[5] / 256);
        for (double& a : acc) a = 0;
        if (dev) {
            std::fprintf(stderr, "[b200sync trace] device, us: h2d %.1f | correlate %.1f | flags %.1f | walk+refine %.1f\n",
                         dacc[0] / 256, dacc[1] / 256, dacc[2] / 256, dacc[3] / 256);
            for (double& a : dacc) a = 0;
        }
    }
    void mark(int i, cudaStream_t st) {
        if (!dev) return;
        if (!ev[i]) cudaEventCreate(&ev[i]);
        cudaEventRecord(ev[i], st);
    }
    void collect() {
        if (!dev) return;
        for (int i = 0; i < 4; ++i) {
            float ms = 0.f;
            if (ev[i] && ev[i + 1] && cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) == cudaSuccess) dacc[i] += 1e3 * ms;
        }
    }
};
thread_local StreamTrace g_trace;
}  // namespace

// delay line of the streaming block (:318-319): out[i] = stream[C + i - delay]; then remember the last `delay` items
static void host_delay_line_stream(b200sync_sd* sd, const float* in, float* out, size_t j) {
    const size_t D = static_cast<size_t>(sd->delay);
    const c64* cin = reinterpret_cast<const c64*>(in);
    if (out) {
        c64* cout = reinterpret_cast<c64*>(out);
        const size_t from_carry = std::min(D, j);
        std::memcpy(cout, sd->carry.data(), from_carry * sizeof(c64));
        if (j > D) {
            // large spans (a flowgraph edge with a 2^18 .. 2^20-item buffer): the copy is spread over a few threads,
            // one memcpy would take as long as everything the GPU does for the span
            const size_t cnt = j - D;
            const size_t nth = std::min<size_t>({4, std::max(1u, std::thread::hardware_concurrency() / 2), cnt >> 17});
            if (nth <= 1) {
                std::memcpy(cout + D, cin, cnt * sizeof(c64));
            } else {
                std::vector<std::thread> pool;
                for (size_t t = 1; t < nth; ++t) {
                    const size_t i0 = cnt * t / nth, i1 = cnt * (t + 1) / nth;
                    pool.emplace_back([=] { std::memcpy(cout + D + i0, cin + i0, (i1 - i0) * sizeof(c64)); });
                }
                std::memcpy(cout + D, cin, (cnt / nth) * sizeof(c64));
                for (auto& th : pool) th.join();
            }
        }
    }
    if (j >= D) {
        std::memcpy(sd->carry.data(), cin + (j - D), D * sizeof(c64));
    } else {
        std::memmove(sd->carry.data(), sd->carry.data() + j, (D - j) * sizeof(c64));
        std::memcpy(sd->carry.data() + (D - j), cin, j * sizeof(c64));
    }
}

// ------------------------------------------------------------------------------------------
int b200sync_sd_process(b200sync_sd* sd, const float* in, size_t n_in, float* out, size_t* n_consumed,
                        b200sync_sd_tag* tags, size_t max_tags, size_t* n_tags) {
    if (!sd || !n_consumed || !n_tags || (!in && n_in)) return fail(B200SYNC_EINVAL, "null argument");
    *n_consumed = 0;
    *n_tags = 0;
    if (n_in < sd->fft_size) return 1;  // INSUFFICIENT_INPUT_ITEMS, consume(0)/publish(0) (:215-227)
    if (!g_trace.init) {
        const char* v = getenv("B200SYNC_TRACE");
        g_trace.on = v && (v[0] == '1' || v[0] == '2');
        g_trace.dev = v && v[0] == '2';
        g_trace.init = true;
    }
    double tt = g_trace.on ? StreamTrace::now() : 0.0;
    CU(cudaSetDevice(sd->device));
    cudaStream_t st = sd->stream;
    const long long S = sd->S, F = sd->fft_size, T = sd->T;
    const long long nb_total = (static_cast<long long>(n_in) - F) / S + 1;
    const long long j_total = nb_total * S;
    const long long C = static_cast<long long>(sd->consumed);
    const float2* hin = reinterpret_cast<const float2*>(in);

    long long done_blocks = 0;
    while (done_blocks < nb_total) {
        const long long nb = std::min(kStreamStepBlocks, nb_total - done_blocks);
        const long long a0 = C + done_blocks * S;            // absolute first sample of this step
        const long long need_end = a0 + (nb - 1) * S + F;    // absolute end of samples needed
        const long long P_prev = a0, P = a0 + nb * S;
        // --- input window: keep every block that may still hold an undecided peak
        long long keep_x = (sd->lo_next / S) * S;
        if (keep_x > a0) keep_x = a0;
        const size_t x_need = static_cast<size_t>(need_end - keep_x);
        if (sd->d_x.cap < x_need) {
            DevBuf<float2> nx;
            CU(nx.ensure(x_need + static_cast<size_t>(kStreamStepBlocks * S)));
            if (sd->x_end > keep_x && sd->d_x.p)
                CU(cudaMemcpyAsync(nx.p, sd->d_x.p + (keep_x - sd->x_base),
                                   sizeof(float2) * static_cast<size_t>(sd->x_end - keep_x),
                                   cudaMemcpyDeviceToDevice, st));
            CU(cudaStreamSynchronize(st));
            std::swap(sd->d_x.p, nx.p);
            std::swap(sd->d_x.cap, nx.cap);
            sd->x_base = keep_x;
            if (sd->x_end < keep_x) sd->x_end = keep_x;
        } else if (static_cast<size_t>(need_end - sd->x_base) > sd->d_x.cap) {
            if (int rc = slide_window(sd->d_x, sd->d_xtmp, sd->x_base, sd->x_end, keep_x, st)) return rc;
            if (sd->x_end < keep_x) sd->x_end = keep_x;
        }
        g_trace.lap(0, tt);
        g_trace.mark(0, st);
        // Optional zero copy for a streaming-sized step (off by default, see zero_copy_enabled()): the correlator
        // pulls the span out of (pinned, mapped) host memory itself and fills the device window on the way; a
        // pageable span is first copied into the context's pinned staging buffer.
        const float2* host_view = nullptr;
        {
            long long hi_s = P - T - 1;
            if (hi_s < sd->lo_next) hi_s = sd->lo_next;
            const size_t nmax_s = static_cast<size_t>((hi_s - sd->lo_next) / (T + 1) + 2);
            const bool small_s = !sd->generic && hi_s > sd->lo_next && hi_s - sd->lo_next <= kSmallRange && nmax_s <= 1024;
            if (small_s && zero_copy_enabled() && correlate_takes_host_input((int)sd->K, sd->fft_arg, nb, sd->num_sms))
                if (int rc = stream_host_view(sd, hin + (a0 - C), static_cast<size_t>(need_end - a0), &host_view)) return rc;
        }
        if (host_view == nullptr)
            if (int rc = stream_h2d(sd, sd->d_x.p + (a0 - sd->x_base), hin + (a0 - C), static_cast<size_t>(need_end - a0), st))
                return rc;
        g_trace.lap(1, tt);
        g_trace.mark(1, st);
        sd->x_end = need_end;
        // --- metric window: [P_prev - 2T - 2, P)
        long long keep_z = P_prev - 2 * T - 2;
        if (keep_z < 0) keep_z = 0;
        const size_t z_need = static_cast<size_t>(P - keep_z);
        if (sd->d_z.cap < z_need) {
            DevBuf<float> nz;
            CU(nz.ensure(z_need + static_cast<size_t>(kStreamStepBlocks * S)));
            if (sd->z_end > keep_z && sd->d_z.p)
                CU(cudaMemcpyAsync(nz.p, sd->d_z.p + (keep_z - sd->z_base),
                                   sizeof(float) * static_cast<size_t>(sd->z_end - keep_z),
                                   cudaMemcpyDeviceToDevice, st));
            CU(cudaStreamSynchronize(st));
            std::swap(sd->d_z.p, nz.p);
            std::swap(sd->d_z.cap, nz.cap);
            sd->z_base = keep_z;
        } else if (static_cast<size_t>(P - sd->z_base) > sd->d_z.cap) {
            if (int rc = slide_window(sd->d_z, sd->d_ztmp, sd->z_base, sd->z_end, keep_z, st)) return rc;
        }
        // --- peaks decidable after this step: [lo_next, P - T - 1)
        const long long lo = sd->lo_next;
        long long hi = P - T - 1;
        if (hi < lo) hi = lo;
        const size_t ws = peak_workspace_bytes_sms(kStreamStepBlocks * S + T + 2, sd->T, sd->num_sms);
        CU(sd->d_ws.ensure(ws));
        if (int rc = ensure_det(sd, static_cast<size_t>((kStreamStepBlocks * S + T + 2) / (T + 1) + 2))) return rc;
        const size_t nmax = static_cast<size_t>((hi - lo) / (T + 1) + 2);
        // (the fused walk keeps the detection list in shared memory: at most 1024 entries, i.e. not for tiny T)
        const bool small = !sd->generic && hi > lo && hi - lo <= kSmallRange && nmax <= 1024 && nmax + 1 <= sd->d_recs.cap;
        if (small) {
            // streaming-sized step: correlator, flags, [walk + refine] — three launches, one D2H, one synchronisation
            CU(launch_correlate(sd->d_x.p, sd->x_base, sd->d_z.p, sd->z_base, sd->d_hperm.p, (int)sd->K, (int)sd->S, sd->fft_arg,
                                a0 / S, nb, sd->d_tw.p, nullptr, 0, 0, 0, (int)sd->delay, sd->num_sms, st, 0, 0, 0, nullptr, 0,
                                0, host_view, a0));
            g_trace.mark(2, st);
            StreamWalk walk;
            CU(launch_peak_flags_stream(sd->d_z.p, sd->z_base, P, lo, hi, sd->T, sd->power_threshold, sd->d_ws.p,
                                        sd->d_ws.cap, sd->num_sms, st, &walk));
            g_trace.mark(3, st);
            walk.r_abs_in = sd->r_abs_host;
            walk.state_out = sd->d_state.p;
            // The records land in MAPPED pinned host memory, written by the kernel itself: slot 0 is the header
            // (search position, count, and — stored last, after every CTA's record is visible — this call's
            // sequence number, which the host polls).  No D2H copy, no stream synchronisation.
            if (sd->h_recs_pin_cap < nmax + 1) {
                CU(cudaStreamSynchronize(st));
                if (sd->h_recs_pin) cudaFreeHost(sd->h_recs_pin);
                sd->h_recs_pin = nullptr;
                sd->h_recs_pin_cap = 0;
                CU(cudaMallocHost(&sd->h_recs_pin, sizeof(DetectionRecord) * (nmax + 65)));
                sd->h_recs_pin_cap = nmax + 65;
                std::memset(sd->h_recs_pin, 0, sizeof(DetectionRecord));
            }
            if (!sd->d_done.p) {
                CU(sd->d_done.ensure(1));
                CU(cudaMemsetAsync(sd->d_done.p, 0, sizeof(unsigned int), st));
            }
            if (++sd->stream_seq == 0) sd->stream_seq = 1;
            walk.header = reinterpret_cast<PeakState*>(sd->h_recs_pin);
            walk.header->_pad = 0;      // (the previous step has completed: nothing else writes here now)
            walk.done = sd->d_done.p;
            walk.seq = sd->stream_seq;
            CU(launch_refine(sd->d_x.p, sd->x_base, sd->d_z.p, sd->z_base, sd->d_hperm.p, (int)sd->K, (int)sd->S, sd->fft_arg,
                             sd->min_bin, sd->d_tw.p, sd->d_det_idx.p, &sd->d_state.p->det_count, (unsigned)nmax,
                             sd->h_recs_pin + 1, sd->num_sms, st, 1, 0, 0, 0, &walk));
            g_trace.mark(4, st);
        } else {
            if (host_view != nullptr) return fail(B200SYNC_ECUDA, "internal: zero-copy span on the bulk path");
            if (int rc = run_chunk(sd, sd->d_x.p, sd->x_base, sd->d_z.p, sd->z_base, a0 / S, nb, lo, hi, nullptr, 0, st))
                return rc;
            if (int rc = records_enqueue(sd, sd->d_x.p, sd->x_base, sd->d_z.p, sd->z_base, nmax, st)) return rc;
        }
        sd->z_end = P;
        sd->lo_next = hi;
        done_blocks += nb;
        // the delay line of host spans (:318-319) is host work: do it while the GPU runs the last step
        g_trace.lap(2, tt);
        if (done_blocks >= nb_total) host_delay_line_stream(sd, in, out, static_cast<size_t>(j_total));
        g_trace.lap(3, tt);
        const DetectionRecord* recs = nullptr;
        size_t nrec = 0;
        if (small) {
            {   // poll the sequence number; the stream is queried now and then so that a failed launch surfaces
                const volatile unsigned int* flag = &reinterpret_cast<const volatile PeakState*>(sd->h_recs_pin)->_pad;
                for (unsigned spins = 0; *flag != sd->stream_seq; ++spins) {
                    if ((spins & 0xfff) == 0xfff) {
                        const cudaError_t q = cudaStreamQuery(st);
                        if (q == cudaSuccess) {
                            if (*flag != sd->stream_seq) return fail(B200SYNC_ECUDA, "streaming step finished without its completion flag");
                            break;
                        }
                        if (q != cudaErrorNotReady) CU(q);
                    }
                    __builtin_ia32_pause();
                }
                std::atomic_thread_fence(std::memory_order_acquire);
            }
            g_trace.lap(4, tt);
            if (g_trace.dev) {
                cudaStreamSynchronize(st);
                g_trace.collect();
            }
            const PeakState* hs = reinterpret_cast<const PeakState*>(sd->h_recs_pin);
            if (hs->det_count > nmax) return fail(B200SYNC_ENOMEM, "internal detection list overflow");
            sd->r_abs_host = hs->r_abs;
            recs = sd->h_recs_pin + 1;
            nrec = hs->det_count;
        } else {
            if (int rc = records_finish(sd, nmax, st, &recs, &nrec)) return rc;
            if (hi > lo) sd->r_abs_host = sd->h_state->r_abs;
        }
        for (size_t i = 0; i < nrec; ++i) sd->pending.push_back(make_tag(sd, recs[i]));
    }
    const size_t j = static_cast<size_t>(j_total);
    sd->consumed += j;
    *n_consumed = j;
    // tags whose output index has now been published (:320-325)
    // The call has committed (state, carry line, consumed count): it can no longer fail.  Tags that do not
    // fit stay queued; b200sync_sd_tags_ready() reports them and b200sync_sd_drain_tags() hands them out.
    size_t nt = 0;
    while (nt < max_tags && !sd->pending.empty() && sd->pending.front().index < sd->consumed) {
        tags[nt++] = sd->pending.front();
        sd->pending.pop_front();
    }
    *n_tags = nt;
    g_trace.lap(5, tt);
    g_trace.end();
    return 0;
}

size_t b200sync_sd_tags_ready(const b200sync_sd* sd) {
    if (!sd) return 0;
    size_t n = 0;
    for (const auto& t : sd->pending) {
        if (t.index >= sd->consumed) break;  // the queue is sorted by index
        ++n;
    }
    return n;
}

int b200sync_sd_drain_tags(b200sync_sd* sd, b200sync_sd_tag* tags, size_t max_tags, size_t* n_tags) {
    if (!sd || !n_tags || (!tags && max_tags)) return fail(B200SYNC_EINVAL, "null argument");
    size_t nt = 0;
    while (nt < max_tags && !sd->pending.empty() && sd->pending.front().index < sd->consumed) {
        tags[nt++] = sd->pending.front();
        sd->pending.pop_front();
    }
    *n_tags = nt;
    return 0;
}

// ------------------------------------------------------------------------------------------
// offline detection = detect_begin + correlator chunks over [0, nb_total) + detect_finish
static int detect_begin(b200sync_sd* sd, size_t n, float2* d_out_delayed, cudaStream_t st, long long* nb_total,
                        long long* P) {
    const long long S = sd->S, F = sd->fft_size, T = sd->T;
    *nb_total = (static_cast<long long>(n) - F) / S + 1;
    *P = *nb_total * S;
    CU(sd->d_zoff.ensure(static_cast<size_t>(*P) + 64));
    sd->gm_b0 = 0;
    sd->gm_blocks = 0;
    if (gm_supported((int)sd->S, sd->T)) {
        CU(sd->d_gm.ensure(static_cast<size_t>(*nb_total) * gm_groups_per_block((int)sd->S) * kGmF2PerGroup));
        sd->gm_blocks = *nb_total;
    }
    const long long hi_total = std::max(0LL, *P - T - 1);
    CU(sd->d_ws.ensure(peak_workspace_bytes_sms(hi_total + 1, sd->T, sd->num_sms)));
    if (int rc = ensure_det(sd, static_cast<size_t>(*P / (T + 1) + 2))) return rc;
    if (int rc = reset_state(sd, st)) return rc;
    if (d_out_delayed) {
        const size_t zeros = std::min<size_t>(sd->delay, n);
        CU(cudaMemsetAsync(d_out_delayed, 0, zeros * sizeof(float2), st));
    }
    sd->ev_valid = false;
    CU(cudaEventRecord(sd->ev[0], st));
    return 0;
}

// the peak stage runs ONCE over the whole decided range (its sequential table scan costs per launch,
// not per sample), then refine + records to the host
static int detect_finish(b200sync_sd* sd, const float2* d_in, long long P, cudaStream_t st,
                         b200sync_detection_record* recs, size_t max_recs, size_t* n_recs, size_t* n_consumed) {
    const long long T = sd->T;
    const long long hi_total = std::max(0LL, P - T - 1);
    CU(cudaEventRecord(sd->ev[1], st));
    if (hi_total > 0 && sd->T > kMaxTimeThreshold) {
        // beyond the parallel chain kernels' range: generic flags + the in-order walk over the whole capture
        CU(launch_peak_stream(sd->d_zoff.p, 0, P, 0, hi_total, sd->T, sd->power_threshold, sd->d_ws.p, sd->d_ws.cap,
                              sd->d_state.p, sd->d_det_idx.p, (unsigned)sd->d_det_idx.cap, sd->num_sms, st));
    } else if (hi_total > 0) {
        CU(launch_peak_phase1(sd->d_zoff.p, 0, P, 0, hi_total, sd->T, sd->power_threshold, sd->d_ws.p,
                              sd->d_ws.cap, nullptr, sd->num_sms, st, 1, 0, 0, sd->gm_blocks > 0 ? sd->d_gm.p : nullptr,
                              sd->gm_b0, sd->gm_blocks, (int)sd->S, 0));
        CU(launch_peak_phase2(0, hi_total, sd->T, sd->d_ws.p, sd->d_ws.cap, -1, sd->d_state.p,
                              sd->d_det_idx.p, (unsigned)sd->d_det_idx.cap, sd->num_sms, st));
    }
    CU(cudaEventRecord(sd->ev[2], st));
    sd->metric_n = static_cast<size_t>(P);
    sd->metric_ptr = sd->d_zoff.p;
    sd->metric_base = 0;
    if (int rc = collect_records(sd, d_in, 0, sd->d_zoff.p, 0, st, sd->h_recs)) return rc;
    CU(cudaEventRecord(sd->ev[3], st));
    CU(cudaEventSynchronize(sd->ev[3]));
    sd->ev_valid = true;
    // only what the reference block would have tagged: output index < items published
    size_t cnt = 0;
    for (const auto& r : sd->h_recs) {
        if (r.index + sd->delay >= static_cast<uint64_t>(P)) continue;
        if (cnt >= max_recs) return fail(B200SYNC_ENOMEM, "record buffer too small");
        recs[cnt++] = r;
    }
    *n_recs = cnt;
    *n_consumed = static_cast<size_t>(P);
    return 0;
}

// How the delayed output span of a HOST-span bulk call is produced when the span is pinned.  Measured on this pool
// (profiles/r2_e2e_output_path.md): one GPU alone is fastest with a host-side copy on a few threads (5.8 vs 5.0 Gsps:
// a concurrent D2H stream slows the H2D stream from 53 to 40 GB/s), while several GPUs of one box saturate the
// host's memory system, where the copy costs 16 B/sample of host traffic and a D2H write only 8.  Default: host
// copy for the single-GPU bulk call, D2H for time shards.  B200SYNC_HOST_OUTPUT=d2h|memcpy overrides both.
static bool host_output_by_d2h(bool sharded) {
    static int forced = -1;
    if (forced < 0) {
        const char* v = std::getenv("B200SYNC_HOST_OUTPUT");
        forced = !v ? 0 : (std::strcmp(v, "d2h") == 0 ? 1 : (std::strcmp(v, "memcpy") == 0 ? 2 : 0));
    }
    return forced == 1 || (forced == 0 && sharded);
}

static bool is_pinned_host(const void* p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

// h_out_pinned != nullptr: the delayed output goes to d_out_delayed (device) AND is sent back to that pinned host
// span piece by piece on the context's D2H stream, behind the correlator chunk that completes each piece
static int detect_resident(b200sync_sd* sd, const float2* d_in, size_t n, float2* d_out_delayed,
                           cudaStream_t st, cudaEvent_t* chunk_ready, long long chunk_samples,
                           b200sync_detection_record* recs, size_t max_recs, size_t* n_recs,
                           size_t* n_consumed, float2* h_out_pinned = nullptr) {
    const long long S = sd->S, F = sd->fft_size;
    *n_recs = 0;
    *n_consumed = 0;
    if (n < static_cast<size_t>(F)) return 0;
    long long nb_total = 0, P = 0;
    if (int rc = detect_begin(sd, n, d_out_delayed, st, &nb_total, &P)) return rc;
    if (h_out_pinned && !sd->d2h_stream) CU(cudaStreamCreateWithFlags(&sd->d2h_stream, cudaStreamNonBlocking));
    long long out_done = 0;  // output items already on their way to the host
    size_t piece_i = 0;
    // correlator in chunks (so it can chase H2D copies)
    for (long long b0 = 0; b0 < nb_total; b0 += kOfflineChunkBlocks) {
        const long long nb = chunk_ready ? std::min(kOfflineChunkBlocks, nb_total - b0) : nb_total;
        if (chunk_ready) {
            // wait until the H2D copy covering the last sample of this chunk has landed
            const long long last_sample = (b0 + nb - 1) * S + F - 1;
            CU(cudaStreamWaitEvent(st, chunk_ready[last_sample / chunk_samples], 0));
        }
        if (int rc = run_chunk(sd, d_in, 0, sd->d_zoff.p, 0, b0, nb, 0, 0, d_out_delayed, P, st)) return rc;
        if (h_out_pinned) {
            // blocks [0, b0+nb) have written output items [0, (b0+nb)*S + delay) (zeros first, detect_begin)
            const long long upto = std::min(P, (b0 + nb) * S + static_cast<long long>(sd->delay));
            if (upto > out_done) {
                if (sd->ev_out.size() <= piece_i) {
                    cudaEvent_t e;
                    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                    sd->ev_out.push_back(e);
                }
                CU(cudaEventRecord(sd->ev_out[piece_i], st));
                CU(cudaStreamWaitEvent(sd->d2h_stream, sd->ev_out[piece_i], 0));
                CU(cudaMemcpyAsync(h_out_pinned + out_done, d_out_delayed + out_done,
                                   static_cast<size_t>(upto - out_done) * sizeof(float2), cudaMemcpyDeviceToHost,
                                   sd->d2h_stream));
                out_done = upto;
                ++piece_i;
            }
        }
        if (!chunk_ready) break;
    }
    const int rc = detect_finish(sd, d_in, P, st, recs, max_recs, n_recs, n_consumed);
    if (h_out_pinned) cudaStreamSynchronize(sd->d2h_stream);
    return rc;
}

int b200sync_sd_detect_device(b200sync_sd* sd, const void* d_in, size_t n, void* d_out_delayed,
                              void* cuda_stream, b200sync_detection_record* recs, size_t max_recs,
                              size_t* n_recs, size_t* n_consumed) {
    if (!sd || !n_recs || !n_consumed || (!d_in && n) || (!recs && max_recs))
        return fail(B200SYNC_EINVAL, "null argument");
    CU(cudaSetDevice(sd->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    return detect_resident(sd, static_cast<const float2*>(d_in), n, static_cast<float2*>(d_out_delayed), st,
                           nullptr, 0, recs, max_recs, n_recs, n_consumed);
}

// the delay line of the block contract for HOST spans (:318-319): out[i] = in[i - delay], zeros first.  The samples
// never needed the GPU: a plain copy, spread over a few host threads so that it keeps up with the PCIe link.
static void host_delay_line(const b200sync_sd* sd, const float* in, float* out, size_t P, std::vector<std::thread>& pool) {
    const size_t D = static_cast<size_t>(sd->delay);
    const size_t zeros = std::min(D, P);
    std::memset(out, 0, zeros * sizeof(c64));
    if (P <= D) return;
    const size_t cnt = P - D;
    const size_t nth = std::max<size_t>(1, std::min<size_t>({8, std::thread::hardware_concurrency() / 2, cnt >> 20}));
    for (size_t t = 0; t < nth; ++t) {
        const size_t i0 = cnt * t / nth, i1 = cnt * (t + 1) / nth;
        pool.emplace_back([=] { std::memcpy(out + 2 * (D + i0), in + 2 * i0, (i1 - i0) * sizeof(c64)); });
    }
}

static int detect_host_core(b200sync_sd* sd, const float* in, size_t n, float2* h_out_pinned,
                            b200sync_detection_record* recs, size_t max_recs, size_t* n_recs, size_t* n_consumed);

int b200sync_sd_detect_host_out(b200sync_sd* sd, const float* in, size_t n, float* out, b200sync_detection_record* recs,
                                size_t max_recs, size_t* n_recs, size_t* n_consumed) {
    if (!sd || !n_recs || !n_consumed || (!in && n) || (!recs && max_recs))
        return fail(B200SYNC_EINVAL, "null argument");
    if (out != nullptr && n >= sd->fft_size && host_output_by_d2h(false) && is_pinned_host(out)) {
        // pinned output span: the GPU writes the delayed stream and sends it back over the other PCIe direction
        CU(cudaSetDevice(sd->device));
        return detect_host_core(sd, in, n, reinterpret_cast<float2*>(out), recs, max_recs, n_recs, n_consumed);
    }
    std::vector<std::thread> pool;
    if (out != nullptr && n >= sd->fft_size) {
        const size_t P = ((n - sd->fft_size) / sd->S + 1) * sd->S;  // items the call will publish (:346)
        host_delay_line(sd, in, out, P, pool);                      // pageable span: host threads copy it
    }
    const int rc = b200sync_sd_detect_host(sd, in, n, recs, max_recs, n_recs, n_consumed);
    for (auto& t : pool) t.join();
    return rc;
}

int b200sync_sd_detect_host(b200sync_sd* sd, const float* in, size_t n, b200sync_detection_record* recs,
                            size_t max_recs, size_t* n_recs, size_t* n_consumed) {
    if (!sd || !n_recs || !n_consumed || (!in && n) || (!recs && max_recs))
        return fail(B200SYNC_EINVAL, "null argument");
    CU(cudaSetDevice(sd->device));
    return detect_host_core(sd, in, n, nullptr, recs, max_recs, n_recs, n_consumed);
}

static int detect_host_core(b200sync_sd* sd, const float* in, size_t n, float2* h_out_pinned,
                            b200sync_detection_record* recs, size_t max_recs, size_t* n_recs, size_t* n_consumed) {
    *n_recs = 0;
    *n_consumed = 0;
    if (n < sd->fft_size) return 0;
    CU(sd->d_xoff.ensure(n));
    if (h_out_pinned) CU(sd->d_outoff.ensure(n));
    // H2D on the context's copy stream in pieces, one (cached) event per piece; compute chases the copies
    const long long piece = 4LL << 20;  // 4 Mi samples = 32 MiB per copy
    const size_t npieces = (n + piece - 1) / piece;
    if (!sd->copy_stream) CU(cudaStreamCreateWithFlags(&sd->copy_stream, cudaStreamNonBlocking));
    cudaStream_t cs = sd->copy_stream;
    while (sd->ev_pieces.size() < npieces) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        sd->ev_pieces.push_back(e);
    }
    int rc = 0;
    for (size_t i = 0; i < npieces; ++i) {
        const size_t off = i * piece;
        const size_t cnt = std::min<size_t>(piece, n - off);
        cudaError_t e = cudaMemcpyAsync(sd->d_xoff.p + off, reinterpret_cast<const float2*>(in) + off,
                                        cnt * sizeof(float2), cudaMemcpyHostToDevice, cs);
        if (e == cudaSuccess) e = cudaEventRecord(sd->ev_pieces[i], cs);
        if (e != cudaSuccess) {
            rc = fail(B200SYNC_ECUDA, std::string("H2D copy: ") + cudaGetErrorString(e));
            break;
        }
    }
    if (rc == 0)
        rc = detect_resident(sd, sd->d_xoff.p, n, h_out_pinned ? sd->d_outoff.p : nullptr, sd->stream,
                             sd->ev_pieces.data(), piece, recs, max_recs, n_recs, n_consumed, h_out_pinned);
    cudaStreamSynchronize(cs);
    cudaStreamSynchronize(sd->stream);
    return rc;
}

// ------------------------------------------------------------------------------------------
// Raw capture ingestion: the on-disk format of FileSource<std::complex<float>>
// (PM/file_source.hpp:47-53: fread of sizeof(T)-byte items, no header; apps/README.md:15-19 "raw complex64").
// The file is read piece by piece into a small ring of pinned staging buffers; each piece goes to the
// device on a copy stream and the correlator chunks it completes are enqueued behind its event, so disk
// reads, PCIe copies and kernels overlap.  The capture ends up resident in HBM (8 B/sample).
int b200sync_sd_detect_file(b200sync_sd* sd, const char* filename, uint64_t first_item, uint64_t max_items,
                            b200sync_detection_record* recs, size_t max_recs, size_t* n_recs,
                            size_t* n_consumed, uint64_t* n_items_read) {
    if (!sd || !filename || !n_recs || !n_consumed || (!recs && max_recs))
        return fail(B200SYNC_EINVAL, "null argument");
    CU(cudaSetDevice(sd->device));
    *n_recs = 0;
    *n_consumed = 0;
    if (n_items_read) *n_items_read = 0;
    FILE* f = std::fopen(filename, "rb");  // PM/file_source.hpp:31-37
    if (!f) return fail(B200SYNC_EINVAL, std::string("error opening file: ") + std::strerror(errno));
    struct Closer {
        FILE* f;
        ~Closer() { std::fclose(f); }
    } closer{f};
    if (fseeko(f, 0, SEEK_END) != 0) return fail(B200SYNC_EUNSUPPORTED, "not a seekable file (FIFOs: use b200sync_sd_process)");
    const uint64_t items_in_file = static_cast<uint64_t>(ftello(f)) / sizeof(float2);  // whole items only, like fread
    if (first_item > items_in_file) first_item = items_in_file;
    const size_t n = static_cast<size_t>(std::min<uint64_t>(items_in_file - first_item, max_items));
    if (fseeko(f, static_cast<off_t>(first_item * sizeof(float2)), SEEK_SET) != 0)
        return fail(B200SYNC_EINVAL, std::string("seek failed: ") + std::strerror(errno));
    if (n_items_read) *n_items_read = n;
    const long long S = sd->S, F = sd->fft_size;
    if (n < static_cast<size_t>(F)) return 0;
    CU(sd->d_xoff.ensure(n));
    constexpr int kSlots = 3;
    const size_t piece = 4u << 20;  // 4 Mi items = 32 MiB per read / copy
    if (!sd->h_stage) {
        CU(cudaMallocHost(&sd->h_stage, kSlots * piece * sizeof(float2)));
        for (int i = 0; i < kSlots; ++i) CU(cudaEventCreateWithFlags(&sd->ev_stage[i], cudaEventDisableTiming));
    }
    if (!sd->copy_stream) CU(cudaStreamCreateWithFlags(&sd->copy_stream, cudaStreamNonBlocking));
    cudaStream_t st = sd->stream, cs = sd->copy_stream;
    long long nb_total = 0, P = 0;
    if (int rc = detect_begin(sd, n, nullptr, st, &nb_total, &P)) return rc;
    long long b_done = 0;  // correlator blocks already enqueued
    size_t off = 0;
    for (size_t i = 0; off < n; ++i) {
        const int slot = static_cast<int>(i % kSlots);
        float2* h = sd->h_stage + static_cast<size_t>(slot) * piece;
        if (i >= kSlots) CU(cudaEventSynchronize(sd->ev_stage[slot]));  // its previous copy has left the buffer
        const size_t cnt = std::min(piece, n - off);
        const size_t got = std::fread(h, sizeof(float2), cnt, f);  // PM/file_source.hpp:52
        if (got != cnt) {
            cudaStreamSynchronize(cs);
            cudaStreamSynchronize(st);
            return fail(B200SYNC_EINVAL, std::string("error reading from file: ") +
                                             (std::feof(f) ? "file shrank while reading" : std::strerror(errno)));
        }
        CU(cudaMemcpyAsync(sd->d_xoff.p + off, h, cnt * sizeof(float2), cudaMemcpyHostToDevice, cs));
        CU(cudaEventRecord(sd->ev_stage[slot], cs));
        off += cnt;
        // blocks whose last sample is now on its way: b * S + F <= off
        const long long b_ready = std::min(nb_total, (static_cast<long long>(off) - F) / S + 1);
        if (b_ready > b_done) {
            CU(cudaStreamWaitEvent(st, sd->ev_stage[slot], 0));
            for (long long b0 = b_done; b0 < b_ready; b0 += kOfflineChunkBlocks)
                if (int rc = run_chunk(sd, sd->d_xoff.p, 0, sd->d_zoff.p, 0, b0,
                                       std::min(kOfflineChunkBlocks, b_ready - b0), 0, 0, nullptr, 0, st))
                    return rc;
            b_done = b_ready;
        }
    }
    return detect_finish(sd, sd->d_xoff.p, P, st, recs, max_recs, n_recs, n_consumed);
}

// ------------------------------------------------------------------------------------------
int b200sync_sd_detect_channels_device(b200sync_sd* sd, const void* d_in, size_t channel_stride,
                                       size_t n_channels, size_t n, void* cuda_stream,
                                       b200sync_detection_record* recs, size_t max_recs_per_channel,
                                       size_t* n_recs, size_t* n_consumed) {
    if (!sd || !n_recs || !n_consumed || (!d_in && n && n_channels) || (!recs && max_recs_per_channel))
        return fail(B200SYNC_EINVAL, "null argument");
    if (n_channels > 1 && channel_stride < n) return fail(B200SYNC_EINVAL, "channel_stride smaller than n");
    if (sd->T > kMaxTimeThreshold)
        return fail(B200SYNC_EUNSUPPORTED, "batched channel mode needs time_threshold <= 1023 (one detect_device call per channel works)");
    CU(cudaSetDevice(sd->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const long long S = sd->S, F = sd->fft_size, T = sd->T;
    *n_consumed = 0;
    for (size_t c = 0; c < n_channels; ++c) n_recs[c] = 0;
    if (n_channels == 0 || n < static_cast<size_t>(F)) return 0;
    const long long nb_total = (static_cast<long long>(n) - F) / S + 1;
    const long long P = nb_total * S;
    const long long hi_total = std::max(0LL, P - T - 1);
    const size_t cap = std::max<size_t>(64, static_cast<size_t>(P / (T + 1) + 2));
    // Channels in groups of at most 65535 (gridDim.y) and of at most ~2^31 samples of metric: ONE launch of every
    // kernel per group — correlator over channel x block, peak stage and refine with the channel on blockIdx.y —
    // where round 1 issued the whole kernel sequence once per channel.
    const size_t z_stride = (static_cast<size_t>(P) + 64 + 31) & ~size_t(31);
    const size_t ws_stride = (peak_plan_bytes(std::max(1LL, hi_total), sd->T, sd->num_sms) + 255) & ~size_t(255);
    size_t per_group = std::max<size_t>(1, (size_t(1) << 31) / z_stride);
    per_group = std::min<size_t>({per_group, n_channels, 65535});
    CU(sd->d_chan_state.ensure(n_channels));
    CU(sd->d_chan_recs.ensure(n_channels * cap));
    CU(sd->d_chan_det.ensure(n_channels * cap));
    CU(sd->d_chan_z.ensure(per_group * z_stride));
    CU(sd->d_chan_ws.ensure(per_group * ws_stride));
    const bool use_gm = gm_supported((int)sd->S, sd->T);
    const size_t gm_stride = static_cast<size_t>(nb_total) * gm_groups_per_block((int)sd->S) * kGmF2PerGroup;
    if (use_gm) CU(sd->d_chan_gm.ensure(per_group * gm_stride));
    if (sd->h_chan_cap < n_channels) {
        if (sd->h_chan_state) cudaFreeHost(sd->h_chan_state);
        sd->h_chan_state = nullptr;
        sd->h_chan_cap = 0;
        CU(cudaMallocHost(&sd->h_chan_state, sizeof(PeakState) * n_channels));
        sd->h_chan_cap = n_channels;
    }
    CU(cudaMemsetAsync(sd->d_chan_state.p, 0, sizeof(PeakState) * n_channels, st));
    const float2* base = static_cast<const float2*>(d_in);
    for (size_t c0 = 0; c0 < n_channels; c0 += per_group) {
        const int nch = static_cast<int>(std::min(per_group, n_channels - c0));
        const float2* x = base + c0 * channel_stride;
        PeakState* state = sd->d_chan_state.p + c0;
        CU(launch_correlate(x, 0, sd->d_chan_z.p, 0, sd->d_hperm.p, (int)sd->K, (int)sd->S, sd->fft_arg, 0, nb_total * nch,
                            sd->d_tw.p, nullptr, 0, 0, 0, (int)sd->delay, sd->num_sms, st, nb_total,
                            static_cast<long long>(channel_stride), static_cast<long long>(z_stride),
                            use_gm ? sd->d_chan_gm.p : nullptr, 0, static_cast<long long>(gm_stride)));
        if (hi_total > 0) {
            CU(launch_peak_phase1(sd->d_chan_z.p, 0, P, 0, hi_total, sd->T, sd->power_threshold, sd->d_chan_ws.p,
                                  ws_stride, nullptr, sd->num_sms, st, nch, static_cast<long long>(z_stride), ws_stride,
                                  use_gm ? sd->d_chan_gm.p : nullptr, 0, nb_total, (int)sd->S,
                                  static_cast<long long>(gm_stride)));
            CU(launch_peak_phase2(0, hi_total, sd->T, sd->d_chan_ws.p, ws_stride, -1, state,
                                  sd->d_chan_det.p + c0 * cap, (unsigned)cap, sd->num_sms, st, nch, ws_stride, cap));
        }
        CU(launch_refine(x, 0, sd->d_chan_z.p, 0, sd->d_hperm.p, (int)sd->K, (int)sd->S, sd->fft_arg, sd->min_bin, sd->d_tw.p,
                         sd->d_chan_det.p + c0 * cap, &state->det_count, (unsigned)cap, sd->d_chan_recs.p + c0 * cap,
                         sd->num_sms, st, nch, static_cast<long long>(channel_stride), static_cast<long long>(z_stride),
                         static_cast<long long>(cap)));
    }
    CU(cudaMemcpyAsync(sd->h_chan_state, sd->d_chan_state.p, sizeof(PeakState) * n_channels, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    // exact-size record copies, then the reference's "tag already published" filter per channel
    size_t total = 0;
    for (size_t c = 0; c < n_channels; ++c) {
        if (sd->h_chan_state[c].det_count > cap) return fail(B200SYNC_ENOMEM, "internal detection list overflow");
        total += sd->h_chan_state[c].det_count;
    }
    // into the context's pinned landing buffer: one copy per channel into pageable memory blocks in the driver
    // (64 channels: 2 ms of a 24 ms step)
    if (sd->h_recs_pin_cap < total) {
        if (sd->h_recs_pin) cudaFreeHost(sd->h_recs_pin);
        sd->h_recs_pin = nullptr;
        sd->h_recs_pin_cap = 0;
        CU(cudaMallocHost(&sd->h_recs_pin, sizeof(DetectionRecord) * (total + total / 4 + 64)));
        sd->h_recs_pin_cap = total + total / 4 + 64;
    }
    const DetectionRecord* h = sd->h_recs_pin;
    size_t off = 0;
    for (size_t c = 0; c < n_channels; ++c) {
        const size_t cnt = sd->h_chan_state[c].det_count;
        if (cnt)
            CU(cudaMemcpyAsync(sd->h_recs_pin + off, sd->d_chan_recs.p + c * cap, sizeof(DetectionRecord) * cnt,
                               cudaMemcpyDeviceToHost, st));
        off += cnt;
    }
    CU(cudaStreamSynchronize(st));
    off = 0;
    for (size_t c = 0; c < n_channels; ++c) {
        const size_t cnt = sd->h_chan_state[c].det_count;
        size_t kept = 0;
        for (size_t i = 0; i < cnt; ++i) {
            const DetectionRecord& r = h[off + i];
            if (r.index + sd->delay >= static_cast<uint64_t>(P)) continue;
            if (kept >= max_recs_per_channel) return fail(B200SYNC_ENOMEM, "record buffer too small");
            recs[c * max_recs_per_channel + kept++] = r;
        }
        n_recs[c] = kept;
        off += cnt;
    }
    *n_consumed = static_cast<size_t>(P);
    sd->ev_valid = false;
    return 0;
}

int b200sync_sd_last_timings(const b200sync_sd* sd, float* correlate_ms, float* peaks_ms, float* refine_ms) {
    if (!sd) return fail(B200SYNC_EINVAL, "null context");
    if (!sd->ev_valid) return fail(B200SYNC_EINVAL, "no offline call has completed yet");
    float a = 0, b = 0, c = 0;
    CU(cudaEventElapsedTime(&a, sd->ev[0], sd->ev[1]));
    CU(cudaEventElapsedTime(&b, sd->ev[1], sd->ev[2]));
    CU(cudaEventElapsedTime(&c, sd->ev[2], sd->ev[3]));
    if (correlate_ms) *correlate_ms = a;
    if (peaks_ms) *peaks_ms = b;
    if (refine_ms) *refine_ms = c;
    return 0;
}

int b200sync_sd_copy_metric(const b200sync_sd* sd, float* zpow, size_t n) {
    if (!sd || !zpow) return fail(B200SYNC_EINVAL, "null argument");
    if (!sd->metric_ptr || n > sd->metric_n) return fail(B200SYNC_EINVAL, "no metric of that length available");
    CU(cudaSetDevice(sd->device));
    CU(cudaMemcpy(zpow, sd->metric_ptr, n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

// ------------------------------------------------------------------------------------------
// shard phase 1.  h_in != nullptr: the shard's samples are in HOST memory; they are copied into the context's
// capture buffer in pieces on the copy stream and the correlator runs in chunks behind the copies' events
// (as b200sync_sd_detect_host does for a whole capture).
// f != nullptr: the shard's samples are the next n_in items of the open capture FILE (positioned by the caller); they are
// read piece by piece into the pinned staging ring and copied from there (as b200sync_sd_detect_file does).
static int shard_phase1_core(b200sync_sd* sd, const float2* d_in, const float2* h_in, uint64_t first_sample_abs,
                             size_t n_in, uint64_t first_block, uint64_t n_blocks, uint64_t total_blocks,
                             cudaStream_t st, uint16_t* table, size_t table_len, FILE* f = nullptr) {
    if (sd->T > kMaxTimeThreshold)
        return fail(B200SYNC_EUNSUPPORTED, "time-sharded operation needs time_threshold <= 1023 (the chain tables have T+1 <= 1024 entries)");
    if (table_len < static_cast<size_t>(sd->T) + 1) return fail(B200SYNC_ENOMEM, "table too small");
    if (first_block + n_blocks > total_blocks || n_blocks == 0) return fail(B200SYNC_EINVAL, "bad shard");
    CU(cudaSetDevice(sd->device));
    const long long S = sd->S, F = sd->fft_size, T = sd->T;
    const long long halo = (T + S) / S;  // blocks of metric context needed on each side
    const long long fb = static_cast<long long>(first_block), nbk = static_cast<long long>(n_blocks);
    const long long tb = static_cast<long long>(total_blocks);
    const long long cb0 = std::max(0LL, fb - halo), cb1 = std::min(tb, fb + nbk + halo);
    const long long P_total = tb * S;
    const long long in_base = static_cast<long long>(first_sample_abs);
    if (cb0 * S < in_base || (cb1 - 1) * S + F > in_base + static_cast<long long>(n_in))
        return fail(B200SYNC_EINVAL, "shard input does not cover its halo blocks");
    const long long dec_end = std::max(0LL, P_total - T - 1);
    const long long lo = std::min(fb == 0 ? 0LL : fb * S, dec_end);
    const long long hi = (fb + nbk == tb) ? dec_end : std::min((fb + nbk) * S, dec_end);
    const long long z_base = cb0 * S;
    CU(sd->d_zoff.ensure(static_cast<size_t>((cb1 - cb0) * S) + 64));
    sd->gm_b0 = cb0;
    sd->gm_blocks = 0;
    if (gm_supported((int)sd->S, sd->T)) {
        CU(sd->d_gm.ensure(static_cast<size_t>(cb1 - cb0) * gm_groups_per_block((int)sd->S) * kGmF2PerGroup));
        sd->gm_blocks = cb1 - cb0;
    }
    float2* const gm = sd->gm_blocks > 0 ? sd->d_gm.p : nullptr;
    CU(sd->d_ws.ensure(peak_workspace_bytes_sms(hi - lo + 1, sd->T, sd->num_sms)));
    if (int rc = ensure_det(sd, static_cast<size_t>((hi - lo) / (T + 1) + 2))) return rc;
    if (int rc = reset_state(sd, st)) return rc;
    CU(sd->d_table.ensure(static_cast<size_t>(T) + 1));
    sd->ev_valid = false;
    CU(cudaEventRecord(sd->ev[0], st));
    // optional delayed output (block contract, :318-319): the shard owns output items [fb*S, (fb+nbk)*S) of the
    // P_total the whole capture publishes; they come from blocks fb-1 .. fb+nbk-1, which the halo covers
    float2* d_out = sd->shard.d_out;
    long long out_base = sd->shard.out_first, out_lo = 0, out_hi = 0;
    sd->shard.d_out = nullptr;  // one call only
    if (d_out != nullptr) {
        out_lo = std::max(fb * S, out_base);
        out_hi = std::min({(fb + nbk) * S, P_total, out_base + sd->shard.out_len});
        if (out_lo < static_cast<long long>(sd->delay)) {  // the zero-initialised history (:194-199)
            const long long z1 = std::min<long long>(sd->delay, out_hi);
            if (z1 > out_lo) CU(cudaMemsetAsync(d_out + (out_lo - out_base), 0, (z1 - out_lo) * sizeof(float2), st));
        }
    }
    // host spans: the shard's slice of the delay line is a host copy of samples it already holds (halo included)
    std::vector<std::thread> host_copy;
    struct Joiner {
        std::vector<std::thread>& p;
        ~Joiner() { for (auto& t : p) t.join(); }
    } joiner{host_copy};
    float* h_out = sd->shard.h_out;
    sd->shard.h_out = nullptr;
    // pinned host output span: the correlator writes the slice into d_outoff and finished pieces travel back on the
    // D2H stream (the other PCIe direction) instead of costing the host memory system a read and a write
    float2* h_out_d2h = nullptr;
    long long d2h_done = 0;   // output items (absolute) already on their way back
    size_t d2h_piece = 0;
    if (h_out != nullptr && h_in != nullptr && d_out == nullptr && host_output_by_d2h(true) && is_pinned_host(h_out)) {
        const long long ob = sd->shard.h_out_first;
        out_lo = std::max(fb * S, ob);
        out_hi = std::min({(fb + nbk) * S, P_total, ob + sd->shard.h_out_len});
        if (out_hi > out_lo) {
            CU(sd->d_outoff.ensure(static_cast<size_t>(out_hi - out_lo)));
            d_out = sd->d_outoff.p;
            out_base = out_lo;
            h_out_d2h = reinterpret_cast<float2*>(h_out) + (out_lo - ob);
            d2h_done = out_lo;
            if (!sd->d2h_stream) CU(cudaStreamCreateWithFlags(&sd->d2h_stream, cudaStreamNonBlocking));
            if (out_lo < static_cast<long long>(sd->delay)) {
                const long long z1 = std::min<long long>(sd->delay, out_hi);
                if (z1 > out_lo) CU(cudaMemsetAsync(d_out, 0, (z1 - out_lo) * sizeof(float2), st));
            }
        }
        h_out = nullptr;
    }
    // outputs complete once blocks [.., b_end) have run: items below b_end*S + delay
    auto send_back = [&](long long b_end) -> int {
        if (!h_out_d2h) return 0;
        const long long upto = std::min(out_hi, b_end * S + static_cast<long long>(sd->delay));
        if (upto <= d2h_done) return 0;
        if (sd->ev_out.size() <= d2h_piece) {
            cudaEvent_t e;
            CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            sd->ev_out.push_back(e);
        }
        CU(cudaEventRecord(sd->ev_out[d2h_piece], st));
        CU(cudaStreamWaitEvent(sd->d2h_stream, sd->ev_out[d2h_piece], 0));
        CU(cudaMemcpyAsync(h_out_d2h + (d2h_done - out_lo), d_out + (d2h_done - out_base),
                           static_cast<size_t>(upto - d2h_done) * sizeof(float2), cudaMemcpyDeviceToHost, sd->d2h_stream));
        d2h_done = upto;
        ++d2h_piece;
        return 0;
    };
    if (h_out != nullptr && h_in != nullptr) {
        const long long D = static_cast<long long>(sd->delay), ob = sd->shard.h_out_first;
        const long long lo_o = std::max(fb * S, ob), hi_o = std::min({(fb + nbk) * S, P_total, ob + sd->shard.h_out_len});
        if (hi_o > lo_o) {
            const long long z1 = std::min(D, hi_o);
            if (z1 > lo_o) std::memset(h_out + 2 * (lo_o - ob), 0, static_cast<size_t>(z1 - lo_o) * sizeof(c64));
            const long long c0 = std::max(lo_o, D), cnt = hi_o - c0;
            if (cnt > 0) {
                const size_t nth = std::max<size_t>(1, std::min<size_t>({8, std::thread::hardware_concurrency() / 2,
                                                                         static_cast<size_t>(cnt >> 20)}));
                const float* src = reinterpret_cast<const float*>(h_in) + 2 * (c0 - D - in_base);
                float* dst = h_out + 2 * (c0 - ob);
                for (size_t t = 0; t < nth; ++t) {
                    const size_t i0 = cnt * t / nth, i1 = cnt * (t + 1) / nth;
                    host_copy.emplace_back([=] { std::memcpy(dst + 2 * i0, src + 2 * i0, (i1 - i0) * sizeof(c64)); });
                }
            }
        }
    }
    if (h_in == nullptr && f == nullptr) {
        CU(launch_correlate(d_in, in_base, sd->d_zoff.p, z_base, sd->d_hperm.p, (int)sd->K, (int)sd->S, sd->fft_arg, cb0,
                            cb1 - cb0, sd->d_tw.p, d_out, out_base, out_lo, out_hi, (int)sd->delay, sd->num_sms, st, 0, 0,
                            0, gm, cb0, 0));
    } else {
        CU(sd->d_xoff.ensure(n_in));
        d_in = sd->d_xoff.p;
        if (!sd->copy_stream) CU(cudaStreamCreateWithFlags(&sd->copy_stream, cudaStreamNonBlocking));
        const long long piece = 4LL << 20;  // 4 Mi samples = 32 MiB per copy
        const size_t npieces = (n_in + piece - 1) / piece;
        while (sd->ev_pieces.size() < npieces) {
            cudaEvent_t e;
            CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            sd->ev_pieces.push_back(e);
        }
        constexpr int kSlots = 3;
        if (f != nullptr && !sd->h_stage) {
            CU(cudaMallocHost(&sd->h_stage, kSlots * static_cast<size_t>(piece) * sizeof(float2)));
            for (int i = 0; i < kSlots; ++i) CU(cudaEventCreateWithFlags(&sd->ev_stage[i], cudaEventDisableTiming));
        }
        long long b_next = cb0;  // first correlator block not yet enqueued
        for (size_t i = 0; i < npieces; ++i) {
            const size_t off = i * piece, cnt = std::min<size_t>(piece, n_in - off);
            const float2* src = h_in ? h_in + off : nullptr;
            if (f != nullptr) {
                const int slot = static_cast<int>(i % kSlots);
                float2* h = sd->h_stage + static_cast<size_t>(slot) * piece;
                if (i >= kSlots) CU(cudaEventSynchronize(sd->ev_stage[slot]));  // its previous copy has left the buffer
                if (std::fread(h, sizeof(float2), cnt, f) != cnt) {               // PM/file_source.hpp:52
                    cudaStreamSynchronize(sd->copy_stream);
                    cudaStreamSynchronize(st);
                    return fail(B200SYNC_EINVAL, std::string("error reading from file: ") +
                                                     (std::feof(f) ? "file shrank while reading" : std::strerror(errno)));
                }
                src = h;
            }
            CU(cudaMemcpyAsync(sd->d_xoff.p + off, src, cnt * sizeof(float2), cudaMemcpyHostToDevice, sd->copy_stream));
            CU(cudaEventRecord(sd->ev_pieces[i], sd->copy_stream));
            if (f != nullptr) CU(cudaEventRecord(sd->ev_stage[i % kSlots], sd->copy_stream));
            // correlator chunks this piece completes: blocks whose last sample lies before off + cnt
            const long long have = in_base + static_cast<long long>(off + cnt);
            const long long b_ready = std::min(cb1, have >= F ? (have - F) / S + 1 : 0LL);
            if (b_ready > b_next) {
                CU(cudaStreamWaitEvent(st, sd->ev_pieces[i], 0));
                for (long long b0 = b_next; b0 < b_ready; b0 += kOfflineChunkBlocks)
                    CU(launch_correlate(d_in, in_base, sd->d_zoff.p, z_base, sd->d_hperm.p, (int)sd->K, (int)sd->S, sd->fft_arg, b0,
                                        std::min(kOfflineChunkBlocks, b_ready - b0), sd->d_tw.p, d_out, out_base, out_lo,
                                        out_hi, (int)sd->delay, sd->num_sms, st, 0, 0, 0, gm, cb0, 0));
                b_next = b_ready;
                if (int rc = send_back(b_ready)) return rc;
            }
        }
    }
    CU(cudaEventRecord(sd->ev[1], st));
    if (hi > lo) {
        CU(launch_peak_phase1(sd->d_zoff.p, z_base, cb1 * S, lo, hi, sd->T, sd->power_threshold, sd->d_ws.p,
                              sd->d_ws.cap, sd->d_table.p, sd->num_sms, st, 1, 0, 0, gm, cb0, cb1 - cb0, (int)sd->S, 0));
        CU(cudaMemcpyAsync(table, sd->d_table.p, sizeof(uint16_t) * (T + 1), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    } else {
        CU(cudaStreamSynchronize(st));
        for (long long j = 0; j <= T; ++j) table[j] = static_cast<uint16_t>(j);  // empty range: identity
    }
    if (h_out_d2h) CU(cudaStreamSynchronize(sd->d2h_stream));
    sd->shard.d_in = d_in;
    sd->shard.in_base = in_base;
    sd->shard.z_base = z_base;
    sd->shard.lo = lo;
    sd->shard.hi = hi;
    sd->shard.P_total = P_total;
    sd->shard.st = st;
    sd->shard.valid = true;
    sd->metric_n = static_cast<size_t>((cb1 - cb0) * S);
    sd->metric_ptr = sd->d_zoff.p;
    sd->metric_base = z_base;
    return 0;
}

int b200sync_sd_shard_output(b200sync_sd* sd, void* d_out_delayed, uint64_t out_first_abs, size_t out_len) {
    if (!sd) return fail(B200SYNC_EINVAL, "null context");
    sd->shard.d_out = static_cast<float2*>(d_out_delayed);
    sd->shard.out_first = static_cast<long long>(out_first_abs);
    sd->shard.out_len = static_cast<long long>(out_len);
    return 0;
}

int b200sync_sd_shard_output_host(b200sync_sd* sd, float* out_delayed, uint64_t out_first_abs, size_t out_len) {
    if (!sd) return fail(B200SYNC_EINVAL, "null context");
    sd->shard.h_out = out_delayed;
    sd->shard.h_out_first = static_cast<long long>(out_first_abs);
    sd->shard.h_out_len = static_cast<long long>(out_len);
    return 0;
}

int b200sync_sd_shard_phase1(b200sync_sd* sd, const void* d_in, uint64_t first_sample_abs, size_t n_in,
                             uint64_t first_block, uint64_t n_blocks, uint64_t total_blocks,
                             void* cuda_stream, uint16_t* table, size_t table_len) {
    if (!sd || !d_in || !table) return fail(B200SYNC_EINVAL, "null argument");
    return shard_phase1_core(sd, static_cast<const float2*>(d_in), nullptr, first_sample_abs, n_in, first_block,
                             n_blocks, total_blocks, static_cast<cudaStream_t>(cuda_stream), table, table_len);
}

int b200sync_sd_shard_phase1_host(b200sync_sd* sd, const float* in, uint64_t first_sample_abs, size_t n_in,
                                  uint64_t first_block, uint64_t n_blocks, uint64_t total_blocks, uint16_t* table,
                                  size_t table_len) {
    if (!sd || !in || !table) return fail(B200SYNC_EINVAL, "null argument");
    return shard_phase1_core(sd, nullptr, reinterpret_cast<const float2*>(in), first_sample_abs, n_in, first_block,
                             n_blocks, total_blocks, sd->stream, table, table_len);
}

int b200sync_sd_shard_phase1_file(b200sync_sd* sd, const char* filename, uint64_t capture_first_item,
                                  uint64_t first_sample_abs, size_t n_in, uint64_t first_block, uint64_t n_blocks,
                                  uint64_t total_blocks, uint16_t* table, size_t table_len) {
    if (!sd || !filename || !table) return fail(B200SYNC_EINVAL, "null argument");
    FILE* f = std::fopen(filename, "rb");  // PM/file_source.hpp:31-37
    if (!f) return fail(B200SYNC_EINVAL, std::string("error opening file: ") + std::strerror(errno));
    struct Closer {
        FILE* f;
        ~Closer() { std::fclose(f); }
    } closer{f};
    if (fseeko(f, static_cast<off_t>((capture_first_item + first_sample_abs) * sizeof(float2)), SEEK_SET) != 0)
        return fail(B200SYNC_EINVAL, std::string("seek failed: ") + std::strerror(errno));
    return shard_phase1_core(sd, nullptr, nullptr, first_sample_abs, n_in, first_block, n_blocks, total_blocks,
                             sd->stream, table, table_len, f);
}

int b200sync_sd_shard_phase2(b200sync_sd* sd, uint32_t entry_offset, b200sync_detection_record* recs,
                             size_t max_recs, size_t* n_recs) {
    if (!sd || !n_recs || (!recs && max_recs)) return fail(B200SYNC_EINVAL, "null argument");
    if (!sd->shard.valid) return fail(B200SYNC_EINVAL, "shard_phase1 has not run");
    if (entry_offset > static_cast<uint32_t>(sd->T)) return fail(B200SYNC_EINVAL, "entry offset out of range");
    CU(cudaSetDevice(sd->device));
    *n_recs = 0;
    const auto& sh = sd->shard;
    if (sh.hi > sh.lo) {
        CU(launch_peak_phase2(sh.lo, sh.hi, sd->T, sd->d_ws.p, sd->d_ws.cap, static_cast<int>(entry_offset),
                              sd->d_state.p, sd->d_det_idx.p, (unsigned)sd->d_det_idx.cap, sd->num_sms, sh.st));
    }
    // stage timing of the shard: correlate = ev0..ev1, peaks = phase-1 rest + phase 2 (the wait for the
    // table exchange in between is host time and not counted: ev[2] closes phase 2 only), refine = ev2..ev3
    CU(cudaEventRecord(sd->ev[2], sh.st));
    if (int rc = collect_records(sd, sh.d_in, sh.in_base, sd->d_zoff.p, sh.z_base, sh.st, sd->h_recs)) return rc;
    CU(cudaEventRecord(sd->ev[3], sh.st));
    CU(cudaEventSynchronize(sd->ev[3]));
    sd->ev_valid = true;
    size_t cnt = 0;
    for (const auto& r : sd->h_recs) {
        if (r.index + sd->delay >= static_cast<uint64_t>(sh.P_total)) continue;
        if (cnt >= max_recs) return fail(B200SYNC_ENOMEM, "record buffer too small");
        recs[cnt++] = r;
    }
    *n_recs = cnt;
    sd->shard.valid = false;
    return 0;
}

}  // extern "C"

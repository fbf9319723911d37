// symbol_filter.cu — SymbolFilter (polyphase RRC matched filter, decimation to one sample per
// symbol, re-timed by syncword tags) and SyncwordDetectionFilter (tag gate) behind the C ABI.
//
// Reference semantics
//   SymbolFilter<c64,c64,float>     PM/symbol_filter.hpp:64-252
//   SyncwordDetectionFilter<c64>    PM/syncword_detection_filter.hpp:54-210
//
// SymbolFilter.  Between tags the reference free-runs: a symbol is produced whenever the clock
// phase is 0, as _scale * inner_product(taps[_pfb_arm], history).  A syncword tag re-phases the
// clock, picks the PFB arm from syncword_time_est and the scale from syncword_amplitude, with two
// special cases that emit / swallow one sample (:160-195).  That state machine only acts at tags,
// so the host replays it per tag (cheap, exact integer logic) and emits SEGMENTS
//     (first input index, first output index, clock phase at the first input, arm, scale)
// and the GPU computes every output symbol independently: one thread per symbol, inputs staged
// de-interleaved by clock phase in shared memory (so the stride-sps reads of a warp are
// bank-conflict free), taps in shared memory with an odd arm stride.  Accumulation is tap by tap,
// float multiply then float add, exactly std::inner_product's order: outputs are BIT-EXACT
// against the oracle.
//
// SyncwordDetectionFilter is pure control logic plus a pass-through copy; it stays on the host
//  — with device-resident data the copy is a no-op on the same buffer.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <deque>
#include <new>
#include <string>
#include <vector>

#include "b200sync_internal.h"
#include "cfc.cuh"

namespace b200sync {

struct SfSegment {
    long long in_start;   // absolute input index of the first sample governed by this segment
    long long out_start;  // absolute output index of the first symbol it produces
    int phase0;           // clock phase at in_start (a symbol is produced when the phase is 0)
    int arm;
    float scale;
    int _pad;
};

constexpr int kSfThreads = 256;
constexpr int kSfThreadsC = kSfThreads;

struct SfParams {
    const float2* in;
    long long in_base;   // absolute index of in[0]
    const float2* hist;  // hist_len samples before in_base
    int hist_len;
    float2* out;
    long long out_base;  // absolute index of out[0]
    long long n_out;
    const SfSegment* segs;
    int n_segs;
    int sps, num_arms, arm_size, stride;
    int tile_in;         // staged input samples per CTA
    const CfcSegment* cfc;  // fused CoarseFrequencyCorrection load stage (cfc.cuh); n_cfc = 0: off
    int n_cfc;
};

// Fused CoarseFrequencyCorrection: rotate the staged samples in place.  Thread t owns a run of
// consecutive staged samples so that one sincosf serves 8 of them (cfc.cuh); IDX maps a staged index to
// its shared-memory slot.  Samples before the stream start (zeros) are left alone.
template <typename IDX>
__device__ __forceinline__ void sf_cfc_inplace(const SfParams& P, float2* xs, long long lo_abs, int span, IDX idx) {
    const int per = (span + kSfThreadsC - 1) / kSfThreadsC;
    const int i0 = threadIdx.x * per;
    const int i1 = min(span, i0 + per);
    CfcCursor c;
    for (int i = i0; i < i1; ++i) {
        const long long n = lo_abs + i;
        if (n < 0) continue;
        float2* slot = xs + idx(i);
        *slot = cfc_apply(P.cfc, P.n_cfc, c, n, *slot);
    }
}

__global__ void __launch_bounds__(kSfThreads)
symbol_filter_kernel(const SfParams P, const float* __restrict__ taps_g /*[num_arms][arm_size]*/) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* taps_s = reinterpret_cast<float*>(smem_raw);                        // [num_arms][stride]
    float2* xs = reinterpret_cast<float2*>(taps_s + P.num_arms * P.stride);    // [sps][ph_stride]
    const int ph_stride = P.tile_in / P.sps + 2;
    const int tid = threadIdx.x;
    for (int i = tid; i < P.num_arms * P.arm_size; i += kSfThreads) {
        const int a = i / P.arm_size, k = i - a * P.arm_size;
        taps_s[a * P.stride + k] = taps_g[i];
    }
    const long long o0 = P.out_base + (long long)blockIdx.x * kSfThreads;
    const long long o_end = min(P.out_base + P.n_out, o0 + kSfThreads);
    // segment of an output index: last segment with out_start <= o and at least one symbol at o
    auto locate = [&](long long o, long long& in_idx, int& arm, float& scale) {
        int lo = 0, hi = P.n_segs - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (P.segs[mid].out_start <= o) lo = mid; else hi = mid - 1;
        }
        const SfSegment s = P.segs[lo];
        const int d0 = (P.sps - s.phase0) % P.sps;
        in_idx = s.in_start + d0 + (o - s.out_start) * P.sps;
        arm = s.arm;
        scale = s.scale;
    };
    long long i_first, i_last;
    int a_dummy;
    float s_dummy;
    locate(o0, i_first, a_dummy, s_dummy);
    locate(o_end - 1, i_last, a_dummy, s_dummy);
    const long long lo_abs = i_first - (P.arm_size - 1);
    const long long span_ll = i_last - lo_abs + 1;
    const bool staged = span_ll <= (long long)P.tile_in;
    auto sample = [&](long long a) -> float2 {
        if (a >= P.in_base) return P.in[a - P.in_base];
        const long long h = a - (P.in_base - P.hist_len);
        return h >= 0 ? P.hist[h] : make_float2(0.f, 0.f);
    };
    if (staged) {
        const int span = (int)span_ll;
        for (int i = tid; i < span; i += kSfThreads) {
            const int ph = i % P.sps, m = i / P.sps;
            xs[ph * ph_stride + m] = sample(lo_abs + i);
        }
        if (P.n_cfc > 0) {
            __syncthreads();
            const int sps = P.sps;
            sf_cfc_inplace(P, xs, lo_abs, span, [=](int i) { return (i % sps) * ph_stride + i / sps; });
        }
    }
    __syncthreads();
    CfcCursor cfc_cur;  // unstaged tiles only
    auto sample_c = [&](long long a) -> float2 {
        const float2 v = sample(a);
        return (P.n_cfc > 0 && a >= 0) ? cfc_apply(P.cfc, P.n_cfc, cfc_cur, a, v) : v;
    };
    const long long o = o0 + tid;
    if (o >= o_end) return;
    long long in_idx;
    int arm;
    float scale;
    locate(o, in_idx, arm, scale);
    const float* tp = taps_s + arm * P.stride;
    float2 acc = make_float2(0.f, 0.f);
    if (staged) {
        int rel = (int)(in_idx - lo_abs);
        int ph = rel % P.sps, m = rel / P.sps;
        for (int k = 0; k < P.arm_size; ++k) {
            const float2 h = xs[ph * ph_stride + m];
            const float t = tp[k];
            acc.x = __fadd_rn(acc.x, __fmul_rn(t, h.x));
            acc.y = __fadd_rn(acc.y, __fmul_rn(t, h.y));
            if (--ph < 0) { ph += P.sps; --m; }
        }
    } else {
        for (int k = 0; k < P.arm_size; ++k) {
            const float2 h = sample_c(in_idx - k);
            const float t = tp[k];
            acc.x = __fadd_rn(acc.x, __fmul_rn(t, h.x));
            acc.y = __fadd_rn(acc.y, __fmul_rn(t, h.y));
        }
    }
    P.out[o - P.out_base] = make_float2(__fmul_rn(scale, acc.x), __fmul_rn(scale, acc.y));
}

// ---------------------------------------------------------------------------------------------
// Fast path for the receiver's configuration (samples_per_symbol = SPS, ARM taps per arm; 4 x 44 in
// PM/packet_receiver.hpp:96-115).  ncu on the generic kernel above showed it latency/ALU bound:
// three binary searches over the segment list per thread and 2 shared loads per multiply-add.
//   * sf_tile_seg_kernel finds the segment of every tile's first symbol once (one thread per tile);
//     threads then walk forward from it (tags are thousands of symbols apart).
//   * a thread computes R = 4 consecutive symbols: their windows overlap, so every staged sample is
//     loaded once and used by up to 4 (symbol, tap) pairs; the arm's ARM taps sit in registers.
//     Per symbol: 14 sample loads instead of 88 shared loads.
//   * consecutive threads start 4*SPS samples apart; samples are staged at i + (i >> 4) so that the
//     64-bit window loads of a warp fall on distinct banks.
// The per-symbol accumulation is still tap 0, 1, 2, ... with a separately rounded multiply and add:
// bit-exact against std::inner_product (PM/symbol_filter.hpp:208-215).
// ---------------------------------------------------------------------------------------------
constexpr int kSfR = 4;
constexpr int kSfTileOut = kSfThreads * kSfR;  // 1024 symbols per CTA
__device__ __forceinline__ int sf_skew(int i) { return i + (i >> 4); }

__global__ void sf_tile_seg_kernel(const SfSegment* __restrict__ segs, int n_segs, long long out_base, int tile_out,
                                   int n_tiles, int* __restrict__ tile_seg) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const long long o = out_base + (long long)t * tile_out;
    int lo = 0, hi = n_segs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (segs[mid].out_start <= o) lo = mid; else hi = mid - 1;
    }
    tile_seg[t] = lo;
}

template <int SPS, int ARM, bool CFC>
__global__ void __launch_bounds__(kSfThreads)
symbol_filter_fast_kernel(const SfParams P, const float* __restrict__ taps_g /*[num_arms][ARM]*/,
                          const int* __restrict__ tile_seg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* taps_s = reinterpret_cast<float*>(smem_raw);                        // [num_arms][stride]
    float2* xs = reinterpret_cast<float2*>(taps_s + P.num_arms * P.stride);    // skewed staging
    [[maybe_unused]] float2* ebase_s = xs + (P.tile_in + (P.tile_in >> 4) + 8);  // CFC: one factor per 8 samples
    __shared__ long long span_s[2];
    const int tid = threadIdx.x;
    for (int i = tid; i < P.num_arms * ARM; i += kSfThreads) {
        const int a = i / ARM, k = i - a * ARM;
        taps_s[a * P.stride + k] = __ldg(taps_g + i);
    }
    const long long o0 = P.out_base + (long long)blockIdx.x * kSfTileOut;
    const long long o_end = min(P.out_base + P.n_out, o0 + kSfTileOut);
    const long long o = o0 + (long long)tid * kSfR;     // this thread's first symbol
    // segment of symbol o: walk forward from the tile's first segment
    int sg = tile_seg[blockIdx.x];
    auto in_index = [&](const SfSegment& sgm, long long oo) -> long long {
        const int d0 = (SPS - sgm.phase0) % SPS;
        return sgm.in_start + d0 + (oo - sgm.out_start) * SPS;
    };
    SfSegment cur = P.segs[sg];
    if (tid == 0) span_s[0] = in_index(cur, o0);
    long long next_start = (sg + 1 < P.n_segs) ? P.segs[sg + 1].out_start : LLONG_MAX;
    const long long o_clamped = o < o_end ? o : o_end - 1;   // idle tail threads follow the last symbol
    while (next_start <= o_clamped) {
        ++sg;
        cur = P.segs[sg];
        next_start = (sg + 1 < P.n_segs) ? P.segs[sg + 1].out_start : LLONG_MAX;
    }
    // the thread that owns the tile's last symbol publishes the end of the input span
    if (o <= o_end - 1 && o_end - 1 < o + kSfR) {
        int s2 = sg;
        SfSegment c2 = cur;
        long long nx = next_start;
        while (nx <= o_end - 1) {
            ++s2;
            c2 = P.segs[s2];
            nx = (s2 + 1 < P.n_segs) ? P.segs[s2 + 1].out_start : LLONG_MAX;
        }
        span_s[1] = in_index(c2, o_end - 1);
    }
    __syncthreads();
    const long long lo_abs = span_s[0] - (ARM - 1);
    const long long span_ll = span_s[1] - lo_abs + 1;
    const bool staged = span_ll <= (long long)P.tile_in;
    auto sample = [&](long long a) -> float2 {
        if (a >= P.in_base) return P.in[a - P.in_base];
        const long long h = a - (P.in_base - P.hist_len);
        return h >= 0 ? P.hist[h] : make_float2(0.f, 0.f);
    };
    if (staged) {
        const int span = (int)span_ll;
        bool rotated = false;
        if constexpr (CFC) {
            // Fused CoarseFrequencyCorrection, common case: the whole tile lies in ONE correction segment
            // (resets are a packet apart).  One sincosf per group of 8 samples goes to shared memory, then
            // every sample is rotated on its way from global to shared memory — no second pass.  A thread's
            // samples are kSfThreads apart, so its position in the group (m mod 8) never changes.
            const int sgc = cfc_search(P.cfc, P.n_cfc, lo_abs);   // same address in every thread: broadcast
            const long long c_start = P.cfc[sgc].start;
            const long long c_next = (sgc + 1 < P.n_cfc) ? P.cfc[sgc + 1].start : LLONG_MAX;
            if (lo_abs >= c_start && lo_abs + span <= c_next && lo_abs >= P.in_base) {
                rotated = true;
                const double theta = P.cfc[sgc].theta, phase0 = P.cfc[sgc].phase0;
                const float a0 = P.cfc[sgc].amp0_eps, a1 = P.cfc[sgc].amp_eps;
                const long long m_lo = lo_abs - c_start;
                const long long g_first = m_lo & ~(long long)(kCfcGroup - 1);
                const int n_groups = (int)((m_lo + span - 1 - g_first) >> 3) + 1;
                for (int k = tid; k < n_groups; k += kSfThreads)
                    ebase_s[k] = cfc_group_base(theta, phase0, g_first + (long long)k * kCfcGroup);
                __syncthreads();
                const float2 w = P.cfc[sgc].w[(int)((m_lo + tid) & (kCfcGroup - 1))];
                const float2* src = P.in + (lo_abs - P.in_base);
                for (int i = tid; i < span; i += kSfThreads) {
                    const long long m = m_lo + i;
                    xs[sf_skew(i)] = cfc_rotate(ebase_s[(int)((m - g_first) >> 3)], w, m, a0, a1, __ldcs(src + i));
                }
            }
        }
        if (!rotated) {
            if (lo_abs >= P.in_base) {
                const float2* src = P.in + (lo_abs - P.in_base);
                for (int i = tid; i < span; i += kSfThreads) xs[sf_skew(i)] = __ldcs(src + i);
            } else {
                for (int i = tid; i < span; i += kSfThreads) xs[sf_skew(i)] = sample(lo_abs + i);
            }
            if constexpr (CFC) {  // a reset inside the tile, or history samples: rotate in place
                __syncthreads();
                sf_cfc_inplace(P, xs, lo_abs, span, [](int i) { return sf_skew(i); });
            }
        }
    }
    __syncthreads();
    CfcCursor cfc_cur;  // unstaged tiles only
    auto sample_c = [&](long long a) -> float2 {
        const float2 v = sample(a);
        return (CFC && a >= 0) ? cfc_apply(P.cfc, P.n_cfc, cfc_cur, a, v) : v;
    };
    if (o >= o_end) return;
    float2* dst = P.out + (o - P.out_base);
    const bool fast = staged && (o + kSfR <= o_end) && (o + kSfR - 1 < next_start);
    if (fast) {
        const float* tp = taps_s + cur.arm * P.stride;
        float t[ARM];
#pragma unroll
        for (int k = 0; k < ARM; ++k) t[k] = tp[k];
        const int b = (int)(in_index(cur, o) - lo_abs);   // staged index of symbol o's newest sample
        float2 acc[kSfR];
#pragma unroll
        for (int r = 0; r < kSfR; ++r) acc[r] = make_float2(0.f, 0.f);
        // samples from newest to oldest: symbol r meets tap k = r*SPS - j, i.e. taps in increasing order
#pragma unroll
        for (int j = (kSfR - 1) * SPS; j > -ARM; --j) {
            const float2 h = xs[sf_skew(b + j)];
#pragma unroll
            for (int r = 0; r < kSfR; ++r) {
                const int k = r * SPS - j;
                if (k >= 0 && k < ARM)  // two scalar multiplies + one packed FP32x2 add (see frontend.cu on FFMA2; the opaque -0 form
                                        // of fe_mac measured no faster here: this kernel waits on its loads, not on issue slots)
                    acc[r] = __fadd2_rn(acc[r], make_float2(__fmul_rn(t[k], h.x), __fmul_rn(t[k], h.y)));
            }
        }
        const float sc = cur.scale;
        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
            for (int r = 0; r < kSfR; r += 2)
                *reinterpret_cast<float4*>(dst + r) =
                    make_float4(__fmul_rn(sc, acc[r].x), __fmul_rn(sc, acc[r].y), __fmul_rn(sc, acc[r + 1].x),
                                __fmul_rn(sc, acc[r + 1].y));
        } else {
#pragma unroll
            for (int r = 0; r < kSfR; ++r) dst[r] = make_float2(__fmul_rn(sc, acc[r].x), __fmul_rn(sc, acc[r].y));
        }
        return;
    }
    // symbols next to a tag (segment change inside the thread's group), tile tails, unstaged tiles
    for (int r = 0; r < kSfR; ++r) {
        const long long oo = o + r;
        if (oo >= o_end) break;
        while (next_start <= oo) {
            ++sg;
            cur = P.segs[sg];
            next_start = (sg + 1 < P.n_segs) ? P.segs[sg + 1].out_start : LLONG_MAX;
        }
        const long long in_idx = in_index(cur, oo);
        const float* tp = taps_s + cur.arm * P.stride;
        float2 acc = make_float2(0.f, 0.f);
        for (int k = 0; k < ARM; ++k) {
            const float2 h = staged ? xs[sf_skew((int)(in_idx - k - lo_abs))] : sample_c(in_idx - k);
            const float tk = tp[k];
            acc.x = __fadd_rn(acc.x, __fmul_rn(tk, h.x));
            acc.y = __fadd_rn(acc.y, __fmul_rn(tk, h.y));
        }
        dst[r] = make_float2(__fmul_rn(cur.scale, acc.x), __fmul_rn(cur.scale, acc.y));
    }
}

__global__ void sf_update_hist_kernel(const float2* __restrict__ in, long long n_consumed,
                                      const float2* __restrict__ hist_old, float2* __restrict__ hist_new, int hist_len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hist_len) return;
    const long long src = n_consumed - hist_len + i;
    hist_new[i] = src >= 0 ? in[src] : hist_old[hist_len + src];
}

}  // namespace b200sync

using namespace b200sync;

namespace {
thread_local std::string g_sf_error;
int sf_fail(int code, const std::string& m) {
    g_sf_error = m;
    return code;
}
#define SCU(expr)                                                                                    \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) return sf_fail(B200SYNC_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

struct PendingTag {
    long long pushed_at;  // absolute input index p where the tag arrived (countdown = delay - (i - p))
    b200sync_stream_tag tag;
};
}  // namespace

struct b200sync_sf {
    // settings (PM/symbol_filter.hpp:53-59)
    uint32_t sps = 4, num_arms = 32, delay = 0;
    std::vector<float> taps;
    int device = 0;
    // derived
    int arm_size = 0;
    int reset_clock_phase = 0;
    // state (:42-51)
    int clock_phase = 0, pfb_arm = 0;
    float scale = 1.0f;
    std::deque<PendingTag> pending;
    unsigned long long abs_in = 0, abs_out = 0;
    // device
    float* d_taps = nullptr;
    size_t taps_cap = 0;
    float2* d_hist[2] = { nullptr, nullptr };
    int hist_cur = 0, hist_len = 0, hist_cap = 0;
    SfSegment* d_segs = nullptr;
    size_t segs_cap = 0;
    int* d_tile_seg = nullptr;
    size_t tile_seg_cap = 0;
    float2* d_in = nullptr;
    float2* d_out = nullptr;
    size_t in_cap = 0, out_cap = 0;
    cudaStream_t stream = nullptr;
    // pinned staging for the segment lists of a pipelined bulk call (sf_process_pipelined)
    unsigned char* h_pin = nullptr;
    size_t h_pin_cap = 0, h_pin_used = 0;
    // fused CoarseFrequencyCorrection in front of the filter (b200sync_sf_fuse_cfc)
    bool cfc_on = false;
    uint32_t cfc_delay = 0;
    CfcPlanner cfc;
    std::vector<CfcSeed> cfc_live;
    CfcSeed* d_cfc_seeds = nullptr;
    CfcSegment* d_cfc = nullptr;
    size_t cfc_cap = 0;
};

namespace {

int sf_setup(b200sync_sf* sf) {
    if (sf->sps == 0) return sf_fail(B200SYNC_EINVAL, "samples_per_symbol cannot be zero");  // :67-69
    if (sf->num_arms == 0) return sf_fail(B200SYNC_EINVAL, "num_arms cannot be zero");       // :71-73
    if (sf->taps.empty()) return sf_fail(B200SYNC_EINVAL, "taps cannot be empty");
    sf->arm_size = static_cast<int>((sf->taps.size() + sf->num_arms - 1) / sf->num_arms);
    if (sf->taps.size() % sf->num_arms != 0)
        return sf_fail(B200SYNC_EUNSUPPORTED, "taps.size() must be a multiple of num_arms on the GPU path");
    sf->reset_clock_phase = static_cast<int>((sf->sps - (sf->delay % sf->sps)) % sf->sps);  // :104-107
    SCU(cudaSetDevice(sf->device));
    if (!sf->stream) SCU(cudaStreamCreateWithFlags(&sf->stream, cudaStreamNonBlocking));
    std::vector<float> t(sf->taps.size());
    for (uint32_t j = 0; j < sf->num_arms; ++j)  // polyphase split :84-90
        for (int k = 0; k < sf->arm_size; ++k) t[static_cast<size_t>(j) * sf->arm_size + k] = sf->taps[j + static_cast<size_t>(k) * sf->num_arms];
    // start() on a live context reuses its allocations (cudaFree is a device-wide synchronisation)
    if (sf->taps_cap < t.size()) {
        if (sf->d_taps) cudaFree(sf->d_taps);
        sf->d_taps = nullptr;
        sf->taps_cap = 0;
        SCU(cudaMalloc(&sf->d_taps, t.size() * sizeof(float)));
        sf->taps_cap = t.size();
    }
    SCU(cudaMemcpy(sf->d_taps, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
    sf->hist_len = sf->arm_size;
    for (auto& h : sf->d_hist) {
        if (sf->hist_cap < sf->hist_len) {
            if (h) cudaFree(h);
            h = nullptr;
            SCU(cudaMalloc(&h, sf->hist_len * sizeof(float2)));
        }
        SCU(cudaMemset(h, 0, sf->hist_len * sizeof(float2)));
    }
    if (sf->hist_cap < sf->hist_len) sf->hist_cap = sf->hist_len;
    sf->hist_cur = 0;
    sf->clock_phase = 0;  // start() :110
    sf->pfb_arm = 0;
    sf->scale = 1.0f;
    sf->pending.clear();
    sf->abs_in = sf->abs_out = 0;
    sf->cfc.reset(sf->cfc_delay);
    return 0;
}

// Replays the tag state machine of processBulk (:127-206) over one span and returns segments,
// the number of symbols produced and the re-indexed output tags (:171-183, 218-228).
int sf_plan(b200sync_sf* sf, size_t n_in, const b200sync_stream_tag* in_tags, size_t n_in_tags,
            std::vector<SfSegment>& segs, unsigned long long& n_out, std::vector<b200sync_stream_tag>& out_tags) {
    const int sps = static_cast<int>(sf->sps);
    const long long base_in = static_cast<long long>(sf->abs_in);
    long long out_pos = static_cast<long long>(sf->abs_out);
    long long pos = base_in;            // next input sample to process
    const long long end = base_in + static_cast<long long>(n_in);
    const long long half = sps / 2;
    // symbols produced by free-running from `from` (phase ph) up to `to` (exclusive)
    auto run_count = [&](long long from, long long to, int ph) -> long long {
        const long long d0 = (sps - ph) % sps;
        const long long len = to - from;
        return len > d0 ? (len - d0 + sps - 1) / sps : 0;
    };
    // publish every pending tag whose countdown is < sps/2 at an output instant in [from, to)
    // of a free-running stretch starting with phase ph and output index o_first
    auto flush = [&](long long from, long long to, int ph, long long o_first) {
        const long long d0 = (sps - ph) % sps;
        while (!sf->pending.empty()) {
            const PendingTag& pt = sf->pending.front();
            // first sample index whose countdown (delay - (i - p)) < sps/2
            long long i_min = pt.pushed_at + static_cast<long long>(sf->delay) - half + 1;
            if (i_min < from) i_min = from;
            // first output instant >= i_min in this stretch
            long long first_inst = from + d0;
            if (first_inst < i_min) first_inst += ((i_min - first_inst + sps - 1) / sps) * sps;
            if (first_inst >= to) break;
            b200sync_stream_tag t = pt.tag;
            t.index = static_cast<uint64_t>(o_first + (first_inst - (from + d0)) / sps - static_cast<long long>(sf->abs_out));
            out_tags.push_back(t);
            sf->pending.pop_front();
        }
    };
    size_t ti = 0;
    while (pos < end) {
        // a tag sits on `pos`?
        if (ti < n_in_tags && base_in + static_cast<long long>(in_tags[ti].index) == pos) {
            b200sync_stream_tag tag = in_tags[ti++];
            enum { NONE, CASE_A, CASE_B } special = NONE;
            if (tag.has_syncword) {
                int new_phase = sf->reset_clock_phase;
                sf->scale = 1.0f / tag.sw.syncword_amplitude;  // :139
                float time_est = tag.sw.syncword_time_est;
                if (time_est < 0.0f) {  // :146-156
                    new_phase = (new_phase + 1) % sps;
                    time_est += 1.0f;
                    tag.sw.syncword_phase =
                        static_cast<float>(static_cast<double>(tag.sw.syncword_phase) - tag.sw.syncword_freq);
                }
                if (sf->clock_phase == 0 && new_phase == 1) {
                    // :160-189 — emit one symbol for the tag sample with the OLD arm and the NEW scale;
                    // pending tags are flushed and counted down for that sample
                    segs.push_back({ pos, out_pos, 0, sf->pfb_arm, sf->scale, 0 });
                    flush(pos, pos + 1, 0, out_pos);
                    out_pos += 1;
                    pos += 1;
                    new_phase += 1;
                    special = CASE_A;
                } else if (sf->clock_phase == 1 && new_phase == 0) {
                    // :192-195 — swallow the tag sample: pushed to the history, no symbol, and (as in
                    // the reference) NO countdown step for the pending tags
                    segs.push_back({ pos, out_pos, 1, sf->pfb_arm, sf->scale, 0 });
                    pos += 1;
                    new_phase += 1;
                    special = CASE_B;
                    for (auto& pt : sf->pending) pt.pushed_at += 1;
                }
                // a phase >= sps produces nothing and wraps to 0 on the next sample (:231-233)
                if (new_phase >= sps) new_phase = sps - 1;
                sf->clock_phase = new_phase;
                const float v = std::round(static_cast<float>(sf->num_arms) * time_est);  // :199-202
                long long a = static_cast<long long>(v);
                if (a < 0) a = 0;
                if (a > static_cast<long long>(sf->num_arms) - 1) a = sf->num_arms - 1;
                sf->pfb_arm = static_cast<int>(a);
            }
            // countdown (:204-205): tag.index = delay + adjust, one step per processed sample.  At the
            // output instant of sample i the countdown reads delay - (i - pushed_at):
            //   ordinary tag: pushed_at = p;  case A (adjust = -1, sample p already processed): p;
            //   case B (no adjust, sample p swallowed without a countdown step): p + 1.
            PendingTag pt;
            pt.tag = tag;
            pt.pushed_at = (special == CASE_A) ? pos - 1 : pos;
            sf->pending.push_back(pt);
        }
        // free-run until the next tag (or the end of the span)
        long long stop = end;
        if (ti < n_in_tags) stop = std::min(end, base_in + static_cast<long long>(in_tags[ti].index));
        if (stop < pos) return sf_fail(B200SYNC_EINVAL, "input tags must be sorted by index and inside the span");
        if (stop > pos) {
            segs.push_back({ pos, out_pos, sf->clock_phase, sf->pfb_arm, sf->scale, 0 });
            flush(pos, stop, sf->clock_phase, out_pos);
            out_pos += run_count(pos, stop, sf->clock_phase);
            sf->clock_phase = static_cast<int>((sf->clock_phase + (stop - pos)) % sps);
            pos = stop;
        }
    }
    n_out = static_cast<unsigned long long>(out_pos) - sf->abs_out;
    return 0;
}

// pinned: the segment lists are staged in the context's pinned buffer (room reserved by the caller) and the call
// returns WITHOUT synchronising — the caller runs several spans back to back and synchronises once.
int sf_run(b200sync_sf* sf, const float2* d_in, size_t n_in, const b200sync_stream_tag* in_tags, size_t n_in_tags,
           float2* d_out, size_t max_out, cudaStream_t st, size_t* n_consumed, size_t* n_produced,
           b200sync_stream_tag* out_tags, size_t max_out_tags, size_t* n_out_tags, bool pinned = false) {
    *n_consumed = *n_produced = 0;
    if (n_out_tags) *n_out_tags = 0;
    if (n_in == 0) return 0;
    std::vector<SfSegment> segs;
    std::vector<b200sync_stream_tag> otags;
    segs.reserve(2 * n_in_tags + 2);
    otags.reserve(n_in_tags + sf->pending.size());
    unsigned long long n_out = 0;
    // keep the state so a failing call leaves the block untouched
    const int cp = sf->clock_phase, arm = sf->pfb_arm;
    const float sc = sf->scale;
    const auto pend = sf->pending;
    if (int rc = sf_plan(sf, n_in, in_tags, n_in_tags, segs, n_out, otags)) {
        sf->clock_phase = cp; sf->pfb_arm = arm; sf->scale = sc; sf->pending = pend;
        return rc;
    }
    if (n_out > max_out || otags.size() > max_out_tags) {
        sf->clock_phase = cp; sf->pfb_arm = arm; sf->scale = sc; sf->pending = pend;
        // two causes, two codes: only the tag buffer is something a caller can grow and retry
        if (n_out > max_out) return sf_fail(B200SYNC_ENOSPC, "output span too small");
        return sf_fail(B200SYNC_ENOMEM, "output tag buffer too small");
    }
    if (n_out > 0) {
        // drop segments that produce nothing so the binary search is over producing segments only
        std::vector<SfSegment> prod;
        for (size_t i = 0; i < segs.size(); ++i) {
            const long long next_out = (i + 1 < segs.size()) ? segs[i + 1].out_start
                                                             : static_cast<long long>(sf->abs_out + n_out);
            if (next_out > segs[i].out_start) prod.push_back(segs[i]);
        }
        if (sf->segs_cap < prod.size()) {
            if (sf->d_segs) cudaFree(sf->d_segs);
            sf->d_segs = nullptr;
            SCU(cudaMalloc(&sf->d_segs, (prod.size() + 64) * sizeof(SfSegment)));
            sf->segs_cap = prod.size() + 64;
        }
        const SfSegment* seg_src = prod.data();
        if (pinned) {
            const size_t bytes = prod.size() * sizeof(SfSegment);
            if (sf->h_pin_used + bytes > sf->h_pin_cap) return sf_fail(B200SYNC_ENOMEM, "internal: pinned staging too small");
            std::memcpy(sf->h_pin + sf->h_pin_used, prod.data(), bytes);
            seg_src = reinterpret_cast<const SfSegment*>(sf->h_pin + sf->h_pin_used);
            sf->h_pin_used += (bytes + 63) & ~size_t(63);
        }
        SCU(cudaMemcpyAsync(sf->d_segs, seg_src, prod.size() * sizeof(SfSegment), cudaMemcpyHostToDevice, st));
        SfParams P{};
        P.in = d_in;
        P.in_base = static_cast<long long>(sf->abs_in);
        P.hist = sf->d_hist[sf->hist_cur];
        P.hist_len = sf->hist_len;
        P.out = d_out;
        P.out_base = static_cast<long long>(sf->abs_out);
        P.n_out = static_cast<long long>(n_out);
        P.segs = sf->d_segs;
        P.n_segs = static_cast<int>(prod.size());
        P.sps = static_cast<int>(sf->sps);
        P.num_arms = static_cast<int>(sf->num_arms);
        P.arm_size = sf->arm_size;
        P.stride = sf->arm_size + ((sf->arm_size & 1) ? 0 : 1);
        if (sf->cfc_on) {
            // the frequency correction sees the same tags, before the filter does (PM/packet_receiver.hpp:195-202)
            const long long back = static_cast<long long>(sf->abs_in) - sf->hist_len;
            sf->cfc.advance(n_in, in_tags, n_in_tags);
            sf->cfc.live_segments(back, sf->cfc_live);
            if (pinned) {
                const size_t bytes = sf->cfc_live.size() * sizeof(CfcSeed);
                if (sf->h_pin_used + bytes > sf->h_pin_cap) return sf_fail(B200SYNC_ENOMEM, "internal: pinned staging too small");
                std::memcpy(sf->h_pin + sf->h_pin_used, sf->cfc_live.data(), bytes);
                const CfcSeed* src = reinterpret_cast<const CfcSeed*>(sf->h_pin + sf->h_pin_used);
                sf->h_pin_used += (bytes + 63) & ~size_t(63);
                SCU(cfc_upload_segments(src, sf->cfc_live.size(), &sf->d_cfc_seeds, &sf->d_cfc, &sf->cfc_cap, st));
            } else {
                SCU(cfc_upload_segments(sf->cfc_live, &sf->d_cfc_seeds, &sf->d_cfc, &sf->cfc_cap, st));
            }
            P.cfc = sf->d_cfc;
            P.n_cfc = static_cast<int>(sf->cfc_live.size());
        }
        if (P.sps == 4 && P.arm_size == 44) {
            // the receiver's configuration: register-blocked kernel
            const int n_tiles = static_cast<int>((n_out + kSfTileOut - 1) / kSfTileOut);
            if (sf->tile_seg_cap < static_cast<size_t>(n_tiles)) {
                if (sf->d_tile_seg) cudaFree(sf->d_tile_seg);
                sf->d_tile_seg = nullptr;
                SCU(cudaMalloc(&sf->d_tile_seg, (static_cast<size_t>(n_tiles) + 64) * sizeof(int)));
                sf->tile_seg_cap = static_cast<size_t>(n_tiles) + 64;
            }
            sf_tile_seg_kernel<<<(n_tiles + 255) / 256, 256, 0, st>>>(sf->d_segs, P.n_segs, P.out_base, kSfTileOut,
                                                                      n_tiles, sf->d_tile_seg);
            count_launch();
            SCU(cudaGetLastError());
            P.tile_in = kSfTileOut * P.sps + P.arm_size + 64;  // slack: every tag may add or drop one sample
            const int slots = P.tile_in + (P.tile_in >> 4) + 8;
            const size_t smem = sizeof(float) * P.num_arms * P.stride + sizeof(float2) * static_cast<size_t>(slots) +
                                (sf->cfc_on ? sizeof(float2) * static_cast<size_t>(P.tile_in / kCfcGroup + 4) : 0);
            auto kern = sf->cfc_on ? symbol_filter_fast_kernel<4, 44, true> : symbol_filter_fast_kernel<4, 44, false>;
            SCU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<n_tiles, kSfThreads, smem, st>>>(P, sf->d_taps, sf->d_tile_seg);
            count_launch();
            SCU(cudaGetLastError());
        } else {
            P.tile_in = kSfThreads * P.sps + P.arm_size + 4 * P.sps + 8;
            const size_t smem = sizeof(float) * P.num_arms * P.stride +
                                sizeof(float2) * static_cast<size_t>(P.sps) * (P.tile_in / P.sps + 2);
            SCU(cudaFuncSetAttribute(symbol_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const unsigned grid = static_cast<unsigned>((n_out + kSfThreads - 1) / kSfThreads);
            symbol_filter_kernel<<<grid, kSfThreads, smem, st>>>(P, sf->d_taps);
            count_launch();
            SCU(cudaGetLastError());
        }
        // the pageable `prod` vector must outlive the async copy
        if (!pinned) SCU(cudaStreamSynchronize(st));
    }
    {
        const int nxt = sf->hist_cur ^ 1;
        sf_update_hist_kernel<<<(sf->hist_len + 127) / 128, 128, 0, st>>>(d_in, static_cast<long long>(n_in),
                                                                         sf->d_hist[sf->hist_cur], sf->d_hist[nxt],
                                                                         sf->hist_len);
        count_launch();
        SCU(cudaGetLastError());
        sf->hist_cur = nxt;
    }
    sf->abs_in += n_in;
    sf->abs_out += n_out;
    if (sf->cfc_on) {
        if (n_out == 0) sf->cfc.advance(n_in, in_tags, n_in_tags);  // no kernel ran, the tags still count
        sf->cfc.prune(static_cast<long long>(sf->abs_in) - sf->hist_len);
    }
    *n_consumed = n_in;
    *n_produced = static_cast<size_t>(n_out);
    for (size_t i = 0; i < otags.size(); ++i) out_tags[i] = otags[i];
    if (n_out_tags) *n_out_tags = otags.size();
    return 0;
}

}  // namespace

namespace {
// A bulk span with thousands of tags (a whole capture): the host replay of the tag state machine costs as much as the
// filter kernel (3-4 ms each at 2^30 samples, 43 000 tags).  Cut the span at tag positions into a few sub-spans and
// run them back to back — exactly what consecutive processBulk calls do, the block's state carries across — with the
// segment lists staged in pinned memory and no synchronisation in between: the replay of sub-span i+1 then runs
// while the GPU filters sub-span i.  Only taken when no sub-call can fail for lack of room (so the call still either
// succeeds or leaves the block untouched).
int sf_process_pipelined(b200sync_sf* sf, const float2* d_in, size_t n_in, const b200sync_stream_tag* in_tags,
                         size_t n_in_tags, float2* d_out, size_t max_out, cudaStream_t st, size_t* n_consumed,
                         size_t* n_produced, b200sync_stream_tag* out_tags, size_t max_out_tags, size_t* n_out_tags) {
    const size_t G = std::min<size_t>(16, std::max<size_t>(2, n_in_tags / 2048));
    const size_t need_pin = (2 * n_in_tags + 4 * G + 64) * sizeof(SfSegment) +
                            (sf->cfc_on ? (n_in_tags + 8 * G + 64) * sizeof(CfcSeed) + 64 * G : 0) + 64 * G;
    if (sf->h_pin_cap < need_pin) {
        if (sf->h_pin) cudaFreeHost(sf->h_pin);
        sf->h_pin = nullptr;
        sf->h_pin_cap = 0;
        SCU(cudaMallocHost(&sf->h_pin, need_pin + need_pin / 4));
        sf->h_pin_cap = need_pin + need_pin / 4;
    }
    sf->h_pin_used = 0;
    size_t produced = 0, ntags_out = 0, ti = 0, pos = 0;
    std::vector<b200sync_stream_tag> sub;
    for (size_t g = 0; g < G; ++g) {
        // sub-span g ends where tag number (g+1) * n_tags / G sits (a tag then opens the next sub-span)
        const size_t t_end = (g + 1 == G) ? n_in_tags : (g + 1) * n_in_tags / G;
        const size_t end = (g + 1 == G) ? n_in : static_cast<size_t>(in_tags[t_end].index);
        if (end <= pos) continue;
        sub.assign(in_tags + ti, in_tags + t_end);
        for (auto& t : sub) t.index -= pos;
        size_t c = 0, p = 0, nt = 0;
        if (int rc = sf_run(sf, d_in + pos, end - pos, sub.data(), sub.size(), d_out + produced, max_out - produced, st, &c,
                            &p, out_tags + ntags_out, max_out_tags - ntags_out, &nt, true)) {
            cudaStreamSynchronize(st);
            return rc;
        }
        for (size_t i = 0; i < nt; ++i) out_tags[ntags_out + i].index += produced;   // offsets in the whole call's output
        ntags_out += nt;
        produced += p;
        pos = end;
        ti = t_end;
    }
    SCU(cudaStreamSynchronize(st));
    *n_consumed = n_in;
    *n_produced = produced;
    if (n_out_tags) *n_out_tags = ntags_out;
    return 0;
}
}  // namespace

// ---------------------------------------------------------------------------------------------
// SyncwordDetectionFilter: host control logic (PM/syncword_detection_filter.hpp:54-210)
// ---------------------------------------------------------------------------------------------
struct b200sync_sdf {
    size_t sps = 4, syncword_size = 64, header_size = 128, allowed_margin = 16;  // :43-47
    bool in_packet = false;                                                      // :35-37
    size_t position = 0, block_until = 0;
};

extern "C" {

const char* b200sync_sf_last_error(void) { return g_sf_error.c_str(); }

int b200sync_sf_create(const b200sync_sf_config* cfg, b200sync_sf** out) {
    if (!cfg || !out) return sf_fail(B200SYNC_EINVAL, "null argument");
    *out = nullptr;
    b200sync_sf* sf = new (std::nothrow) b200sync_sf();
    if (!sf) return sf_fail(B200SYNC_ENOMEM, "out of memory");
    sf->sps = cfg->samples_per_symbol;
    sf->num_arms = cfg->num_arms;
    sf->delay = cfg->delay;
    if (cfg->taps && cfg->n_taps) sf->taps.assign(cfg->taps, cfg->taps + cfg->n_taps);
    sf->device = cfg->device;
    const int rc = sf_setup(sf);
    if (rc != 0) {
        const std::string keep = g_sf_error;
        b200sync_sf_destroy(sf);
        g_sf_error = keep;
        return rc;
    }
    *out = sf;
    return 0;
}

void b200sync_sf_destroy(b200sync_sf* sf) {
    if (!sf) return;
    cudaSetDevice(sf->device);
    if (sf->stream) {
        cudaStreamSynchronize(sf->stream);
        cudaStreamDestroy(sf->stream);
    }
    if (sf->d_taps) cudaFree(sf->d_taps);
    for (auto& h : sf->d_hist)
        if (h) cudaFree(h);
    if (sf->d_segs) cudaFree(sf->d_segs);
    if (sf->d_tile_seg) cudaFree(sf->d_tile_seg);
    if (sf->d_cfc) cudaFree(sf->d_cfc);
    if (sf->d_cfc_seeds) cudaFree(sf->d_cfc_seeds);
    if (sf->d_in) cudaFree(sf->d_in);
    if (sf->d_out) cudaFree(sf->d_out);
    if (sf->h_pin) cudaFreeHost(sf->h_pin);
    delete sf;
}

int b200sync_sf_start(b200sync_sf* sf) {
    if (!sf) return sf_fail(B200SYNC_EINVAL, "null context");
    return sf_setup(sf);
}

int b200sync_sf_fuse_cfc(b200sync_sf* sf, int enable, uint32_t cfc_delay) {
    if (!sf) return sf_fail(B200SYNC_EINVAL, "null context");
    if (sf->abs_in != 0) return sf_fail(B200SYNC_EINVAL, "fuse_cfc must be called before the first process call (or start())");
    sf->cfc_on = enable != 0;
    sf->cfc_delay = cfc_delay;
    sf->cfc.reset(cfc_delay);
    return 0;
}

int b200sync_sf_process_device(b200sync_sf* sf, const void* d_in, size_t n_in, const b200sync_stream_tag* in_tags,
                               size_t n_in_tags, void* d_out, size_t max_out, void* cuda_stream, size_t* n_consumed,
                               size_t* n_produced, b200sync_stream_tag* out_tags, size_t max_out_tags,
                               size_t* n_out_tags) {
    if (!sf || !n_consumed || !n_produced || (!d_in && n_in) || (!d_out && max_out) || (!in_tags && n_in_tags) ||
        (!out_tags && max_out_tags))
        return sf_fail(B200SYNC_EINVAL, "null argument");
    SCU(cudaSetDevice(sf->device));
    // many tags and room to spare: pipeline the host replay against the kernel (see sf_process_pipelined)
    if (n_in_tags >= 4096 && max_out >= n_in / sf->sps + n_in_tags + 2 && max_out_tags >= n_in_tags + sf->pending.size())
        return sf_process_pipelined(sf, static_cast<const float2*>(d_in), n_in, in_tags, n_in_tags,
                                    static_cast<float2*>(d_out), max_out, static_cast<cudaStream_t>(cuda_stream), n_consumed,
                                    n_produced, out_tags, max_out_tags, n_out_tags);
    return sf_run(sf, static_cast<const float2*>(d_in), n_in, in_tags, n_in_tags, static_cast<float2*>(d_out), max_out,
                  static_cast<cudaStream_t>(cuda_stream), n_consumed, n_produced, out_tags, max_out_tags, n_out_tags);
}

int b200sync_sf_process(b200sync_sf* sf, const float* in, size_t n_in, const b200sync_stream_tag* in_tags,
                        size_t n_in_tags, float* out, size_t max_out, size_t* n_consumed, size_t* n_produced,
                        b200sync_stream_tag* out_tags, size_t max_out_tags, size_t* n_out_tags) {
    if (!sf || !n_consumed || !n_produced || (!in && n_in) || (!out && max_out) || (!in_tags && n_in_tags) ||
        (!out_tags && max_out_tags))
        return sf_fail(B200SYNC_EINVAL, "null argument");
    SCU(cudaSetDevice(sf->device));
    if (sf->in_cap < n_in) {
        if (sf->d_in) cudaFree(sf->d_in);
        sf->d_in = nullptr;
        SCU(cudaMalloc(&sf->d_in, n_in * sizeof(float2)));
        sf->in_cap = n_in;
    }
    if (sf->out_cap < max_out) {
        if (sf->d_out) cudaFree(sf->d_out);
        sf->d_out = nullptr;
        SCU(cudaMalloc(&sf->d_out, max_out * sizeof(float2)));
        sf->out_cap = max_out;
    }
    SCU(cudaMemcpyAsync(sf->d_in, in, n_in * sizeof(float2), cudaMemcpyHostToDevice, sf->stream));
    const int rc = sf_run(sf, sf->d_in, n_in, in_tags, n_in_tags, sf->d_out, max_out, sf->stream, n_consumed,
                          n_produced, out_tags, max_out_tags, n_out_tags);
    if (rc != 0) return rc;
    SCU(cudaMemcpyAsync(out, sf->d_out, *n_produced * sizeof(float2), cudaMemcpyDeviceToHost, sf->stream));
    SCU(cudaStreamSynchronize(sf->stream));
    return 0;
}

// ---- SyncwordDetectionFilter ----
int b200sync_sdf_create(uint32_t samples_per_symbol, uint32_t syncword_size, uint32_t header_size,
                        b200sync_sdf** out) {
    if (!out) return sf_fail(B200SYNC_EINVAL, "null argument");
    b200sync_sdf* f = new (std::nothrow) b200sync_sdf();
    if (!f) return sf_fail(B200SYNC_ENOMEM, "out of memory");
    f->sps = samples_per_symbol ? samples_per_symbol : 4;
    f->syncword_size = syncword_size ? syncword_size : 64;
    f->header_size = header_size ? header_size : 128;
    *out = f;
    return 0;
}
void b200sync_sdf_destroy(b200sync_sdf* f) { delete f; }
int b200sync_sdf_start(b200sync_sdf* f) {
    if (!f) return sf_fail(B200SYNC_EINVAL, "null context");
    f->in_packet = false;  // :52
    return 0;
}

int b200sync_sdf_process(b200sync_sdf* f, const float* in, size_t n_in, float* out, size_t n_out,
                         const b200sync_stream_tag* tag_in, const b200sync_sdf_header* header, size_t n_ignored,
                         size_t* n_consumed, size_t* header_consumed, size_t* ignored_consumed,
                         b200sync_stream_tag* tag_out, int* tag_forwarded, int* in_packet) {
    if (!f || !n_consumed || !header_consumed || !ignored_consumed || !tag_forwarded)
        return sf_fail(B200SYNC_EINVAL, "null argument");
    *n_consumed = *header_consumed = *ignored_consumed = 0;
    *tag_forwarded = 0;
    auto copy = [&](size_t off, size_t n) {
        if (out && in && out != in) std::memcpy(out + 2 * off, in + 2 * off, n * 2 * sizeof(float));
    };
    if (tag_in) {  // :76-108
        b200sync_stream_tag o{};
        bool any = false, new_in_packet = false;
        if (tag_in->has_syncword && !f->in_packet) {
            new_in_packet = true;
            o = *tag_in;
            o.other = 0;
            any = true;
        }
        if (tag_in->other != 0) {
            o.other = tag_in->other;
            if (!any) o.has_syncword = 0;
            any = true;
        }
        if (new_in_packet) {
            f->in_packet = true;
            f->position = 0;
            f->block_until = 0;
        }
        if (any && tag_out) {
            *tag_out = o;
            tag_out->index = 0;  // published at offset 0 (:105)
            *tag_forwarded = 1;
        }
    }
    if (!f->in_packet) {  // :110-132
        const size_t n = std::min(n_in, n_out);
        copy(0, n);
        *n_consumed = n;
        if (in_packet) *in_packet = 0;
        return 0;
    }
    if (f->block_until == 0 && header) {  // :136-154
        *header_consumed = 1;
        if (header->invalid_header) {
            f->block_until = 1;
        } else {
            if (header->packet_length == 0) return sf_fail(B200SYNC_EINVAL, "received packet_length = 0");
            const size_t payload_symbols = (static_cast<size_t>(header->packet_length) + 4) * 4;
            f->block_until = f->sps * (f->header_size + f->syncword_size - f->allowed_margin + payload_symbols);
        }
    }
    if (f->block_until == 0 && n_ignored > 0) {  // :158-161
        *ignored_consumed = 1;
        f->block_until = 1;
    }
    size_t consumed = 0;
    const size_t allowed = f->sps * (f->syncword_size + f->header_size + f->allowed_margin);
    if (f->position < allowed) {  // :166-172
        const size_t n = std::min({ n_in, n_out, allowed - f->position });
        copy(0, n);
        f->position += n;
        consumed = n;
    }
    if (f->position >= allowed && f->block_until != 0) {  // :174-185
        const size_t n = std::min(n_in, n_out) - consumed;
        copy(consumed, n);
        f->position += n;
        consumed += n;
        if (f->position >= f->block_until) f->in_packet = false;
    }
    *n_consumed = consumed;
    if (in_packet) *in_packet = f->in_packet ? 1 : 0;
    return 0;
}

}  // extern "C"

// cfc.cuh — CoarseFrequencyCorrection (PM/coarse_frequency_correction.hpp:40-98) in closed form.
//
// The reference block is an NCO whose phase and frequency are reset by "syncword_freq" tags:
//     delay samples after a tag:  _exp = (cos(f*delay), -sin(f*delay)),  _exp_incr = (cos f, -sin f),  _counter = 0
//     every sample:               out = in * _exp;  _exp *= _exp_incr;  every 512 samples _exp /= |_exp|
// Between two resets the float recurrence is sequential; here a reset opens a SEGMENT and the factor
// of sample n = start + m is a function of (segment, m) alone:
//     phase(m) = phase0 + m * theta     theta  = atan2(incr.im, incr.re)  of the FLOAT-rounded incr
//                                       phase0 = atan2(exp0.im, exp0.re)  of the float-rounded exp0
//     |exp|(m) = (m < 512 ? |exp0| : 1) * (1 + (m mod 512) * (|incr| - 1))          (renormalisation :88-91)
// evaluated as one sincosf per aligned group of 8 samples (m - m mod 8, argument reduced in double)
// times a per-segment table w[r] = (cos, sin)(r * theta).  Because nothing depends on how the stream
// was cut into calls, streaming == offline and the stand-alone kernel == the load stage fused into
// SymbolFilter, bit for bit.  Against the reference's float recurrence the parity is toleranced
// (rel-L2 < 1e-5 over a packet, like the Rotator: DESIGN.md §3.6).
#pragma once
#include <cuda_runtime.h>

#include <deque>
#include <vector>

#include "b200sync_internal.h"

namespace b200sync {

constexpr int kCfcGroup = 8;

struct CfcSegment {
    long long start;     // absolute index of the first sample the segment governs
    double phase0;       // phase of _exp at `start`
    double theta;        // phase advance per sample
    float amp0_eps;      // |exp0| - 1 (acts on the first 512 samples only)
    float amp_eps;       // |incr| - 1
    float2 w[kCfcGroup]; // (cos, sin)(r * theta)
};

// What the host knows about a segment: where it starts and the four float values set_freq() computes
// (:53-57, host libm like the reference).  cfc_expand_kernel derives the CfcSegment fields from it on
// the device (double atan2 / hypot / cos / sin), so the per-tag host cost is four float libm calls.
struct CfcSeed {
    long long start;
    float e_re, e_im;  // _exp at the reset
    float i_re, i_im;  // _exp_incr
};

// Host side: replays the tag logic of processBulk (:73-79, 81-83, 94-96) and keeps the segments a
// later call may still need (the SymbolFilter history reaches `keep_back` samples behind a call).
class CfcPlanner {
public:
    void reset(uint32_t delay);
    // tags: (index relative to the span, syncword_freq) sorted by index.  Appends the segments that
    // start inside [abs_pos, abs_pos + n) and advances abs_pos.
    void advance(size_t n, const b200sync_stream_tag* tags, size_t n_tags);
    // segments covering [abs_pos_before_call - keep_back, now): call after advance()
    void live_segments(long long from_abs, std::vector<CfcSeed>& out) const;
    void prune(long long from_abs);
    long long abs_pos() const { return abs_pos_; }

private:
    static CfcSeed make_segment(long long start, float freq, uint32_t delay);
    uint32_t delay_ = 0;
    long long abs_pos_ = 0;
    bool pending_ = false;
    float pending_freq_ = 0.0f;
    long long pending_at_ = 0;
    std::deque<CfcSeed> segs_;
};

// seeds (host memory) -> segments in d_segs (capacity grown as needed), asynchronous on st.  The caller
// keeps `seeds` alive until the stream has been synchronised.
cudaError_t cfc_upload_segments(const std::vector<CfcSeed>& seeds, CfcSeed** d_seeds, CfcSegment** d_segs, size_t* cap,
                                cudaStream_t st);
cudaError_t cfc_upload_segments(const CfcSeed* seeds, size_t n_seeds, CfcSeed** d_seeds, CfcSegment** d_segs, size_t* cap,
                                cudaStream_t st);

#ifdef __CUDACC__
// Walks forward from segment index `sg` to the segment of absolute sample n (segments sorted by start).
__device__ __forceinline__ int cfc_seek(const CfcSegment* __restrict__ segs, int n_segs, int sg, long long n) {
    while (sg + 1 < n_segs && segs[sg + 1].start <= n) ++sg;
    return sg;
}
__device__ __forceinline__ int cfc_search(const CfcSegment* __restrict__ segs, int n_segs, long long n) {
    int lo = 0, hi = n_segs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (segs[mid].start <= n) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Per-thread cursor: remembers the group base factor so consecutive samples cost one table rotation.
struct CfcCursor {
    int sg = -1;
    long long gbase = -1;  // m - m mod 8 of the cached factor
    long long seg_start = 0, seg_next = 0;
    float2 ebase;
    float amp0_eps = 0.f, amp_eps = 0.f;
};

// (cos, sin) of the phase at the group base g = m - m mod 8 of a segment
__device__ __forceinline__ float2 cfc_group_base(double theta, double phase0, long long g) {
    double ph = fma((double)g, theta, phase0);
    ph -= 6.283185307179586476925 * rint(ph * 0.15915494309189533577);
    float s, co;
    sincosf((float)ph, &s, &co);
    return make_float2(co, s);
}
// v * ebase * w[m mod 8] * |exp|(m): every path that applies the correction goes through this function
__device__ __forceinline__ float2 cfc_rotate(float2 ebase, float2 w, long long m, float amp0_eps, float amp_eps, float2 v) {
    const float2 e = make_float2(__fmaf_rn(ebase.x, w.x, -__fmul_rn(ebase.y, w.y)),
                                 __fmaf_rn(ebase.x, w.y, __fmul_rn(ebase.y, w.x)));
    float amp = __fmaf_rn((float)(m & 511), amp_eps, 1.0f);
    if (m < 512) amp = __fmaf_rn(amp, amp0_eps, amp);
    const float cr = __fmul_rn(e.x, amp), si = __fmul_rn(e.y, amp);
    return make_float2(__fsub_rn(__fmul_rn(v.x, cr), __fmul_rn(v.y, si)),
                       __fadd_rn(__fmul_rn(v.x, si), __fmul_rn(v.y, cr)));
}

__device__ __forceinline__ float2 cfc_apply(const CfcSegment* __restrict__ segs, int n_segs, CfcCursor& c,
                                            long long n, float2 v) {
    if (c.sg < 0 || n >= c.seg_next || n < c.seg_start) {
        c.sg = (c.sg < 0 || n < c.seg_start) ? cfc_search(segs, n_segs, n) : cfc_seek(segs, n_segs, c.sg, n);
        c.seg_start = segs[c.sg].start;
        c.seg_next = (c.sg + 1 < n_segs) ? segs[c.sg + 1].start : 0x7fffffffffffffffLL;
        c.amp0_eps = segs[c.sg].amp0_eps;
        c.amp_eps = segs[c.sg].amp_eps;
        c.gbase = -1;
    }
    const long long m = n - c.seg_start;
    const long long g = m & ~(long long)(kCfcGroup - 1);
    if (g != c.gbase) {
        c.ebase = cfc_group_base(segs[c.sg].theta, segs[c.sg].phase0, g);
        c.gbase = g;
    }
    return cfc_rotate(c.ebase, segs[c.sg].w[(int)(m & (kCfcGroup - 1))], m, c.amp0_eps, c.amp_eps, v);
}
#endif

}  // namespace b200sync

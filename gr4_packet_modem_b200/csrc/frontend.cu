// frontend.cu — fused PfbArbResampler + Rotator (the SFO/CFO conditioning stage in front of
// SyncwordDetection: apps/packet_transceiver.cpp:71-75) for sm_100a.
//
// Reference semantics
//   PfbArbResampler<c64,c64,float,float>   PM/pfb_arb_resampler.hpp:67-182
//   Rotator<float>                         PM/rotator.hpp:44-65
//
// The reference resampler is a sequential loop over (_last_filter, _phase_acc).  With TRate = float
// every partial sum of the accumulator is exactly representable (all values are multiples of
// 2^-24 and < 2 as long as filter_size/rate >= 1), so output n has a closed form
//     tot = n*Kf,  acc = tot ? ((tot-1) mod 2^24) + 1 : 0,  wraps = (tot-acc) / 2^24      [units 2^-24]
//     Lf  = L0 + n*decim + wraps,  inputs consumed c = Lf / filter_size,  arm = Lf % filter_size
//     y[n] = sum_k taps[arm][k] x[c-1-k]  +  float(acc) * sum_k diff[arm][k] x[c-1-k]
// (verified against the sequential loop by tests/test_oracle_golden.py).  TRate = double (the reference's own QA
// uses it, test/qa_pfb_arb_resampler.cpp:45-69) is the same statement with units of 2^-52 and 128-bit products
// (filter_size/rate >= 1 keeps every partial sum below 2 and a multiple of 2^-52, so the double accumulator is
// exact as well); the interpolation factor is then float(double(acc)), as in the reference (:156-159).
// Each output is then
// independent: one thread computes R consecutive outputs; when they share one arm and advance one
// input per output (always, except at the rare wrap events, for the |rate-1| << 1 of an SFO model)
// a register window slides over the inputs and each tap pair is read once per R outputs.
// The tap-by-tap accumulation order and rounding (float multiply, then float add) are those of
// std::inner_product in the reference, so the resampler output is BIT-EXACT against the oracle.
//
// The rotator's float recurrence (_exp *= _exp_incr, renormalised every 512 samples) is inherently
// sequential; here the phase is closed-form: n * atan2(sin_f, cos_f) reduced in double, amplitude
// |incr|^(n mod 512).  Parity is therefore toleranced.
#include <cmath>
#include <cstring>
#include <new>
#include <numbers>
#include <string>
#include <type_traits>
#include <vector>

#include "b200sync_internal.h"
#include "tma.cuh"

namespace b200sync {

constexpr int kFeThreads = 256;
constexpr int kFeR = 8;                              // outputs per thread
constexpr int kFeTileOut = kFeThreads * kFeR;        // 2048 outputs per CTA
constexpr int kFeTileIn = 2 * kFeTileOut + 128;      // input samples a CTA can stage
// staged inputs live at skew(i) = i + (i >> 3): thread t's register window starts 8 samples after
// thread t-1's, so without the skew the 64-bit window loads of a warp would be 16-way bank conflicted
// (ncu on the unskewed version: 76 % of all shared wavefronts were conflicts)
constexpr int kFeXinSlots = kFeTileIn + (kFeTileIn >> 3) + 8;
__device__ __forceinline__ int fe_skew(int i) { return i + (i >> 3); }

struct FeParams {
    const float2* in;        // in[i] is absolute input sample in_base + i
    long long in_base;       // absolute index of in[0]
    long long in_avail;      // number of valid samples behind `in`
    const float2* hist;      // the hist_len samples that precede in_base (zeros at stream start)
    int hist_len;
    float2* out;             // out[i] is absolute output sample out_base + i
    long long out_base;
    long long n_out;
    unsigned long long Kf;   // filt_rate in units of 2^-24 (TRate = float) or 2^-52 (TRate = double)
    int decim, L0, fs, arm;  // decim_rate, initial _last_filter, filter_size, taps per arm
    int do_resample, do_rotate;
    double theta;            // effective rotation per sample
    float amp_eps;           // |incr| - 1
    float2 wr[kFeR];         // (cos, sin)(r * theta), r = 0..R-1: rotation of output r relative to output 0
    float neg_zero;          // -0.0f, opaque to the compiler (fe_mac)
};

// WIDE = false: TRate = float, units 2^-24, 64-bit products.  WIDE = true: TRate = double, units 2^-52, 128-bit.
template <bool WIDE>
struct FeTime {
    using U = typename std::conditional<WIDE, unsigned __int128, unsigned long long>::type;
    static constexpr int Q = WIDE ? 52 : 24;
    static __device__ __forceinline__ U tot(long long n, unsigned long long Kf) { return (U)(unsigned long long)n * (U)Kf; }
    static __device__ __forceinline__ unsigned long long frac(U t) {   // the accumulator after the step, in units
        return t ? (unsigned long long)((t - 1) & (((U)1 << Q) - 1)) + 1 : 0ull;
    }
    static __device__ __forceinline__ unsigned long long wraps(U t, unsigned long long a) { return (unsigned long long)((t - a) >> Q); }
    static __device__ __forceinline__ float acc(unsigned long long a) {
        // exact in the accumulator's type, then static_cast<float>(_phase_acc) (:156-159)
        return WIDE ? (float)((double)a * (1.0 / 4503599627370496.0)) : (float)a * (1.0f / 16777216.0f);
    }
};

template <bool WIDE>
__device__ __forceinline__ void timing(const FeParams& P, long long n, long long& c, int& arm, float& acc) {
    using FT = FeTime<WIDE>;
    const typename FT::U tot = FT::tot(n, P.Kf);
    const unsigned long long a = FT::frac(tot);
    const unsigned long long wraps = FT::wraps(tot, a);
    const unsigned long long Lf = (unsigned long long)P.L0 + (unsigned long long)n * (unsigned long long)P.decim + wraps;
    c = (long long)(Lf / (unsigned)P.fs);
    arm = (int)(Lf % (unsigned)P.fs);
    acc = FT::acc(a);
}

// (cos, sin) of n * theta, reduced in double (PM/rotator.hpp:56-65 with the recurrence in closed form)
__device__ __forceinline__ float2 rot_phase(const FeParams& P, long long n) {
    double ph = (double)n * P.theta;
    ph -= 6.283185307179586476925 * rint(ph * 0.15915494309189533577);
    float s, c;
    sincosf((float)ph, &s, &c);
    return make_float2(c, s);
}
// v * e * amplitude(n): the reference renormalises _exp every 512 samples, in between |_exp| drifts as
// |incr|^(n mod 512) ~ 1 + (n mod 512) * (|incr| - 1)
__device__ __forceinline__ float2 rot_apply(const FeParams& P, float2 v, float2 e, long long n) {
    const float amp = __fmaf_rn((float)(n & 511), P.amp_eps, 1.0f);
    const float c = e.x * amp, s = e.y * amp;
    return make_float2(__fsub_rn(__fmul_rn(v.x, c), __fmul_rn(v.y, s)), __fadd_rn(__fmul_rn(v.x, s), __fmul_rn(v.y, c)));
}
// Rotation factors of outputs nt .. nt+R-1.  One sincos per ABSOLUTE group of R outputs (n - n mod R),
// the others by a table rotation: the factor of output n is a function of n alone, so the stream does
// not depend on how it was cut into calls (streaming == offline == fused, bit for bit).  nt mod R is
// uniform over the CTA; it is 0 for every CTA of an aligned (offline) call.
__device__ __forceinline__ void rot_factors(const FeParams& P, long long nt, float2 (&e)[kFeR]) {
    const int d = (int)(nt & (kFeR - 1));
    const float2 ea = rot_phase(P, nt - d);
    float2 eb = ea;
    if (d != 0) eb = rot_phase(P, nt - d + kFeR);
#pragma unroll
    for (int r = 0; r < kFeR; ++r) {
        const float2 b = (d + r < kFeR) ? ea : eb;
        const float2 w = P.wr[(d + r) & (kFeR - 1)];
        e[r] = make_float2(__fmaf_rn(b.x, w.x, -__fmul_rn(b.y, w.y)), __fmaf_rn(b.x, w.y, __fmul_rn(b.y, w.x)));
    }
}

// FMA (b200sync_fe_config::fp_contract): every multiply-accumulate of the filter as ONE fused operation (FFMA2, a third of
// the issue slots and half the FP32 pipe passes) instead of std::inner_product's separately rounded multiply and add —
// not bit-identical to the reference any more (one rounding per tap instead of two), opt-in.
// Bit-exact form: the product must be rounded on its own before the add.  A packed multiply feeding a packed add is
// contracted into FFMA2 by ptxas (and so is fma(h, t, -0.0f) with a LITERAL -0, which it first folds into a multiply);
// with the -0 in a register whose value the compiler cannot see (nz, a kernel parameter), the FFMA2 stays a rounded
// product — fma(a, b, -0) == a * b bit for bit, signed zeros included — and the add stays an add: 2 issue slots per
// complex multiply-accumulate instead of the 3 of two scalar multiplies + one packed add.
template <bool FMA>
__device__ __forceinline__ float2 fe_mac(float2 acc, float2 h, float t, float nz) {
    if constexpr (FMA) return __ffma2_rn(h, make_float2(t, t), acc);
    else return __fadd2_rn(acc, __ffma2_rn(h, make_float2(t, t), make_float2(nz, nz)));
}

template <bool WIDE, bool FMA>
__global__ void __launch_bounds__(kFeThreads, 4)
frontend_kernel(const FeParams P, const float2* __restrict__ td_g /*[fs][arm] (tap, diff tap)*/) {
    using FT = FeTime<WIDE>;
    extern __shared__ __align__(128) unsigned char smem_raw[];  // 128 B: half-warp LDS.64 rows never straddle a bank row, whatever static shared data precedes
    float2* td_s = reinterpret_cast<float2*>(smem_raw);   // [fs][arm] (tap, diff tap): one broadcast LDS.64 per tap
    float2* xin = td_s + P.fs * P.arm;                    // [kFeXinSlots], skewed
    const int tid = threadIdx.x;
    const long long n0 = P.out_base + (long long)blockIdx.x * kFeTileOut;     // first output of this CTA
    const long long n_end = min(P.out_base + P.n_out, n0 + kFeTileOut);
    if (!P.do_resample) {
        // Rotator alone: R consecutive outputs per thread share one sincos
        const long long nt = n0 + (long long)tid * kFeR;
        if (nt >= n_end) return;
        const float2* src = P.in + (nt - P.in_base);
        float2* dst = P.out + (nt - P.out_base);
        if (!P.do_rotate) {
#pragma unroll
            for (int r = 0; r < kFeR; ++r)
                if (nt + r < n_end) dst[r] = src[r];
            return;
        }
        float2 e[kFeR];
        rot_factors(P, nt, e);
#pragma unroll
        for (int r = 0; r < kFeR; ++r)
            if (nt + r < n_end) dst[r] = rot_apply(P, src[r], e[r], nt + r);
        return;
    }
    // taps: the global layout is the shared layout, so the whole (tap, diff tap) table is ONE bulk copy by
    // the TMA engine (cp.async.bulk -> UBLKCP) that runs while the threads stage the input samples below
    __shared__ __align__(8) unsigned long long tap_bar;
    const uint32_t tap_bytes = (uint32_t)(P.fs * P.arm) * (uint32_t)sizeof(float2);
    const bool tap_tma = (tap_bytes & 15u) == 0;  // an odd entry count (8 bytes over) takes the plain copy
    if (tap_tma) {
        if (tid == 0) mbar_init(&tap_bar, 1);
        __syncthreads();
        if (tid == 0) tma_load_1d(td_s, td_g, tap_bytes, &tap_bar);
    } else {
        for (int i = tid; i < P.fs * P.arm; i += kFeThreads) td_s[i] = td_g[i];
    }
    // input span of the tile: [c(n0) - arm - 1, c(n_end-1))
    long long c_first, c_last;
    int arm_dummy;
    float acc_dummy;
    timing<WIDE>(P, n0, c_first, arm_dummy, acc_dummy);
    timing<WIDE>(P, n_end - 1, c_last, arm_dummy, acc_dummy);
    const long long lo = c_first - P.arm - 1;  // one spare element: the window prefetch reads x[c-1-arm]
    const int span = (int)min((long long)kFeTileIn + 1, c_last - lo);
    const bool staged = span <= kFeTileIn;
    auto sample = [&](long long a) -> float2 {  // absolute input index -> value (zeros before the stream)
        if (a >= P.in_base) return P.in[a - P.in_base];
        const long long h = a - (P.in_base - P.hist_len);
        return h >= 0 ? P.hist[h] : make_float2(0.f, 0.f);
    };
    if (staged) {
        if (lo >= P.in_base) {
            const float2* src = P.in + (lo - P.in_base);
            for (int i = tid; i < span; i += kFeThreads) xin[fe_skew(i)] = __ldcs(src + i);
        } else {
            for (int i = tid; i < span; i += kFeThreads) xin[fe_skew(i)] = sample(lo + i);
        }
    }
    if (tap_tma) mbar_wait(&tap_bar, 0);
    __syncthreads();

    const long long nt = n0 + (long long)tid * kFeR;  // this thread's first output
    if (nt >= n_end) return;
    // timing of the R outputs: one closed-form evaluation, then increments (no further divisions on
    // the fast path: output r stays on arm0 and one input later iff Lf advanced by exactly r*fs)
    long long c0;
    int arm0;
    float acc[kFeR];
    timing<WIDE>(P, nt, c0, arm0, acc[0]);
    const typename FT::U tot0 = FT::tot(nt, P.Kf);
    const unsigned long long a0 = FT::frac(tot0);
    const unsigned long long wraps0 = FT::wraps(tot0, a0);
    int dL[kFeR];  // Lf(nt + r) - Lf(nt)
    dL[0] = 0;
    bool fast = staged && (nt + kFeR <= n_end);
#pragma unroll
    for (int r = 1; r < kFeR; ++r) {
        const typename FT::U tot = tot0 + (typename FT::U)((unsigned long long)r * P.Kf);
        const unsigned long long aa = FT::frac(tot);
        const unsigned long long wraps = FT::wraps(tot, aa);
        dL[r] = r * P.decim + (int)(wraps - wraps0);
        acc[r] = FT::acc(aa);
        fast = fast && (dL[r] == r * P.fs);
    }
    float2 y[kFeR];
    if (fast) {
        // register window: slot of x[c0-1+j] is (j & 7); output r at tap k uses j = r - k
        const float2* tp = td_s + arm0 * P.arm;
        const int b = (int)(c0 - 1 - lo);  // staged index of x[c0-1]
        float2 W[kFeR];
#pragma unroll
        for (int j = 0; j < kFeR; ++j) W[j] = xin[fe_skew(b + j)];
        float2 af[kFeR], ad[kFeR];
#pragma unroll
        for (int r = 0; r < kFeR; ++r) af[r] = ad[r] = make_float2(0.f, 0.f);
        int k = 0;
        for (; k + kFeR <= P.arm; k += kFeR) {
#pragma unroll
            for (int kk = 0; kk < kFeR; ++kk) {
                const float2 t = tp[k + kk];
#pragma unroll
                for (int r = 0; r < kFeR; ++r) {
                    const float2 h = W[(r - kk) & (kFeR - 1)];
                    // scalar multiplies + one packed FP32x2 add per (re, im) pair: 3 issue slots instead of 4,
                    // each half rounded like the multiply-then-add of std::inner_product.  (A packed
                    // multiply as well would be 2 slots, but ptxas 12.9 contracts mul.rn.f32x2 +
                    // add.rn.f32x2 into FFMA2 even with --fmad=false, which changes the rounding.)
                    af[r] = fe_mac<FMA>(af[r], h, t.x, P.neg_zero);
                    ad[r] = fe_mac<FMA>(ad[r], h, t.y, P.neg_zero);
                }
                // x[c0-1-(k+kk+1)] replaces the element leaving the window
                W[(-(kk + 1)) & (kFeR - 1)] = xin[fe_skew(b - (k + kk + 1))];
            }
        }
        for (; k < P.arm; ++k) {  // arm sizes that are not a multiple of R
            const float2 t = tp[k];
#pragma unroll
            for (int r = 0; r < kFeR; ++r) {
                const float2 h = xin[fe_skew(b + r - k)];
                af[r] = fe_mac<FMA>(af[r], h, t.x, P.neg_zero);
                ad[r] = fe_mac<FMA>(ad[r], h, t.y, P.neg_zero);
            }
        }
#pragma unroll
        for (int r = 0; r < kFeR; ++r) y[r] = fe_mac<FMA>(af[r], ad[r], acc[r], P.neg_zero);
    } else {
#pragma unroll
        for (int r = 0; r < kFeR; ++r) {
            if (nt + r >= n_end) continue;
            const unsigned q = (unsigned)arm0 + (unsigned)dL[r];
            const long long cr = c0 + (long long)(q / (unsigned)P.fs);
            const int armr = (int)(q % (unsigned)P.fs);
            const float2* tp = td_s + armr * P.arm;
            float2 af = make_float2(0.f, 0.f), ad = make_float2(0.f, 0.f);
            for (int k = 0; k < P.arm; ++k) {
                const long long a = cr - 1 - k;
                const float2 h = staged ? xin[fe_skew((int)(a - lo))] : sample(a);
                const float2 t = tp[k];
                af = fe_mac<FMA>(af, h, t.x, P.neg_zero);
                ad = fe_mac<FMA>(ad, h, t.y, P.neg_zero);
            }
            y[r] = fe_mac<FMA>(af, ad, acc[r], P.neg_zero);
        }
    }
    float2* dst = P.out + (nt - P.out_base);
    if (P.do_rotate) {
        float2 e[kFeR];
        rot_factors(P, nt, e);
#pragma unroll
        for (int r = 0; r < kFeR; ++r) y[r] = rot_apply(P, y[r], e[r], nt + r);
    }
    if (nt + kFeR <= n_end && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
        for (int r = 0; r < kFeR; r += 2)
            *reinterpret_cast<float4*>(dst + r) = make_float4(y[r].x, y[r].y, y[r + 1].x, y[r + 1].y);
    } else {
#pragma unroll
        for (int r = 0; r < kFeR; ++r)
            if (nt + r < n_end) dst[r] = y[r];
    }
}

// keep the last hist_len consumed inputs for the next call
__global__ void fe_update_hist_kernel(const float2* __restrict__ in, long long n_consumed,
                                      const float2* __restrict__ hist_old, float2* __restrict__ hist_new, int hist_len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hist_len) return;
    const long long src = n_consumed - hist_len + i;  // index into in[] (may be negative -> old history)
    hist_new[i] = src >= 0 ? in[src] : hist_old[hist_len + src];
}

}  // namespace b200sync

using namespace b200sync;

namespace {
thread_local std::string g_fe_error;
int fe_fail(int code, const std::string& m) {
    g_fe_error = m;
    return code;
}
#define FCU(expr)                                                                                    \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) return fe_fail(B200SYNC_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)
}  // namespace

struct b200sync_fe {
    // settings
    float rate = 1.0f, phase_incr = 0.0f;
    double rate_f64 = 1.0;      // TRate = double (rate_is_f64)
    bool wide = false;
    bool fma = false;           // fp_contract: fused multiply-add in the filter (not bit-identical to the reference)
    std::vector<float> taps;
    uint32_t fs = 32;
    bool do_resample = true, do_rotate = true;
    int device = 0;
    // derived (PM/pfb_arb_resampler.hpp:77-119, PM/rotator.hpp:44-48)
    int arm = 0, decim = 0, L0 = 0;
    unsigned long long Kf = 0;
    double theta = 0.0;
    float amp_eps = 0.0f;
    // state
    unsigned long long abs_in = 0, abs_out = 0;
    float* d_taps = nullptr;
    size_t taps_cap = 0;
    float2* d_hist[2] = { nullptr, nullptr };
    int hist_cur = 0, hist_cap = 0;
    float2* d_in = nullptr;
    float2* d_out = nullptr;
    size_t in_cap = 0, out_cap = 0;
    cudaStream_t stream = nullptr;
    std::string err;
};

namespace {

// number of outputs the reference loop produces when n_in more inputs arrive (PM/pfb_arb_resampler.hpp:129-167):
// output n is produced iff c(n-1) < total_in (the outer `while (in_item < end)`) and c(n) <= total_in
void host_timing(const b200sync_fe* fe, unsigned long long n, unsigned long long& c) {
    const int Q = fe->wide ? 52 : 24;
    const unsigned __int128 tot = (unsigned __int128)n * fe->Kf;
    const unsigned long long a = tot ? (unsigned long long)((tot - 1) & (((unsigned __int128)1 << Q) - 1)) + 1 : 0ull;
    const unsigned long long wraps = (unsigned long long)((tot - a) >> Q);
    const unsigned long long Lf = (unsigned long long)fe->L0 + n * (unsigned long long)fe->decim + wraps;
    c = Lf / fe->fs;
}

unsigned long long outputs_until(const b200sync_fe* fe, unsigned long long total_in) {
    // smallest n_end such that output n_end is NOT produced; c() is non-decreasing in n
    if (total_in == 0) return 0;
    auto produced = [&](unsigned long long n) {
        unsigned long long cn, cp = 0;
        host_timing(fe, n, cn);
        if (n > 0) host_timing(fe, n - 1, cp);
        return cp < total_in && cn <= total_in;
    };
    unsigned long long lo = 0, hi = 1;
    if (!produced(0)) return 0;
    while (produced(hi)) hi *= 2;
    while (hi - lo > 1) {
        const unsigned long long mid = lo + (hi - lo) / 2;
        if (produced(mid)) lo = mid; else hi = mid;
    }
    return hi;
}

int fe_setup(b200sync_fe* fe) {
    if (fe->do_resample) {
        if (fe->fs == 0) return fe_fail(B200SYNC_EINVAL, "filter_size cannot be 0");
        if (fe->taps.size() < 2) return fe_fail(B200SYNC_EINVAL, "taps must have at least 2 entries");
        if (!fe->wide && !(fe->rate > 0.0f)) return fe_fail(B200SYNC_EINVAL, "rate must be positive");
        fe->arm = static_cast<int>((fe->taps.size() + fe->fs - 1) / fe->fs);
        if (fe->wide) {
            // TRate = double (PM/pfb_arb_resampler.hpp:115-118 in double)
            if (!(fe->rate_f64 > 0.0)) return fe_fail(B200SYNC_EINVAL, "rate must be positive");
            const double float_rate = static_cast<double>(fe->fs) / fe->rate_f64;
            if (!(float_rate >= 1.0))
                return fe_fail(B200SYNC_EUNSUPPORTED, "rate > filter_size is not implemented on the GPU path");
            fe->decim = static_cast<int>(std::floor(float_rate));
            const double filt = float_rate - static_cast<double>(fe->decim);
            const double kf = filt * 4503599627370496.0;  // units of 2^-52: exact scaling
            if (kf != std::floor(kf)) return fe_fail(B200SYNC_EUNSUPPORTED, "filt_rate is not a multiple of 2^-52");
            fe->Kf = static_cast<unsigned long long>(kf);
        } else {
            const float float_rate = static_cast<float>(fe->fs) / fe->rate;  // :115
            if (!(float_rate >= 1.0f))
                return fe_fail(B200SYNC_EUNSUPPORTED, "rate > filter_size is not implemented on the GPU path");
            fe->decim = static_cast<int>(std::floor(float_rate));
            const float filt = float_rate - static_cast<float>(fe->decim);
            const double kf = static_cast<double>(filt) * 16777216.0;
            if (kf != std::floor(kf)) return fe_fail(B200SYNC_EUNSUPPORTED, "filt_rate is not a multiple of 2^-24");
            fe->Kf = static_cast<unsigned long long>(kf);
        }
        fe->L0 = static_cast<int>((fe->taps.size() / 2) % fe->fs);
        if (static_cast<size_t>(fe->fs) * (fe->arm + 1) * 8 > 96 * 1024)
            return fe_fail(B200SYNC_EUNSUPPORTED, "tap set too large for shared memory");
    } else {
        fe->arm = 0;
    }
    // Rotator::settingsChanged: _exp_incr = { cos(phase_incr), sin(phase_incr) } in float
    const float ci = std::cos(fe->phase_incr), si = std::sin(fe->phase_incr);
    fe->theta = std::atan2(static_cast<double>(si), static_cast<double>(ci));
    fe->amp_eps = static_cast<float>(std::hypot(static_cast<double>(ci), static_cast<double>(si)) - 1.0);

    FCU(cudaSetDevice(fe->device));
    if (!fe->stream) FCU(cudaStreamCreateWithFlags(&fe->stream, cudaStreamNonBlocking));
    if (fe->do_resample) {
        // polyphase split with zero padding and the derivative filter (:77-102)
        // device layout: td[j][k] = (taps[j][k], diff_taps[j][k]) interleaved
        std::vector<float> t(static_cast<size_t>(2) * fe->fs * fe->arm, 0.0f);
        const size_t nt = fe->taps.size();
        for (uint32_t j = 0; j < fe->fs; ++j) {
            int k = 0;
            for (size_t i = j; i < nt; i += fe->fs) t[2 * (static_cast<size_t>(j) * fe->arm + k++)] = fe->taps[i];
            k = 0;
            for (size_t i = j; i < nt - 1; i += fe->fs)
                t[2 * (static_cast<size_t>(j) * fe->arm + k++) + 1] = fe->taps[i + 1] - fe->taps[i];
        }
        // start() on a live context reuses its allocations (cudaFree is a device-wide synchronisation)
        if (fe->taps_cap < t.size()) {
            if (fe->d_taps) cudaFree(fe->d_taps);
            fe->d_taps = nullptr;
            fe->taps_cap = 0;
            FCU(cudaMalloc(&fe->d_taps, t.size() * sizeof(float)));
            fe->taps_cap = t.size();
        }
        FCU(cudaMemcpy(fe->d_taps, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    const int hl = fe->arm > 0 ? fe->arm : 1;
    for (auto& h : fe->d_hist) {
        if (fe->hist_cap < hl) {
            if (h) cudaFree(h);
            h = nullptr;
            FCU(cudaMalloc(&h, hl * sizeof(float2)));
        }
        FCU(cudaMemset(h, 0, hl * sizeof(float2)));
    }
    if (fe->hist_cap < hl) fe->hist_cap = hl;
    fe->hist_cur = 0;
    fe->abs_in = fe->abs_out = 0;
    return 0;
}

int fe_run(b200sync_fe* fe, const float2* d_in, size_t n_in, float2* d_out, size_t max_out, cudaStream_t st,
           size_t* n_consumed, size_t* n_produced) {
    *n_consumed = *n_produced = 0;
    if (n_in == 0) return 0;
    unsigned long long n_out, consumed = n_in;
    if (fe->do_resample) {
        const unsigned long long end = outputs_until(fe, fe->abs_in + n_in);
        n_out = end - fe->abs_out;
        if (n_out > max_out) {  // output span full: stop after max_out outputs (:129)
            n_out = max_out;
            unsigned long long c;
            host_timing(fe, fe->abs_out + n_out - 1, c);
            consumed = c > fe->abs_in ? c - fe->abs_in : 0;
        }
    } else {
        n_out = n_in < max_out ? n_in : max_out;
        consumed = n_out;
    }
    if (n_out > 0) {
        FeParams P{};
        P.in = d_in;
        P.in_base = static_cast<long long>(fe->abs_in);
        P.in_avail = static_cast<long long>(n_in);
        P.hist = fe->d_hist[fe->hist_cur];
        P.hist_len = fe->arm;
        P.out = d_out;
        P.out_base = static_cast<long long>(fe->abs_out);
        P.n_out = static_cast<long long>(n_out);
        P.Kf = fe->Kf;
        P.decim = fe->decim;
        P.L0 = fe->L0;
        P.fs = static_cast<int>(fe->fs);
        P.arm = fe->arm;
        P.do_resample = fe->do_resample;
        P.do_rotate = fe->do_rotate;
        P.theta = fe->theta;
        P.amp_eps = fe->amp_eps;
        P.neg_zero = -0.0f;
        for (int r = 0; r < kFeR; ++r)
            P.wr[r] = make_float2(static_cast<float>(std::cos(r * fe->theta)), static_cast<float>(std::sin(r * fe->theta)));
        const size_t smem = fe->do_resample
                                ? sizeof(float2) * (static_cast<size_t>(P.fs) * P.arm + static_cast<size_t>(kFeXinSlots))
                                : 0;
        auto kern = fe->wide ? (fe->fma ? frontend_kernel<true, true> : frontend_kernel<true, false>)
                             : (fe->fma ? frontend_kernel<false, true> : frontend_kernel<false, false>);
        FCU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned grid = static_cast<unsigned>((n_out + kFeTileOut - 1) / kFeTileOut);
        kern<<<grid, kFeThreads, smem, st>>>(P, reinterpret_cast<const float2*>(fe->d_taps));
        count_launch();
        FCU(cudaGetLastError());
    }
    if (fe->do_resample && consumed > 0) {
        const int nxt = fe->hist_cur ^ 1;
        fe_update_hist_kernel<<<(fe->arm + 127) / 128, 128, 0, st>>>(d_in, static_cast<long long>(consumed),
                                                                     fe->d_hist[fe->hist_cur], fe->d_hist[nxt], fe->arm);
        count_launch();
        FCU(cudaGetLastError());
        fe->hist_cur = nxt;
    }
    fe->abs_in += consumed;
    fe->abs_out += n_out;
    *n_consumed = static_cast<size_t>(consumed);
    *n_produced = static_cast<size_t>(n_out);
    return 0;
}

}  // namespace

extern "C" {

const char* b200sync_fe_last_error(void) { return g_fe_error.c_str(); }

int b200sync_fe_create(const b200sync_fe_config* cfg, b200sync_fe** out) {
    if (!cfg || !out) return fe_fail(B200SYNC_EINVAL, "null argument");
    *out = nullptr;
    b200sync_fe* fe = new (std::nothrow) b200sync_fe();
    if (!fe) return fe_fail(B200SYNC_ENOMEM, "out of memory");
    fe->rate = cfg->rate;
    fe->wide = cfg->rate_is_f64 != 0;
    fe->fma = cfg->fp_contract != 0;
    fe->rate_f64 = fe->wide ? cfg->rate_f64 : static_cast<double>(cfg->rate);
    fe->phase_incr = cfg->phase_incr;
    if (cfg->taps && cfg->n_taps) fe->taps.assign(cfg->taps, cfg->taps + cfg->n_taps);
    fe->do_resample = cfg->enable_resampler != 0;
    // 0 is an error for the resampler, as in the reference (:70-72); the rotator alone has no filter
    fe->fs = (cfg->filter_size || fe->do_resample) ? cfg->filter_size : 32;
    fe->do_rotate = cfg->enable_rotator != 0;
    fe->device = cfg->device;
    const int rc = fe_setup(fe);
    if (rc != 0) {
        const std::string keep = g_fe_error;
        b200sync_fe_destroy(fe);
        g_fe_error = keep;
        return rc;
    }
    *out = fe;
    return 0;
}

void b200sync_fe_destroy(b200sync_fe* fe) {
    if (!fe) return;
    cudaSetDevice(fe->device);
    if (fe->stream) {
        cudaStreamSynchronize(fe->stream);
        cudaStreamDestroy(fe->stream);
    }
    if (fe->d_taps) cudaFree(fe->d_taps);
    for (auto& h : fe->d_hist)
        if (h) cudaFree(h);
    if (fe->d_in) cudaFree(fe->d_in);
    if (fe->d_out) cudaFree(fe->d_out);
    delete fe;
}

int b200sync_fe_start(b200sync_fe* fe) {
    if (!fe) return fe_fail(B200SYNC_EINVAL, "null context");
    return fe_setup(fe);
}

size_t b200sync_fe_max_output(const b200sync_fe* fe, size_t n_in) {
    if (!fe) return 0;
    if (!fe->do_resample) return n_in;
    return static_cast<size_t>(static_cast<double>(n_in) * (fe->wide ? fe->rate_f64 : static_cast<double>(fe->rate)) * 1.0001) + 64;
}

int b200sync_fe_process_device(b200sync_fe* fe, const void* d_in, size_t n_in, void* d_out, size_t max_out,
                               void* cuda_stream, size_t* n_consumed, size_t* n_produced) {
    if (!fe || !n_consumed || !n_produced || (!d_in && n_in) || (!d_out && max_out))
        return fe_fail(B200SYNC_EINVAL, "null argument");
    FCU(cudaSetDevice(fe->device));
    return fe_run(fe, static_cast<const float2*>(d_in), n_in, static_cast<float2*>(d_out), max_out,
                  static_cast<cudaStream_t>(cuda_stream), n_consumed, n_produced);
}

int b200sync_fe_process(b200sync_fe* fe, const float* in, size_t n_in, float* out, size_t max_out,
                        size_t* n_consumed, size_t* n_produced) {
    if (!fe || !n_consumed || !n_produced || (!in && n_in) || (!out && max_out))
        return fe_fail(B200SYNC_EINVAL, "null argument");
    FCU(cudaSetDevice(fe->device));
    if (fe->in_cap < n_in) {
        if (fe->d_in) cudaFree(fe->d_in);
        fe->d_in = nullptr;
        FCU(cudaMalloc(&fe->d_in, n_in * sizeof(float2)));
        fe->in_cap = n_in;
    }
    if (fe->out_cap < max_out) {
        if (fe->d_out) cudaFree(fe->d_out);
        fe->d_out = nullptr;
        FCU(cudaMalloc(&fe->d_out, max_out * sizeof(float2)));
        fe->out_cap = max_out;
    }
    FCU(cudaMemcpyAsync(fe->d_in, in, n_in * sizeof(float2), cudaMemcpyHostToDevice, fe->stream));
    const int rc = fe_run(fe, fe->d_in, n_in, fe->d_out, max_out, fe->stream, n_consumed, n_produced);
    if (rc != 0) return rc;
    FCU(cudaMemcpyAsync(out, fe->d_out, *n_produced * sizeof(float2), cudaMemcpyDeviceToHost, fe->stream));
    FCU(cudaStreamSynchronize(fe->stream));
    return 0;
}

}  // extern "C"

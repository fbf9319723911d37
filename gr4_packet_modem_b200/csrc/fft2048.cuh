// fft2048.cuh — register/shared-memory FFT-2048 building blocks for sm_100a.
//
// The overlap-save correlator (PM/syncword_detection.hpp:236-252 in the reference)
// needs, per 2048-sample block, one forward FFT and K "inverse-as-forward" FFTs.
// On the CPU that is FFTW; here it is a 3-pass 16 x 16 x 8 decomposition executed by
// a 128-thread group: every thread holds 16 complex points in registers, small DFTs
// run entirely in registers, and the two transposes between passes go through a
// padded shared-memory exchange buffer laid out so that every LDS.64/STS.64 warp
// access is the minimum 2 wavefronts (no bank conflicts).
//
//   FFT "A" (samples -> spectrum), n = 128 n1 + 8 n2 + n3, k = k1 + 16 k2 + 256 k3
//     pass A1  DFT16 over n1   x W_256^{n2 k1}            thread = (n2,n3) = tid
//     pass A2  DFT16 over n2   x W_2048^{n3 (k1+16 k2)}   thread = (n3,k1)
//     pass A3  DFT8  over n3                              thread owns p = k1+16 k2 in {tid, tid+128}
//   FFT "B" (spectrum product -> correlation), f = f1 + 16 f2 + 256 f3, m = 128 m1 + 8 m2 + m3
//     pass B1  DFT8  over f3   x W_2048^{(f1+16 f2) m3}   same ownership as A3 (no exchange!)
//     pass B2  DFT16 over f2   x W_256^{f1 m2}            thread = (m3,f1)
//     pass B3  DFT16 over f1                              thread = (m2,m3) = tid -> C[128 m1 + tid]
//
// A's output ordering is exactly B's input ordering, so the spectrum never has to be
// un-permuted: the template spectra are stored pre-permuted instead.
//
// ARITHMETIC CONTRACT (mirrored bit-for-bit by oracle/oracle_fft.hpp, FftKind::Mirror):
//   * every add/sub/mul/fma is a separately rounded IEEE binary32 op (explicit _rn intrinsics), issued as
//     packed FP32x2 instructions; the complex multiply is
//         cmul(a,w) = ( fma(a.x, w.x, -(a.y*w.y)),  fma(a.x, w.y, a.y*w.x) )
//     and the squared magnitude  norm2(z) = fma(z.y, z.y, z.x*z.x);
//   * small DFTs are radix-2 decimation-in-frequency stages: (u,v) -> (u+v, (u-v)*w);
//     w = 1 is skipped, w = -i is a swap/negate; W16^{1,3,5,7} use cmul with constants equal to
//     twiddle-table entries; w = W8^1 / W8^3 leave the UNSCALED terms
//         e1(d) = (d.x + d.y, d.y - d.x),   e3(d) = (d.y - d.x, (-d.x) + (-d.y))
//     and the factor c = float(sqrt(1/2)) is applied by the next stage's butterfly, which always pairs the
//     two such elements:  a = c*ea;  sum = fma(c, eb, a);  dif = fma(-c, eb, a);
//   * all inter-pass twiddles are entries of one table  Wt[j] = (float(cos(2 pi j/2048)), float(-sin(2 pi j/2048)))
//     computed in double on the host; a statically-zero exponent skips the multiply.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "b200sync_internal.h"
#include "tmem.cuh"

namespace b200sync {

// Twiddle tables.  Every entry is an entry of the one table Wt[j] of the arithmetic contract; they
// are only re-ordered so that, for each unrolled pass iteration, every half-warp of a 64-bit
// shared-memory load reads 16 consecutive float2 (or a broadcast): no bank conflicts.  (The first
// version indexed Wt directly: tw[8*f1*m2] puts 16 lanes on one bank; ncu showed 63% of all shared
// wavefronts were load conflicts — profiles/r1_ncu_summary.md.)
//   twS[a*16 + b]  = Wt[8*a*b]   (256)    pass B2: a = m2, b = f1;   pass A1: a = k1, b = n2
//   twL[a*256 + p] = Wt[a*p]     (2048)   pass B1: a = m3, p;        pass A2: a = n3, p = k1 + 16 k2
// The forward and the inverse-as-forward transform need the same two sets of factors, so they
// share the tables (18 KiB of shared memory instead of 36: the rest is L1 for the template spectra).
constexpr int kTwS = 0, kTwL = 256;
constexpr int kTwTotal = 256 + 2048;  // 2304 float2 = 18432 B

constexpr int kXchgStrideA = 129;   // exchange layout 1: (a*129 + tid)
constexpr int kXchgStrideP = 9;     // exchange layout 2: (p*9 + d)
constexpr int kXchgFloat2 = 256 * kXchgStrideP;  // 2304 float2 = 18432 B (>= 16*129 = 2064)
constexpr int kXchgFloat2A = 16 * kXchgStrideA;  // 2064 float2 = 16512 B: layout 1 alone

// Complex arithmetic on Blackwell's packed FP32x2 pipe forms (FADD2 / FMUL2 / FFMA2, sm_100_rt.h): a float2
// is one aligned register pair, one instruction works on both halves, each half rounded exactly like the
// scalar IEEE op.  The operand modifiers ptxas folds for free (seen in SASS, profiles/r2_sass_excerpt.txt):
// scalar broadcast `R.F32`, swapped halves `.LO_HI`, whole-pair negation and a per-half sign `.NP`.  With
// them a complex multiply is TWO instructions and a butterfly's sum and difference one each — half the
// issue slots of the scalar form; FP32 pipe time is unchanged (a packed instruction occupies the pipe for two
// passes), so the kernel moves from issue-bound to FP32-pipe-bound.
// Caution, measured with ptxas 12.9: `mul.rn.f32x2` feeding `add.rn.f32x2` IS contracted into FFMA2 even with
// -fmad=false.  The contract below therefore never lets a rounded packed product feed a packed add: every
// multiply-then-add is written as an explicit fma (w8pair), so there is nothing left for ptxas to fuse.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
// a * w = ( fma(a.x, w.x, -(a.y*w.y)),  fma(a.x, w.y, a.y*w.x) ):
//   FMUL2 p = a.y * w.LO_HI ;  FFMA2 r = a.x * w + (-p.x, +p.y)
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
    const float2 p = __fmul2_rn(make_float2(a.y, a.y), make_float2(w.y, w.x));
    return __ffma2_rn(make_float2(a.x, a.x), w, make_float2(-p.x, p.y));
}
__device__ __forceinline__ float norm2(float2 z) { return __fmaf_rn(z.y, z.y, __fmul_rn(z.x, z.x)); }
// (x,y) * -i
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

#define B200_SQRT1_2 0.70710678118654752440f
#define B200_COS_PI_8 0.92387953251128675613f
#define B200_SIN_PI_8 0.38268343236508977173f

// W8^1 = c(1-i) and W8^3 = -c(1+i), c = sqrt(1/2): d * W8^1 = c * e1(d), d * W8^3 = c * e3(d) with
//   e1(d) = (d.x + d.y, d.y - d.x),   e3(d) = (d.y - d.x, (-d.x) + (-d.y)).
// In a radix-2 DIF DFT-8/16 the two elements that carry these factors always meet in the NEXT stage's
// butterfly, so the scale is applied there (w8pair) instead of being rounded on its own.
__device__ __forceinline__ float2 w8e1(float2 d) { return __fadd2_rn(d, make_float2(d.y, -d.x)); }
__device__ __forceinline__ float2 w8e3(float2 d) { return __fadd2_rn(make_float2(d.y, -d.x), make_float2(-d.x, -d.y)); }
// butterfly of u = c*ea and v = c*eb:  a = rn(c*ea);  sum = fma(c, eb, a);  dif = fma(-c, eb, a)
__device__ __forceinline__ void w8pair(float2& ea, float2& eb) {
    const float2 cc = make_float2(B200_SQRT1_2, B200_SQRT1_2);
    const float2 a = __fmul2_rn(cc, ea);
    ea = __ffma2_rn(cc, eb, a);
    eb = __ffma2_rn(make_float2(-B200_SQRT1_2, -B200_SQRT1_2), eb, a);
}
// plain butterfly: (u, v) -> (u + v, u - v)
__device__ __forceinline__ void bf(float2& u, float2& v) {
    const float2 s = cadd(u, v);
    v = csub(u, v);
    u = s;
}
// W16^e for odd e: table-entry constants, general complex multiply
template <int E>
__device__ __forceinline__ float2 mul_w16_odd(float2 a) {
    static_assert(E == 1 || E == 3 || E == 5 || E == 7, "odd exponents only");
    if constexpr (E == 1) return cmul(a, make_float2(B200_COS_PI_8, -B200_SIN_PI_8));
    else if constexpr (E == 3) return cmul(a, make_float2(B200_SIN_PI_8, -B200_COS_PI_8));
    else if constexpr (E == 5) return cmul(a, make_float2(-B200_SIN_PI_8, -B200_COS_PI_8));
    else return cmul(a, make_float2(-B200_COS_PI_8, -B200_SIN_PI_8));
}

// In-place radix-2 DIF DFT-16.  Input v[n], output v[bitrev4(k)] = X[k].
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
    // half = 8, twiddle W16^i on the difference
#pragma unroll
    for (int i = 0; i < 8; ++i) bf(v[i], v[i + 8]);
    v[9] = mul_w16_odd<1>(v[9]);
    v[10] = w8e1(v[10]);                 // x c, applied in the next stage
    v[11] = mul_w16_odd<3>(v[11]);
    v[12] = mul_mi(v[12]);
    v[13] = mul_w16_odd<5>(v[13]);
    v[14] = w8e3(v[14]);                 // x c, applied in the next stage
    v[15] = mul_w16_odd<7>(v[15]);
    // half = 4, twiddle W8^i
    bf(v[0], v[4]); bf(v[1], v[5]); bf(v[2], v[6]); bf(v[3], v[7]);
    v[5] = w8e1(v[5]); v[6] = mul_mi(v[6]); v[7] = w8e3(v[7]);
    bf(v[8], v[12]); bf(v[9], v[13]); w8pair(v[10], v[14]); bf(v[11], v[15]);
    v[13] = w8e1(v[13]); v[14] = mul_mi(v[14]); v[15] = w8e3(v[15]);
    // half = 2, twiddle W4^i
    bf(v[0], v[2]); bf(v[1], v[3]); v[3] = mul_mi(v[3]);
    bf(v[4], v[6]); w8pair(v[5], v[7]); v[7] = mul_mi(v[7]);
    bf(v[8], v[10]); bf(v[9], v[11]); v[11] = mul_mi(v[11]);
    bf(v[12], v[14]); w8pair(v[13], v[15]); v[15] = mul_mi(v[15]);
    // half = 1
#pragma unroll
    for (int g = 0; g < 16; g += 2) bf(v[g], v[g + 1]);
}

// In-place radix-2 DIF DFT-8.  Input v[n], output v[bitrev3(k)] = X[k].
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
    bf(v[0], v[4]); bf(v[1], v[5]); bf(v[2], v[6]); bf(v[3], v[7]);
    v[5] = w8e1(v[5]); v[6] = mul_mi(v[6]); v[7] = w8e3(v[7]);
    bf(v[0], v[2]); bf(v[1], v[3]); v[3] = mul_mi(v[3]);
    bf(v[4], v[6]); w8pair(v[5], v[7]); v[7] = mul_mi(v[7]);
#pragma unroll
    for (int g = 0; g < 8; g += 2) bf(v[g], v[g + 1]);
}

__host__ __device__ constexpr int bitrev4(int k) {
    return ((k & 1) << 3) | ((k & 2) << 1) | ((k & 4) >> 1) | ((k & 8) >> 3);
}
__host__ __device__ constexpr int bitrev3(int k) { return ((k & 1) << 2) | (k & 2) | ((k & 4) >> 2); }

// 128-thread named barrier for one FFT group (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void group_sync(int bar_id) {
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(kGroupThreads) : "memory");
}

// ---- FFT A: 2048 samples (registers v[n1] = x[128 n1 + tid]) -> spectrum -------------
// On return xs[pi*8 + k3] = X[(tid + 128 pi) + 256 k3]   (pi in {0,1}).
// `tw` are the per-pass twiddle tables (shared memory), `xb` this group's exchange buffer.
//
// TWO_BUF: the two exchanges use two separate buffers (xb: layout 2 "p*9+d", xb2: layout 1
// "a*129+tid"), so a buffer is never rewritten before every thread has passed the barrier that
// follows its reads: the two "buffer free again" barriers of each transform disappear.  The caller
// must place one barrier between a transform that read xb2 last (fft_b) and the next fft_a.
template <bool TWO_BUF = false>
__device__ __forceinline__ void fft_a(float2 (&v)[16], float2 (&xs)[16], const float2* __restrict__ tw,
                                      float2* __restrict__ xb, int tid, int bar_id,
                                      float2* __restrict__ xb2 = nullptr) {
    float2* const xa = TWO_BUF ? xb2 : xb;
    // pass A1
    dft16(v);
    {
        const int n2 = tid >> 3;
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
            float2 val = v[bitrev4(k1)];
            if (k1 != 0) val = cmul(val, tw[kTwS + k1 * 16 + n2]);
            xa[k1 * kXchgStrideA + tid] = val;
        }
    }
    group_sync(bar_id);
    // pass A2
    const int n3 = tid >> 4, k1 = tid & 15;
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = xa[k1 * kXchgStrideA + n2 * 8 + n3];
    if constexpr (!TWO_BUF) group_sync(bar_id);
    dft16(v);
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
        float2 val = v[bitrev4(k2)];
        val = cmul(val, tw[kTwL + n3 * 256 + k1 + 16 * k2]);
        xb[(k1 + 16 * k2) * kXchgStrideP + n3] = val;
    }
    group_sync(bar_id);
    // pass A3
#pragma unroll
    for (int pi = 0; pi < 2; ++pi) {
        float2 w[8];
        const int p = tid + 128 * pi;
#pragma unroll
        for (int d = 0; d < 8; ++d) w[d] = xb[p * kXchgStrideP + d];
        dft8(w);
#pragma unroll
        for (int k3 = 0; k3 < 8; ++k3) xs[pi * 8 + k3] = w[bitrev3(k3)];
    }
    group_sync(bar_id);  // exchange buffer free again
}

// ---- FFT A with its inter-pass factors in TMEM (correlate_kernel, K <= kTmHypA) ---------------------------
// Same passes, same table entries, same multiplies as fft_a (bit-identical): the 15 + 16 factors of a thread come
// from its TMEM lane (columns kTmA1 / kTmA2 of the layout below) instead of shared memory.
__device__ __forceinline__ void fft_a_tm(float2 (&v)[16], float2 (&xs)[16], float2* __restrict__ xb, int tid,
                                         int bar_id, uint32_t tm_tw) {
    constexpr int TmA1 = 256, TmA2 = 288;   // == kTmA1, kTmA2 (declared with the layout further down)
    float2 t[8];
    // pass A1
    tmem_ld8(tm_tw + TmA1, t);
    dft16(v);
    tmem_wait_ld();
    tmem_use(t);
#pragma unroll
    for (int k1 = 0; k1 < 9; ++k1) {
        float2 val = v[bitrev4(k1)];
        if (k1 != 0) val = cmul(val, t[k1 > 0 ? k1 - 1 : 0]);
        xb[k1 * kXchgStrideA + tid] = val;
    }
    tmem_ld8(tm_tw + TmA1 + 16, t);
    tmem_wait_ld();
    tmem_use(t);
#pragma unroll
    for (int k1 = 9; k1 < 16; ++k1) xb[k1 * kXchgStrideA + tid] = cmul(v[bitrev4(k1)], t[k1 - 9]);
    group_sync(bar_id);
    // pass A2
    const int n3 = tid >> 4, k1 = tid & 15;
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = xb[k1 * kXchgStrideA + n2 * 8 + n3];
    tmem_ld8(tm_tw + TmA2, t);
    group_sync(bar_id);
    dft16(v);
    tmem_wait_ld();
    tmem_use(t);
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) xb[(k1 + 16 * k2) * kXchgStrideP + n3] = cmul(v[bitrev4(k2)], t[k2]);
    tmem_ld8(tm_tw + TmA2 + 16, t);
    tmem_wait_ld();
    tmem_use(t);
#pragma unroll
    for (int k2 = 8; k2 < 16; ++k2) xb[(k1 + 16 * k2) * kXchgStrideP + n3] = cmul(v[bitrev4(k2)], t[k2 - 8]);
    group_sync(bar_id);
    // pass A3
#pragma unroll
    for (int pi = 0; pi < 2; ++pi) {
        float2 w[8];
        const int p = tid + 128 * pi;
#pragma unroll
        for (int d = 0; d < 8; ++d) w[d] = xb[p * kXchgStrideP + d];
        dft8(w);
#pragma unroll
        for (int k3 = 0; k3 < 8; ++k3) xs[pi * 8 + k3] = w[bitrev3(k3)];
    }
    group_sync(bar_id);  // exchange buffer free again
}

// ---- FFT B: y[pi*8 + f3] = Y[(tid + 128 pi) + 256 f3] -> c[m1] = C[128 m1 + tid] ------
// y is destroyed.  The exchange buffer must be free on entry and is free on return.
//
// REG_B1 / REG_B2: the inter-pass factors of pass B1 / B2 depend only on the thread, not on the
// hypothesis, so a caller that runs many transforms back to back (correlate_kernel: K per block) can
// hold them in registers (load_b1_twiddles / load_b2_twiddles) instead of re-reading shared memory
// K times.  Same table entries, same multiplies: bit-identical.
__device__ __forceinline__ void load_b1_twiddles(float2 (&t1)[14], const float2* __restrict__ tw, int tid) {
#pragma unroll
    for (int pi = 0; pi < 2; ++pi)
#pragma unroll
        for (int m3 = 1; m3 < 8; ++m3) t1[pi * 7 + m3 - 1] = tw[kTwL + m3 * 256 + tid + 128 * pi];
}
__device__ __forceinline__ void load_b2_twiddles(float2 (&t2)[15], const float2* __restrict__ tw, int tid) {
#pragma unroll
    for (int m2 = 1; m2 < 16; ++m2) t2[m2 - 1] = tw[kTwS + m2 * 16 + (tid & 15)];
}

template <bool TWO_BUF = false, bool REG_B1 = false, bool REG_B2 = false>
__device__ __forceinline__ void fft_b(float2 (&y)[16], float2 (&c)[16], const float2* __restrict__ tw,
                                      float2* __restrict__ xb, int tid, int bar_id,
                                      float2* __restrict__ xb2 = nullptr, const float2* t1 = nullptr,
                                      const float2* t2 = nullptr) {
    float2* const xa = TWO_BUF ? xb2 : xb;
    // pass B1
#pragma unroll
    for (int pi = 0; pi < 2; ++pi) {
        float2 w[8];
        const int p = tid + 128 * pi;
#pragma unroll
        for (int d = 0; d < 8; ++d) w[d] = y[pi * 8 + d];
        dft8(w);
#pragma unroll
        for (int m3 = 0; m3 < 8; ++m3) {
            float2 val = w[bitrev3(m3)];
            if (m3 != 0) val = cmul(val, REG_B1 ? t1[pi * 7 + (m3 > 0 ? m3 - 1 : 0)] : tw[kTwL + m3 * 256 + p]);
            xb[p * kXchgStrideP + m3] = val;
        }
    }
    group_sync(bar_id);
    // pass B2
    const int m3 = tid >> 4, f1 = tid & 15;
#pragma unroll
    for (int f2 = 0; f2 < 16; ++f2) c[f2] = xb[(f1 + 16 * f2) * kXchgStrideP + m3];
    if constexpr (!TWO_BUF) group_sync(bar_id);
    dft16(c);
#pragma unroll
    for (int m2 = 0; m2 < 16; ++m2) {
        float2 val = c[bitrev4(m2)];
        if (m2 != 0) val = cmul(val, REG_B2 ? t2[m2 > 0 ? m2 - 1 : 0] : tw[kTwS + m2 * 16 + f1]);
        xa[f1 * kXchgStrideA + m2 * 8 + m3] = val;
    }
    group_sync(bar_id);
    // pass B3
#pragma unroll
    for (int ff = 0; ff < 16; ++ff) y[ff] = xa[ff * kXchgStrideA + tid];
    if constexpr (!TWO_BUF) group_sync(bar_id);  // exchange buffer free again
    dft16(y);
#pragma unroll
    for (int m1 = 0; m1 < 16; ++m1) c[m1] = y[bitrev4(m1)];
}

// ---- FFT B with its operands in TMEM (correlate_kernel) ------------------------------------------------
// Same passes, same table entries, same multiplies as fft_b (bit-identical); what changes is where the
// operands come from.  Everything a thread needs that does not depend on the other threads of the group —
// its 16 spectrum points, its 16 template points, its 14 + 15 inter-pass factors — is read from the thread's
// own TMEM lane with tcgen05.ld (tmem.cuh) instead of from registers / L1 / shared memory: the LSU pipe, which
// bound the kernel (83 % of its wavefront peak in round 1), keeps only the two exchanges, and the spectrum no
// longer occupies 32 registers for the whole hypothesis loop.
// TMEM column layout of a CTA (512 columns; the lane is the thread's index inside its FFT group):
//   [  0,  32)  pass-B1 factors: 16 columns per pi, entry m3-1 = Wt[(tid + 128 pi) m3], m3 = 1..7
//   [ 32,  64)  pass-B2 factors: entry m2-1 = Wt[8 (tid & 15) m2], m2 = 1..15
//   [ 64, 320)  conj template spectra of the first kTmHyp = 8 hypotheses, 32 columns each (layout of hperm)
//   [320, 512)  block spectrum of each of the 6 FFT groups, 32 columns per group
constexpr int kTmT1 = 0, kTmT2 = 32, kTmH = 64, kTmHyp = 8, kTmXs = 320, kTmCols = 512;
// With K <= kTmHypA hypotheses the last two template slots are free and hold FFT A's inter-pass factors instead
// (fft_a_tm): then the forward transform, too, keeps only its two exchanges on the LSU pipe.  It is half of the
// work at K = 1, the one HBM-relevant configuration.
//   [256, 288)  pass-A1 factors: entry k1-1 = Wt[8 (tid >> 3) k1], k1 = 1..15
//   [288, 320)  pass-A2 factors: entry k2 = Wt[(tid >> 4) ((tid & 15) + 16 k2)], k2 = 0..15
constexpr int kTmHypA = 6, kTmA1 = 256, kTmA2 = 288;
static_assert(kTmH + 32 * kTmHypA <= kTmA1 && kTmA2 + 32 <= kTmXs, "FFT A's factors live in the last two template slots");

// load_h(pi, h): fills h[0..8) with the conj template points of half pi (a tcgen05.ld or eight global loads)
template <class LoadH>
__device__ __forceinline__ void fft_b_tm(float2 (&c)[16], float2* __restrict__ xb, int tid, int bar_id,
                                         uint32_t tm_xs, uint32_t tm_tw, LoadH&& load_h) {
    // pass B1
#pragma unroll
    for (int pi = 0; pi < 2; ++pi) {
        float2 xsr[8], h[8], t[8], w[8];
        const int p = tid + 128 * pi;
        tmem_ld8(tm_xs + 16 * pi, xsr);
        load_h(pi, h);
        tmem_ld8(tm_tw + kTmT1 + 16 * pi, t);
        tmem_wait_ld();
        tmem_use(xsr);
        tmem_use(h);
        tmem_use(t);
#pragma unroll
        for (int d = 0; d < 8; ++d) w[d] = cmul(xsr[d], h[d]);  // PM/syncword_detection.hpp:247-249
        dft8(w);
#pragma unroll
        for (int m3 = 0; m3 < 8; ++m3) {
            float2 val = w[bitrev3(m3)];
            if (m3 != 0) val = cmul(val, t[m3 > 0 ? m3 - 1 : 0]);
            xb[p * kXchgStrideP + m3] = val;
        }
    }
    group_sync(bar_id);
    // pass B2
    const int m3 = tid >> 4, f1 = tid & 15;
    float2 y[16];
    {
        float2 t[8];
#pragma unroll
        for (int f2 = 0; f2 < 16; ++f2) y[f2] = xb[(f1 + 16 * f2) * kXchgStrideP + m3];
        tmem_ld8(tm_tw + kTmT2, t);
        group_sync(bar_id);
        dft16(y);
        tmem_wait_ld();
        tmem_use(t);
#pragma unroll
        for (int m2 = 0; m2 < 9; ++m2) {
            float2 val = y[bitrev4(m2)];
            if (m2 != 0) val = cmul(val, t[m2 > 0 ? m2 - 1 : 0]);
            xb[f1 * kXchgStrideA + m2 * 8 + m3] = val;
        }
        tmem_ld8(tm_tw + kTmT2 + 16, t);
        tmem_wait_ld();
        tmem_use(t);
#pragma unroll
        for (int m2 = 9; m2 < 16; ++m2) {
            const float2 val = cmul(y[bitrev4(m2)], t[m2 - 9]);
            xb[f1 * kXchgStrideA + m2 * 8 + m3] = val;
        }
    }
    group_sync(bar_id);
    // pass B3
#pragma unroll
    for (int ff = 0; ff < 16; ++ff) y[ff] = xb[ff * kXchgStrideA + tid];
    group_sync(bar_id);  // exchange buffer free again
    dft16(y);
#pragma unroll
    for (int m1 = 0; m1 < 16; ++m1) c[m1] = y[bitrev4(m1)];
}

}  // namespace b200sync

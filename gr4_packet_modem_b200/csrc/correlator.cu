// correlator.cu — overlap-save syncword correlator for sm_100a.
//
// Replaces the hot loops of SyncwordDetection::processBulk
// (PM/syncword_detection.hpp:236-252 forward FFT + per-hypothesis product + FFT,
//  :299-313 per-sample best hypothesis) and the template construction FFT of
// start() (:183-188).  One 128-thread group owns one 2048-sample block at a time:
//   global (coalesced, streaming) -> registers -> FFT A -> spectrum stays in
//   registers -> for each hypothesis: x conj-template (pre-permuted, L1/L2
//   resident) -> FFT B -> |.|^2 -> running max in registers -> zpow (4 B/sample).
// Nothing but the per-sample winning power leaves the SM.  The fields the estimator
// needs (winning complex correlation, neighbour-bin powers, bin, block noise power,
// PM/syncword_detection.hpp:326-342) are recomputed bit-identically by
// refine_kernel for the sparse set of detected peaks only.
#include "b200sync_internal.h"
#include "fft2048.cuh"
#include "tma.cuh"

namespace b200sync {

__device__ __forceinline__ void load_twiddles(float2* tw_s, const float2* __restrict__ tw_g) {
    for (int i = threadIdx.x; i < kTwTotal; i += blockDim.x) tw_s[i] = tw_g[i];
}

// ---------------------------------------------------------------------------------
// Template spectra: Hperm[k][j][tid] = conj(FFT_A(shifted_syncword_k))[(tid+128*(j>>3)) + 256*(j&7)]
// grid = K groups of 128 threads (one CTA each).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGroupThreads)
template_spectra_kernel(const float2* __restrict__ td /*[K][2048] zero padded*/,
                        float2* __restrict__ hperm, const float2* __restrict__ tw_g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];  // 128 B: half-warp LDS.64 rows never straddle a bank row, whatever static shared data precedes
    float2* tw_s = reinterpret_cast<float2*>(smem_raw);
    float2* xb = tw_s + kTwTotal;
    load_twiddles(tw_s, tw_g);
    __syncthreads();
    const int tid = threadIdx.x;
    const int k = blockIdx.x;
    float2 v[16], xs[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) v[n1] = td[(size_t)k * kFft + 128 * n1 + tid];
    fft_a(v, xs, tw_s, xb, tid, 1);
#pragma unroll
    for (int j = 0; j < 16; ++j)
        hperm[((size_t)k * 16 + j) * kGroupThreads + tid] = make_float2(xs[j].x, -xs[j].y);
}

// ---------------------------------------------------------------------------------
// Main correlator.  Persistent: gridDim.x CTAs x (blockDim.x/128) groups; group gg
// handles blocks gg, gg+G, ... of [b0, b0+nb).  Block b covers absolute samples
// [b*S, b*S+2048) and produces zpow for [b*S, (b+1)*S).
// ---------------------------------------------------------------------------------
#ifndef B200_TWO_XB
constexpr bool kCorrTwoBuf = false;
#else
// one buffer per exchange layout (2 barriers per transform instead of 4) needs 228 KB of shared memory at 6
// groups: measured SLOWER (6.66 vs 6.17 ms at 2^28, K = 9) because it leaves no L1 for the template spectra
constexpr bool kCorrTwoBuf = true;
#endif
constexpr int kCorrXchg = kCorrTwoBuf ? kXchgFloat2 + kXchgFloat2A : kXchgFloat2;
#ifdef B200_REG_B1
constexpr bool kRegB1 = true;
#else
constexpr bool kRegB1 = false;
#endif
#ifdef B200_REG_B2
constexpr bool kRegB2 = true;
#else
constexpr bool kRegB2 = false;
#endif

#ifdef B200_NO_TMEM
constexpr bool kCorrTmem = false;  // round-1 form: spectrum in registers, templates through L1, factors in shared memory
#else
constexpr bool kCorrTmem = true;
static_assert(kCorrThreads == 6 * kGroupThreads, "the TMEM column layout (fft2048.cuh) is for 6 FFT groups per CTA");
#endif

__global__ void __launch_bounds__(kCorrThreads, 1)
correlate_kernel(const float2* __restrict__ in, long long in_base, float* __restrict__ zpow,
                 long long z_base, const float2* __restrict__ hperm, int K, int S, long long b0,
                 long long nb, const float2* __restrict__ tw_g, float2* __restrict__ out_delayed,
                 long long out_base, long long out_lo, long long out_hi, int delay) {
    extern __shared__ __align__(128) unsigned char smem_raw[];  // 128 B: half-warp LDS.64 rows never straddle a bank row, whatever static shared data precedes
    float2* tw_s = reinterpret_cast<float2*>(smem_raw);
    const int g = threadIdx.x >> 7;
    const int tid = threadIdx.x & 127;
    float2* xb = tw_s + kTwTotal + g * kCorrXchg;
    float2* xb2 = kCorrTwoBuf ? xb + kXchgFloat2 : xb;
#ifdef B200_TMA_TW
    // TMA staging of the twiddle table is available but NOT the default here: the copy itself is free (once
    // per persistent CTA), yet the build with it measured 1.7 % slower over the whole launch (24.94 vs 24.52 ms
    // at 2^30, same box, A/B twice; not shared-memory alignment — a 128-B aligned base gave the same time —
    // but a slightly different register allocation / schedule of the steady-state loop).  The front end's tap
    // table does use it (frontend.cu), where it is neutral to slightly faster.
    {   // the 18 KiB twiddle table: one TMA bulk copy per persistent CTA
        __shared__ __align__(8) unsigned long long tw_bar;
        if (threadIdx.x == 0) mbar_init(&tw_bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) tma_load_1d(tw_s, tw_g, kTwTotal * sizeof(float2), &tw_bar);
        mbar_wait(&tw_bar, 0);
    }
#else
    load_twiddles(tw_s, tw_g);
#endif
    // ---- TMEM: allocate the CTA's 512 columns and park the per-thread constants there (fft2048.cuh) ----
    __shared__ uint32_t tm_slot;
    uint32_t tm_base = 0, tm_xs = 0;
    const int Ktm = K < kTmHyp ? K : kTmHyp;  // hypotheses whose template lives in TMEM; the rest come through L1
    if constexpr (kCorrTmem) {
        if (threadIdx.x < 32) tmem_alloc(&tm_slot, kTmCols);
        tmem_fence_before_sync();
    }
    __syncthreads();  // twiddles staged, TMEM address published
    if constexpr (kCorrTmem) {
        tmem_fence_after_sync();
        tm_base = tmem_addr(tm_slot, threadIdx.x >> 5, 0);
        tm_xs = tm_base + kTmXs + 32 * g;
        const int ngr = blockDim.x >> 7;
        float2 r[8];
        // templates: group g parks hypotheses g, g + ngr, ... (every group reads all of them later)
        for (int k = g; k < Ktm; k += ngr) {
            const float2* h = hperm + (size_t)k * 16 * kGroupThreads + tid;
#pragma unroll
            for (int pi = 0; pi < 2; ++pi) {
#pragma unroll
                for (int d = 0; d < 8; ++d) r[d] = __ldg(h + (pi * 8 + d) * kGroupThreads);
                tmem_st8(tm_base + kTmH + 32 * k + 16 * pi, r);
            }
        }
        if (g == ngr - 1) {  // inter-pass factors of FFT B, the entries fft_b reads from shared memory
#pragma unroll
            for (int pi = 0; pi < 2; ++pi) {
#pragma unroll
                for (int m3 = 1; m3 < 8; ++m3) r[m3 - 1] = tw_s[kTwL + m3 * 256 + tid + 128 * pi];
                r[7] = make_float2(0.f, 0.f);
                tmem_st8(tm_base + kTmT1 + 16 * pi, r);
            }
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int m2 = 1 + 8 * hf + e;
                    r[e] = m2 < 16 ? tw_s[kTwS + m2 * 16 + (tid & 15)] : make_float2(0.f, 0.f);
                }
                tmem_st8(tm_base + kTmT2 + 16 * hf, r);
            }
        }
        tmem_wait_st();
        tmem_fence_before_sync();
        __syncthreads();
        tmem_fence_after_sync();
    }
    const int ngroups = blockDim.x >> 7;
    const long long gstride = (long long)gridDim.x * ngroups;
    const int bar_id = 1 + g;

    float2 t1[14], t2[15];
    if constexpr (kRegB1) load_b1_twiddles(t1, tw_s, tid);
    if constexpr (kRegB2) load_b2_twiddles(t2, tw_s, tid);

    for (long long blk = (long long)blockIdx.x * ngroups + g; blk < nb; blk += gstride) {
        const long long s0 = (b0 + blk) * (long long)S;  // absolute first sample of the block
        const float2* src = in + (s0 - in_base);
        float2 v[16], xs[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) v[n1] = __ldcs(src + 128 * n1 + tid);
        if (out_delayed != nullptr) {
            // block contract: out[n] = in[n - delay] (PM/syncword_detection.hpp:318-319).  Only output items
            // [out_lo, out_hi) are stored: out_hi = the number of items the call publishes (:346) — nothing is
            // stored at or past it — and a time shard stores only the slice it owns.
            // This block owns samples [s0, s0+S); it has them in registers already.
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) {
                const int i = 128 * n1 + tid;
                const long long o = s0 + i + delay;
                if (i < S && o >= out_lo && o < out_hi) out_delayed[o - out_base] = v[n1];
            }
        }
        if constexpr (kCorrTwoBuf) group_sync(bar_id);  // previous block's last reads of xb2 are done
        fft_a<kCorrTwoBuf>(v, xs, tw_s, xb, tid, bar_id, xb2);

        float best[16];
#pragma unroll
        for (int m1 = 0; m1 < 16; ++m1) best[m1] = -1.0f;  // :303
        if constexpr (kCorrTmem) {
            // the spectrum moves to the thread's TMEM lane: 32 registers free for the hypothesis loop
            {
                float2 r[8];
#pragma unroll
                for (int pi = 0; pi < 2; ++pi) {
#pragma unroll
                    for (int d = 0; d < 8; ++d) r[d] = xs[pi * 8 + d];
                    tmem_st8(tm_xs + 16 * pi, r);
                }
                tmem_wait_st();
            }
            for (int k = 0; k < K; ++k) {
                float2 c[16];
                const uint32_t tm_h = tm_base + kTmH + 32 * k;
                const float2* hg = hperm + (size_t)k * 16 * kGroupThreads + tid;
                const bool in_tm = k < Ktm;  // uniform over the CTA
                fft_b_tm(c, xb, tid, bar_id, tm_xs, tm_base, [&](int pi, float2 (&h)[8]) {   // :250-251
                    if (in_tm) {
                        tmem_ld8(tm_h + 16 * pi, h);
                    } else {
#pragma unroll
                        for (int d = 0; d < 8; ++d) h[d] = __ldg(hg + (pi * 8 + d) * kGroupThreads);
                    }
                });
#pragma unroll
                for (int m1 = 0; m1 < 16; ++m1) {
                    const float p = norm2(c[m1]);  // :307
                    best[m1] = fmaxf(best[m1], p);  // == the strict '>' update of :308 for the power itself
                }
            }
        } else {
            for (int k = 0; k < K; ++k) {
#ifdef B200_WHATIF_H0
                const float2* h = hperm + tid;
#else
                const float2* h = hperm + (size_t)k * 16 * kGroupThreads + tid;
#endif
                float2 y[16], c[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) y[j] = cmul(xs[j], __ldg(h + j * kGroupThreads));  // :247-249
                fft_b<kCorrTwoBuf, kRegB1, kRegB2>(y, c, tw_s, xb, tid, bar_id, xb2, t1, t2);   // :250-251
#pragma unroll
                for (int m1 = 0; m1 < 16; ++m1) {
                    const float p = norm2(c[m1]);  // :307
                    best[m1] = fmaxf(best[m1], p);  // == the strict '>' update of :308 for the power itself
                }
            }
        }
        // time reversal: lag kk lives at index (F - kk) mod F (:300)
        float* zdst = zpow + (s0 - z_base);
#pragma unroll
        for (int m1 = 0; m1 < 16; ++m1) {
            const int m = 128 * m1 + tid;
            const int kk = (kFft - m) & (kFft - 1);
            if (kk < S) zdst[kk] = best[m1];
        }
    }
    if constexpr (kCorrTmem) {
        tmem_fence_before_sync();
        __syncthreads();
        if (threadIdx.x < 32) tmem_dealloc(tm_slot, kTmCols);
    }
}

// ---------------------------------------------------------------------------------
// Refine: recompute, for each detected sample p, the HistoryItem fields of
// PM/syncword_detection.hpp:326-342 with exactly the arithmetic of correlate_kernel.
// One CTA per detection at a time, grid-stride over the device-side detection count.
//
// Only ONE output sample of each hypothesis' FFT B is needed (index m = (F - lag) mod F), so the
// transform is pruned to the cone of that output, using the same butterflies in the same order
// (bit-identical to the full transform):
//   pass B1  all 256 DFT-8 columns are needed, but only output m3 of each  -> 256 values
//   pass B2  only the 16 threads (m3, f1 = 0..15), only output m2          ->  16 values
//   pass B3  only thread (m2, m3), only output m1                          ->   1 value
// B1 runs on the 128 FFT threads for kRefineChunk hypotheses back to back; B2 and B3 of the whole
// chunk then run side by side (16 resp. 1 thread per hypothesis): about 1/3 of the arithmetic and
// 1/8 of the shared-memory traffic of full transforms, and two short tails per chunk instead of K.
// ---------------------------------------------------------------------------------
constexpr int kRefineThreads = kGroupThreads + 32;  // 4 FFT warps + 1 warp for the sequential noise sum
constexpr int kRefineChunk = kRefineThreads / 16;   // hypotheses whose B2 passes fit side by side (10)

template <int N>
__device__ __forceinline__ float2 pick(const float2 (&v)[N], int idx) {
    float2 r = v[0];
#pragma unroll
    for (int i = 1; i < N; ++i) r = (idx == i) ? v[i] : r;
    return r;
}

__global__ void __launch_bounds__(kRefineThreads)
refine_kernel(const float2* __restrict__ in, long long in_base, const float* __restrict__ zpow,
              long long z_base, const float2* __restrict__ hperm, int K, int S, int min_freq_bin,
              const float2* __restrict__ tw_g, const unsigned long long* __restrict__ det_idx,
              const unsigned int* __restrict__ det_count, unsigned int det_cap,
              DetectionRecord* __restrict__ recs) {
    extern __shared__ __align__(128) unsigned char smem_raw[];  // 128 B: half-warp LDS.64 rows never straddle a bank row, whatever static shared data precedes
    float2* tw_s = reinterpret_cast<float2*>(smem_raw);
    float2* xb = tw_s + kTwTotal;
    float2* b1out = xb + kXchgFloat2;                      // [kRefineChunk][256]
    float2* b2out = b1out + kRefineChunk * 256;            // [kRefineChunk][16]
    float2* corr_s = b2out + kRefineChunk * 16;            // [kMaxHyp + 1]
    float* xpow = reinterpret_cast<float*>(corr_s + kMaxHyp + 1);  // [2048]
    __shared__ float noise_s;
    unsigned int n = *det_count;
    if (n > det_cap) n = det_cap;
    if (blockIdx.x >= n) return;  // the grid is sized for the worst case; idle CTAs leave at once
    load_twiddles(tw_s, tw_g);
    __syncthreads();
    const int tid = threadIdx.x;
    const bool fft_warp = tid < kGroupThreads;
    for (unsigned int d = blockIdx.x; d < n; d += gridDim.x) {
        const long long p = (long long)det_idx[d];
        const long long b = p / S;
        const int kk = (int)(p - b * S);
        const int m = (kFft - kk) & (kFft - 1);            // output index of FFT B, m = 128 m1 + 8 m2 + m3
        const int m1 = m >> 7, m2 = (m >> 3) & 15, m3 = m & 7;
        const long long s0 = b * (long long)S;
        float2 xs[16];
        if (fft_warp) {
            const float2* src = in + (s0 - in_base);
            float2 v[16];
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) v[n1] = __ldcs(src + 128 * n1 + tid);
            fft_a(v, xs, tw_s, xb, tid, 1);
#pragma unroll
            for (int j = 0; j < 16; ++j) xpow[(tid + 128 * (j >> 3)) + 256 * (j & 7)] = norm2(xs[j]);
        }
        __syncthreads();
        for (int c0 = 0; c0 < K; c0 += kRefineChunk) {
            const int nk = min(kRefineChunk, K - c0);
            if (fft_warp) {
                // pass B1 of every hypothesis of the chunk: column p keeps only its output m3
                for (int kq = 0; kq < nk; ++kq) {
                    const float2* h = hperm + (size_t)(c0 + kq) * 16 * kGroupThreads + tid;
#pragma unroll
                    for (int pi = 0; pi < 2; ++pi) {
                        float2 w[8];
#pragma unroll
                        for (int dd = 0; dd < 8; ++dd)
                            w[dd] = cmul(xs[pi * 8 + dd], __ldg(h + (pi * 8 + dd) * kGroupThreads));
                        dft8(w);
                        float2 val = w[0];  // output m3 sits at w[bitrev3(m3)]
#pragma unroll
                        for (int q = 1; q < 8; ++q) val = (m3 == q) ? w[bitrev3(q)] : val;
                        const int pp = tid + 128 * pi;
                        if (m3 != 0) val = cmul(val, tw_s[kTwL + m3 * 256 + pp]);
                        b1out[kq * 256 + pp] = val;
                    }
                }
            } else if (c0 == 0 && tid == kGroupThreads) {
                // noise power: sequential float sum over k = F/4 .. 3F/4-1 in index order (:257-265),
                // overlapped with the first chunk's B1 passes of the other four warps
                float acc = 0.0f;
                for (int f = kFft / 4; f < 3 * kFft / 4; f += 16) {
                    float t[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) t[u] = xpow[f + u];
#pragma unroll
                    for (int u = 0; u < 16; ++u) acc = __fadd_rn(acc, t[u]);
                }
                noise_s = __fdiv_rn(acc, __fmul_rn((float)(kFft / 2), (float)kFft));
            }
            __syncthreads();
            // pass B2: thread (kq, f1) transforms over f2 and keeps output m2
            if (tid < 16 * nk) {
                const int kq = tid >> 4, f1 = tid & 15;
                float2 c[16];
#pragma unroll
                for (int f2 = 0; f2 < 16; ++f2) c[f2] = b1out[kq * 256 + f1 + 16 * f2];
                dft16(c);
                float2 val = c[0];
#pragma unroll
                for (int q = 1; q < 16; ++q) val = (m2 == q) ? c[bitrev4(q)] : val;
                if (m2 != 0) val = cmul(val, tw_s[kTwS + m2 * 16 + f1]);
                b2out[kq * 16 + f1] = val;
            }
            __syncthreads();
            // pass B3: one thread per hypothesis transforms over f1 and keeps output m1
            if (tid < nk) {
                float2 y[16];
#pragma unroll
                for (int f1 = 0; f1 < 16; ++f1) y[f1] = b2out[tid * 16 + f1];
                dft16(y);
                float2 val = y[0];
#pragma unroll
                for (int q = 1; q < 16; ++q) val = (m1 == q) ? y[bitrev4(q)] : val;
                corr_s[c0 + tid] = val;
            }
            // (b1out / b2out are rewritten only after the next chunk's first barrier)
        }
        __syncthreads();
        if (tid == 0) {
            int best_freq = 0;  // :301-313
            float2 z = make_float2(0.f, 0.f);
            float zp = -1.0f;
            for (int k = 0; k < K; ++k) {
                const float q = norm2(corr_s[k]);
                if (q > zp) { best_freq = k; z = corr_s[k]; zp = q; }
            }
            DetectionRecord r;
            r.index = (unsigned long long)p;
            r.corr_re = z.x;
            r.corr_im = z.y;
            r.pow = zp;
            r.pow_left = best_freq > 0 ? norm2(corr_s[best_freq - 1]) : 0.0f;
            r.pow_right = best_freq < K - 1 ? norm2(corr_s[best_freq + 1]) : 0.0f;
            r.pow_prev = (p - 1 >= 0 && p - 1 >= z_base) ? zpow[p - 1 - z_base] : 0.0f;
            r.pow_next = zpow[p + 1 - z_base];
            r.noise_power = noise_s;
            r.freq_bin = min_freq_bin + best_freq;
            r._pad = 0;
            recs[d] = r;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------
cudaError_t launch_template_spectra(const float2* d_td, float2* d_hperm, int K, const float2* d_tw,
                                    cudaStream_t st) {
    const size_t smem = sizeof(float2) * (size_t)(kTwTotal + kXchgFloat2);
    cudaError_t e = cudaFuncSetAttribute(template_spectra_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    template_spectra_kernel<<<K, kGroupThreads, smem, st>>>(d_td, d_hperm, d_tw);
    count_launch();
    return cudaGetLastError();
}

size_t correlate_smem_bytes(int groups) {
    return sizeof(float2) * (size_t)(kTwTotal + groups * kCorrXchg);
}

cudaError_t launch_correlate(const float2* d_in, long long in_base, float* d_zpow, long long z_base,
                             const float2* d_hperm, int K, int S, long long b0, long long nb,
                             const float2* d_tw, float2* d_out_delayed, long long out_base, long long out_lo,
                             long long out_hi, int delay, int num_sms, cudaStream_t st) {
    if (nb <= 0) return cudaSuccess;
    // function attributes are per device: one flag per ordinal (a process may hold contexts on several GPUs)
    static bool attr_set[64] = {};
    const int groups = kCorrThreads / kGroupThreads;
    const size_t smem = correlate_smem_bytes(groups);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        e = cudaFuncSetAttribute(correlate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    long long want = (nb + groups - 1) / groups;
    int grid = (int)(want < num_sms ? want : num_sms);
    correlate_kernel<<<grid, kCorrThreads, smem, st>>>(d_in, in_base, d_zpow, z_base, d_hperm, K, S, b0,
                                                        nb, d_tw, d_out_delayed, out_base, out_lo, out_hi, delay);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_refine(const float2* d_in, long long in_base, const float* d_zpow, long long z_base,
                          const float2* d_hperm, int K, int S, int min_freq_bin, const float2* d_tw,
                          const unsigned long long* d_det_idx, const unsigned int* d_det_count,
                          unsigned int det_cap, DetectionRecord* d_recs, int num_sms, cudaStream_t st) {
    const size_t smem = sizeof(float2) * (size_t)(kTwTotal + kXchgFloat2 + kRefineChunk * (256 + 16) + kMaxHyp + 1) +
                        sizeof(float) * kFft;
    cudaError_t e = cudaFuncSetAttribute(refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // exactly one resident wave (shared memory allows 3 CTAs per SM): a grid of 4 per SM ran a second,
    // one-third-full wave that took as long as the first
    static int per_sm = 0;
    if (per_sm == 0) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, refine_kernel, kRefineThreads, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
    }
    int grid = num_sms * per_sm;
    if ((unsigned)grid > det_cap) grid = (int)det_cap;
    if (grid < 1) grid = 1;
    refine_kernel<<<grid, kRefineThreads, smem, st>>>(d_in, in_base, d_zpow, z_base, d_hperm, K, S,
                                                  min_freq_bin, d_tw, d_det_idx, d_det_count, det_cap,
                                                  d_recs);
    count_launch();
    return cudaGetLastError();
}

}  // namespace b200sync

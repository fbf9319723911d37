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
#include "peak_walk.cuh"

#include <cstdlib>

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

// SPLIT: all groups of a CTA share one block and split the hypotheses (low latency, few blocks);
// BATCH: channel x block grid of the batched channel mode.  Separate instantiations keep the persistent
// single-stream shape free of their registers.
// ZC (with SPLIT): the span is read from mapped host memory (in_host), see the block loop.
template <bool SPLIT, bool BATCH, bool ZC = false>
__global__ void __launch_bounds__(kCorrThreads, 1)
correlate_kernel(const float2* __restrict__ in, long long in_base, float* __restrict__ zpow,
                 long long z_base, const float2* __restrict__ hperm, int K, int S, long long b0,
                 long long nb, const float2* __restrict__ tw_g, float2* __restrict__ out_delayed,
                 long long out_base, long long out_lo, long long out_hi, int delay, long long nb_chan,
                 long long in_chan_stride, long long z_chan_stride, float2* __restrict__ gm, long long gm_b0,
                 int gm_ng, long long gm_chan_stride, const float2* __restrict__ in_host, long long in_host_base,
                 float2* __restrict__ x_copy) {
    extern __shared__ __align__(128) unsigned char smem_raw[];  // 128 B: half-warp LDS.64 rows never straddle a bank row, whatever static shared data precedes
    pdl_launch_dependents();   // streaming: the flags launch may become resident (and wait) behind this grid
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
        if (K <= kTmHypA && g == (ngr > 1 ? ngr - 2 : 0)) {
            // few hypotheses: FFT A's inter-pass factors take the two free template slots (fft_a_tm)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int k1 = 1 + 8 * hf + e;
                    r[e] = k1 < 16 ? tw_s[kTwS + k1 * 16 + (tid >> 3)] : make_float2(0.f, 0.f);
                }
                tmem_st8(tm_base + kTmA1 + 16 * hf, r);
#pragma unroll
                for (int e = 0; e < 8; ++e) r[e] = tw_s[kTwL + (tid >> 4) * 256 + (tid & 15) + 16 * (8 * hf + e)];
                tmem_st8(tm_base + kTmA2 + 16 * hf, r);
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

    // ksplit > 1 (few blocks, e.g. one processBulk span of the streaming path): ALL groups of the CTA work on the
    // same block — each transforms it (FFT A, redundantly) and takes every ngroups-th hypothesis — and the
    // per-sample maxima are merged through shared memory.  A block's latency drops from 1 + K transforms to
    // 1 + ceil(K / ngroups); max() commutes, so the metric is bit-identical.
    constexpr bool split = SPLIT;
    const long long blk_first = split ? (long long)blockIdx.x : (long long)blockIdx.x * ngroups + g;
    const long long blk_step = split ? (long long)gridDim.x : gstride;
    const int k_first = split ? g : 0, k_step = split ? ngroups : 1;
    for (long long blk = blk_first; blk < nb; blk += blk_step) {
        // batched channel mode (nb_chan > 0): nb = channels x nb_chan blocks, channel c's stream and metric lie
        // c * stride after the first channel's; every channel has its own block grid starting at b0
        long long bb = blk, ch = 0;
        if constexpr (BATCH) {
            ch = blk / nb_chan;
            bb = blk - ch * nb_chan;
        }
        const long long s0 = (b0 + bb) * (long long)S;  // absolute first sample of the block
        const float2* src = in + ch * in_chan_stride + (s0 - in_base);
        float2 v[16], xs[16];
        if constexpr (SPLIT && ZC) {
            // Streaming, zero copy: the span still lies in (pinned, mapped) HOST memory.  Group 0 pulls the block
            // over PCIe itself — no separate H2D copy in front of the kernel, and the blocks of the span arrive
            // side by side — leaves it in the device window x_copy (same indexing as `in`: the refine stage and
            // the next call read it there), and hands it to the other groups through its exchange buffer.
            float2* stage = tw_s + kTwTotal;     // group 0's exchange buffer, free until FFT A
            if (g == 0) {
                const float2* hsrc = in_host + (s0 - in_host_base) + tid;
                float2* dcopy = x_copy + (s0 - in_base) + tid;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {   // two batches of eight rows: short live ranges (the kernel sits at its register cap)
                    float2 t[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) t[r] = __ldcs(hsrc + 128 * (8 * hf + r));
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        dcopy[128 * (8 * hf + r)] = t[r];
                        stage[128 * (8 * hf + r) + tid] = t[r];
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) v[n1] = stage[128 * n1 + tid];
            __syncthreads();
        } else {
        // The first and last L - 1 samples of a block are also the neighbouring blocks' (overlap-save): those rows are
        // loaded with the default policy so that the second reader — another group of this CTA, microseconds
        // later — finds them in L2; the rows in between are read once and stream through (evict-first).
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
#ifdef B200_NO_OVERLAP_L2
            v[n1] = __ldcs(src + 128 * n1 + tid);
#else
            // (rows 0-2 and 13-15 cover the overlap of the default syncword, L = 297; a pure cache hint for any other)
            constexpr bool kSharedRow[16] = {1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1};
            v[n1] = kSharedRow[n1] ? __ldg(src + 128 * n1 + tid) : __ldcs(src + 128 * n1 + tid);
#endif
        }
        }
#ifdef B200_BULK_PREFETCH   // measured SLOWER as well (K = 1: 4.68 vs 4.47 ms, K = 9: 18.7 vs 18.3): off by default
        if constexpr (!SPLIT && !BATCH) {
            // this group's NEXT block on its way into L2 while the current one is transformed: ONE bulk prefetch
            // (cp.async.bulk.prefetch.L2, the TMA engine walks the 16 KiB) issued by one thread of the group — with
            // few hypotheses a block lasts about one DRAM latency and only ~2.5 of the 6 groups of an SM have
            // loads in flight at any time.  (Sixteen per-line prefetch instructions per 16 lanes cost registers
            // and were slower: B200_L2_PREFETCH below.)
            if (tid == 0 && blk + blk_step < nb) {
                const float2* nsrc = in + ((b0 + blk + blk_step) * (long long)S - in_base);
                if ((reinterpret_cast<uintptr_t>(nsrc) & 15) == 0)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nsrc), "r"((int)(kFft * sizeof(float2))) : "memory");
            }
        }
#endif
#ifdef B200_L2_PREFETCH   // measured SLOWER (K = 1: 4.82 vs 4.47 ms at 2^30; K = 9: 18.8 vs 18.3): register pressure, off by default
        if (!split && (tid & 15) == 0 && blk + blk_step < nb) {
            // this group's NEXT block on its way into L2 while the current one is transformed (one request per
            // 128-byte line): with few hypotheses a block lasts about one DRAM latency
            long long bbn = blk + blk_step, chn = 0;
            if constexpr (BATCH) {
                chn = bbn / nb_chan;
                bbn -= chn * nb_chan;
            }
            const float2* nsrc = in + chn * in_chan_stride + ((b0 + bbn) * (long long)S - in_base) + tid;
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) asm volatile("prefetch.global.L2 [%0];" ::"l"(nsrc + 128 * n1));
        }
#endif
        if (out_delayed != nullptr && (!split || g == 0)) {
            // block contract: out[n] = in[n - delay] (PM/syncword_detection.hpp:318-319).  Only output items
            // [out_lo, out_hi) are stored: out_hi = the number of items the call publishes (:346) — nothing is
            // stored at or past it — and a time shard stores only the slice it owns.
            // This block owns samples [s0, s0+S); it has them in registers already.
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) {
                const int i = 128 * n1 + tid;
                const long long o = s0 + i + delay;
                if (i < S && o >= out_lo && o < out_hi) __stcs(out_delayed + (o - out_base), v[n1]);   // written once, read by a later stage
            }
        }
        if constexpr (kCorrTwoBuf) group_sync(bar_id);  // previous block's last reads of xb2 are done
        if (kCorrTmem && !kCorrTwoBuf && K <= kTmHypA) fft_a_tm(v, xs, xb, tid, bar_id, tm_base);   // uniform over the launch
        else fft_a<kCorrTwoBuf>(v, xs, tw_s, xb, tid, bar_id, xb2);

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
            for (int k = k_first; k < K; k += k_step) {
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
            for (int k = k_first; k < K; k += k_step) {
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
        if (split) {   // merge the groups' maxima: every group parks its 16 values in its own exchange buffer
            float* mine = reinterpret_cast<float*>(xb);
#pragma unroll
            for (int m1 = 0; m1 < 16; ++m1) mine[m1 * kGroupThreads + tid] = best[m1];
            __syncthreads();
            if (g == 0) {
                for (int gg = 1; gg < ngroups; ++gg) {
                    const float* other = reinterpret_cast<const float*>(tw_s + kTwTotal + gg * kCorrXchg);
#pragma unroll
                    for (int m1 = 0; m1 < 16; ++m1) best[m1] = fmaxf(best[m1], other[m1 * kGroupThreads + tid]);
                }
            }
        }
        // time reversal: lag kk lives at index (F - kk) mod F (:300)
        if (!split || g == 0) {
            float* zdst = zpow + ch * z_chan_stride + (s0 - z_base);
#pragma unroll
            for (int m1 = 0; m1 < 16; ++m1) {
                const int m = 128 * m1 + tid;
                const int kk = (kFft - m) & (kFft - 1);
                if (kk < S) __stcs(zdst + kk, best[m1]);
            }
#ifdef B200_GM_FLAGS   // measured dead end, see gm_supported() in b200sync_internal.h
            if (gm != nullptr) {
                // Group extrema for the peak stage (peaks.cu: peak_flags_gm_kernel): a warp holds 32 CONSECUTIVE
                // samples per m1 — lags 32q-31 .. 32q with q = 64 - 4 m1 - warp — so REDUX gives the maximum of that
                // group of the metric and three butterfly shuffles the minimum of each of its four 8-sample
                // quarters; lag 0 (thread 0, m1 = 0) is group 0 on its own.  The peak stage then reads 20 B per 32
                // samples instead of every sample.  (zpow >= +0: unsigned bit order.)
                // Row layout: [max, -, -, -, min of lags 32q-31.., min of 32q-23.., min of 32q-15.., min of 32q-7..]
                float* grow = reinterpret_cast<float*>(gm + ch * gm_chan_stride + ((b0 + bb) - gm_b0) * gm_ng * kGmF2PerGroup);
                const int wq = tid >> 5, lane = tid & 31;
#pragma unroll
                for (int m1 = 0; m1 < 16; ++m1) {
                    const int m = 128 * m1 + tid;
                    const int kk = (kFft - m) & (kFft - 1);
                    const bool valid = kk < S;
                    const unsigned bits = __float_as_uint(best[m1]);
                    const unsigned mx = __reduce_max_sync(0xffffffffu, valid ? bits : 0u);
                    unsigned mn = valid ? bits : 0x7f800000u;   // (a REDUX with quarter-warp masks measured 3x slower)
                    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, 1));
                    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, 2));
                    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, 4));
                    const int q = (m1 == 0 && wq == 0) ? 0 : 64 - 4 * m1 - wq;
                    if (q < gm_ng) {
                        if (lane == 0) grow[8 * q] = __uint_as_float(mx);
                        if ((lane & 7) == 0) grow[8 * q + 4 + (3 - (lane >> 3))] = __uint_as_float(mn);   // ascending lag order
                    }
                }
            }
#endif
        }
        if (split) __syncthreads();   // the exchange buffers are free again
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

// STREAM: the in-order peak walk fused in front (streaming path); its own instantiation, so the bulk form carries none of it
template <bool STREAM>
__global__ void __maxnreg__(88)   // 160 threads: 3 CTAs per SM (what shared memory allows); ptxas 12.9 otherwise wanders between 88 and 164 registers with the surrounding code
refine_kernel(const float2* __restrict__ in, long long in_base, const float* __restrict__ zpow,
              long long z_base, const float2* __restrict__ hperm, int K, int S, int min_freq_bin,
              const float2* __restrict__ tw_g, const unsigned long long* __restrict__ det_idx,
              const unsigned int* __restrict__ det_count, unsigned int det_cap,
              DetectionRecord* __restrict__ recs, long long in_chan_stride, long long z_chan_stride,
              long long det_chan_stride, StreamWalk walk) {
    extern __shared__ __align__(128) unsigned char smem_raw[];  // 128 B: half-warp LDS.64 rows never straddle a bank row, whatever static shared data precedes
    {   // batched channel mode: blockIdx.y is the channel; det_count points into an array of PeakState
        const long long ch = blockIdx.y;
        in += ch * in_chan_stride;
        zpow += ch * z_chan_stride;
        det_idx += ch * det_chan_stride;
        recs += ch * det_chan_stride;
        det_count += ch * (long long)(sizeof(PeakState) / sizeof(unsigned int));
    }
    float2* tw_s = reinterpret_cast<float2*>(smem_raw);
    float2* xb = tw_s + kTwTotal;
    // (sized by the launch for THIS K — refine_smem_float2(K) — not for the largest one: with few hypotheses a fourth
    //  CTA fits on the SM, and the stage is latency-bound, a CTA per detection)
    const int chunk_alloc = K < kRefineChunk ? K : kRefineChunk;
    float2* b1out = xb + kXchgFloat2;                      // [chunk_alloc][256]
    float2* b2out = b1out + chunk_alloc * 256;             // [chunk_alloc][16]
    float2* corr_s = b2out + chunk_alloc * 16;             // [K + 1]
    float* xpow = reinterpret_cast<float*>(corr_s + K + 1 + ((K + 1) & 1));  // [2048] (16-byte aligned)
    __shared__ float noise_s;
    __shared__ unsigned int n_walk;
    unsigned int n;
    const unsigned long long* dets = det_idx;
    if constexpr (STREAM) {
        // launched behind the flags kernel (launch_pdl): the twiddle table — a constant of the context — is staged
        // while that kernel still runs; everything below reads what it and the correlator wrote
        load_twiddles(tw_s, tw_g);
        pdl_wait();
        // Streaming: the in-order walk of the peak detector fused in front of the refine stage.  EVERY CTA walks
        // the (small) bitmaps itself, from shared memory, and so knows the whole detection list without a
        // kernel boundary; CTA d then refines detection d.  CTA 0 publishes the new search position and the count.
        unsigned long long* det_s = reinterpret_cast<unsigned long long*>(xpow + kFft);   // [det_cap]
        uint32_t* cand_s = reinterpret_cast<uint32_t*>(det_s + det_cap);
        const int nwords = (int)((walk.range + 31) >> 5);
        uint32_t* pass_s = cand_s + nwords;
        for (int i = threadIdx.x; i < nwords; i += kRefineThreads) {
            cand_s[i] = walk.cand[i];
            pass_s[i] = walk.pass[i];
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            const long long j0 = (walk.r_abs_in > (unsigned long long)walk.lo) ? (long long)(walk.r_abs_in - (unsigned long long)walk.lo) : 0;
            unsigned int cnt = 0;
            const long long j = peak_walk_warp(cand_s, pass_s, nwords, walk.range, walk.T, j0, [&](long long found) {
                if (threadIdx.x == 0 && cnt < det_cap) det_s[cnt] = (unsigned long long)(walk.lo + found);
                ++cnt;
            });
            if (threadIdx.x == 0) {
                n_walk = cnt;
                if (blockIdx.x == 0) {
                    const long long r_end = walk.lo + j;
                    PeakState st;
                    st.r_abs = (unsigned long long)(r_end > walk.hi ? r_end : walk.hi);
                    st.det_count = cnt;
                    st._pad = 0;
                    if (walk.done == nullptr) {
                        *walk.header = st;      // the host reads the count (and the records behind it) from here
                    } else {                    // mapped host memory: count and position now, the sequence number last
                        walk.header->r_abs = st.r_abs;
                        walk.header->det_count = cnt;
                    }
                    st.det_count = 0;           // the device-side list counts as drained
                    *walk.state_out = st;
                }
            }
        }
        __syncthreads();
        n = n_walk;
        dets = det_s;
    } else {
        n = *det_count;
    }
    if (n > det_cap) n = det_cap;
    auto count_done = [&]() {   // (thread 0 of the CTA) everything this CTA wrote is visible system-wide before it counts
        __threadfence_system();
        if (atomicAdd(walk.done, 1u) == gridDim.x - 1) {
            *walk.done = 0u;                                        // ready for the next launch (stream order)
            __threadfence_system();
            *reinterpret_cast<volatile unsigned int*>(&walk.header->_pad) = walk.seq;
        }
    };
    if (blockIdx.x >= n) {  // the grid is sized for the worst case; idle CTAs leave at once
        if (STREAM && walk.done != nullptr && threadIdx.x == 0) count_done();   // (its barrier above ordered CTA 0's header stores)
        return;
    }
    if constexpr (!STREAM) load_twiddles(tw_s, tw_g);
    __syncthreads();
    const int tid = threadIdx.x;
    const bool fft_warp = tid < kGroupThreads;
    for (unsigned int d = blockIdx.x; d < n; d += gridDim.x) {
        const long long p = (long long)dets[d];
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
    if (STREAM && walk.done != nullptr && threadIdx.x == 0) count_done();
}

// ---------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------
cudaError_t launch_template_spectra(const float2* d_td, float2* d_hperm, int K, const float2* d_tw,
                                    cudaStream_t st, int fft) {
    if (fft != kFft) return launch_template_spectra_generic(fft & ~kGenericFlag, d_td, d_hperm, K, d_tw, st);
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

// true when launch_correlate would take the split shape for this span (a CTA per block, all groups on the same
// block) — the shape that can pull its input from mapped host memory (in_host)
bool correlate_takes_host_input(int K, int fft, long long nb, int num_sms) {
    if (getenv("B200SYNC_SMALL_GROUPS") != nullptr) return false;   // development override of the shape
    return fft == kFft && nb > 0 && nb <= num_sms && K >= 2;
}

cudaError_t launch_correlate(const float2* d_in, long long in_base, float* d_zpow, long long z_base,
                             const float2* d_hperm, int K, int S, int fft, long long b0, long long nb,
                             const float2* d_tw, float2* d_out_delayed, long long out_base, long long out_lo,
                             long long out_hi, int delay, int num_sms, cudaStream_t st, long long nb_chan,
                             long long in_chan_stride, long long z_chan_stride, float2* d_gm, long long gm_b0,
                             long long gm_chan_stride, const float2* in_host, long long in_host_base) {
    if (nb <= 0) return cudaSuccess;
    if (fft != kFft)   // another fft_size, or 2048 forced onto the generic path (kGenericFlag)
        return launch_correlate_generic(fft & ~kGenericFlag, d_in, in_base, d_zpow, z_base, d_hperm, K, S, b0, nb, d_tw,
                                        d_out_delayed, out_base, out_lo, out_hi, delay, num_sms, st, nb_chan,
                                        in_chan_stride, z_chan_stride);
    // function attributes are per device: one flag per ordinal (a process may hold contexts on several GPUs)
    static bool attr_set[64] = {};
    const int max_groups = kCorrThreads / kGroupThreads;
    const size_t smem_max = correlate_smem_bytes(max_groups);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        e = cudaFuncSetAttribute(correlate_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(correlate_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(correlate_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(correlate_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    // Few blocks (one processBulk span of the streaming path is 37 blocks): what the caller waits for is the
    // LATENCY of a block.  Up to one block per SM: a CTA per block, its groups splitting the hypotheses (ksplit);
    // more: fewer groups per CTA so the blocks spread over all SMs; many blocks: the persistent 6-group shape.
    int groups = max_groups, ksplit = 1;
    static int forced = -1;  // development override: B200SYNC_SMALL_GROUPS=1..6
    if (forced < 0) {
        const char* v = getenv("B200SYNC_SMALL_GROUPS");
        forced = v ? atoi(v) : 0;
    }
    if (nb_chan == 0 && nb <= num_sms && K >= 2 && d_out_delayed == nullptr) {
        groups = K < max_groups ? K : max_groups;
        if (K > max_groups && K <= 2 * max_groups) groups = (K + 1) / 2;   // e.g. K = 9: 5 groups, two rounds at most
        if (forced > 0) groups = forced;
        if (groups > max_groups) groups = max_groups;
        ksplit = groups;
    } else if (nb < (long long)max_groups * num_sms) {
        groups = (int)((nb + num_sms - 1) / num_sms);
        if (forced > 0) groups = forced;
        if (groups < 1) groups = 1;
        if (groups > max_groups) groups = max_groups;
    }
    const size_t smem = correlate_smem_bytes(groups);
    long long want = ksplit > 1 ? nb : (nb + groups - 1) / groups;
    int grid = (int)(want < num_sms ? want : num_sms);
    auto kern = ksplit > 1 ? (in_host != nullptr ? correlate_kernel<true, false, true> : correlate_kernel<true, false>)
                           : (nb_chan > 0 ? correlate_kernel<false, true> : correlate_kernel<false, false>);
    kern<<<grid, groups * kGroupThreads, smem, st>>>(d_in, in_base, d_zpow, z_base, d_hperm, K, S, b0, nb, d_tw,
                                                     d_out_delayed, out_base, out_lo, out_hi, delay, nb_chan,
                                                     in_chan_stride, z_chan_stride, d_gm, gm_b0, gm_groups_per_block(S),
                                                     gm_chan_stride, ksplit > 1 ? in_host : nullptr, in_host_base,
                                                     const_cast<float2*>(d_in));
    if (in_host != nullptr && ksplit <= 1) return cudaErrorInvalidValue;   // callers ask correlate_takes_host_input() first
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_refine(const float2* d_in, long long in_base, const float* d_zpow, long long z_base,
                          const float2* d_hperm, int K, int S, int fft, int min_freq_bin, const float2* d_tw,
                          const unsigned long long* d_det_idx, const unsigned int* d_det_count,
                          unsigned int det_cap, DetectionRecord* d_recs, int num_sms, cudaStream_t st, int nch,
                          long long in_chan_stride, long long z_chan_stride, long long det_chan_stride,
                          const StreamWalk* walk) {
    if (fft != kFft) {
        if (walk != nullptr) return cudaErrorInvalidValue;   // the fused walk exists on the 2048 path only
        return launch_refine_generic(fft & ~kGenericFlag, d_in, in_base, d_zpow, z_base, d_hperm, K, S, min_freq_bin, d_tw,
                                     d_det_idx, d_det_count, det_cap, d_recs, num_sms, st, nch, in_chan_stride,
                                     z_chan_stride, det_chan_stride);
    }
    auto refine_smem = [](int k) {   // the kernel's carve-up: tables, exchange buffer, chunk outputs, K + 1 correlations, |X|^2
        const int chunk = k < kRefineChunk ? k : kRefineChunk;
        return sizeof(float2) * (size_t)(kTwTotal + kXchgFloat2 + chunk * (256 + 16) + k + 1 + ((k + 1) & 1)) + sizeof(float) * kFft;
    };
    const size_t smem_base = refine_smem(K);
    // the attribute is for the largest K and the largest streaming walk: 2^19-sample bitmaps + the list
    const size_t smem = refine_smem(kMaxHyp) + 2 * sizeof(uint32_t) * ((1u << 19) / 32 + 1) + sizeof(unsigned long long) * 1024;
    size_t smem_launch = smem_base;
    StreamWalk w{};
    if (walk != nullptr) {
        w = *walk;
        smem_launch = smem_base + 2 * sizeof(uint32_t) * (size_t)((w.range + 31) / 32 + 1) + sizeof(unsigned long long) * det_cap;
        if (smem_launch > smem || nch != 1) return cudaErrorInvalidValue;
    }
    static bool attr_set[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        e = cudaFuncSetAttribute(refine_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(refine_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    // exactly one resident wave (shared memory allows 3 CTAs per SM at K >= 7, 4 below): a grid of 4 per SM where 3 fit
    // ran a second, one-third-full wave that took as long as the first
    static int per_sm_k[kMaxHyp + 1] = {};   // resident CTAs per SM for K hypotheses (benign race: same value)
    int per_sm = per_sm_k[K];
    if (per_sm == 0) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, refine_kernel<false>, kRefineThreads, smem_base);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        per_sm_k[K] = per_sm;
    }
    int grid = num_sms * per_sm;
    if (nch > 1) grid = (grid + nch - 1) / nch;  // the channels share the one resident wave
    if ((unsigned)grid > det_cap) grid = (int)det_cap;
    if (grid < 1) grid = 1;
    if (walk != nullptr) {
        e = launch_pdl(refine_kernel<true>, dim3((unsigned)grid, (unsigned)nch), dim3(kRefineThreads), smem_launch, st, d_in, in_base,
                       d_zpow, z_base, d_hperm, K, S, min_freq_bin, d_tw, d_det_idx, d_det_count, det_cap, d_recs,
                       in_chan_stride, z_chan_stride, det_chan_stride, w);
        if (e != cudaSuccess) return e;
    } else {
        refine_kernel<false><<<dim3((unsigned)grid, (unsigned)nch), kRefineThreads, smem_launch, st>>>(
            d_in, in_base, d_zpow, z_base, d_hperm, K, S, min_freq_bin, d_tw, d_det_idx, d_det_count, det_cap, d_recs,
            in_chan_stride, z_chan_stride, det_chan_stride, w);
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace b200sync

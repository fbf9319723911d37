// correlator_generic.cu — the overlap-save syncword correlator for every fft_size OTHER than 2048.
//
// The reference accepts any power of two (PM/syncword_detection.hpp:133, "FFT size must be 2^N",
// ALG/fourier/fftw.hpp:182-184); correlator.cu is hand-scheduled for the default 2048 only.  This file covers
// 64 <= fft_size <= 8192 with the same three entry points (template spectra of start() :183-188, the block loop
// :236-252 + per-sample best hypothesis :299-313, and the record fields :257-265, :326-342 for detected samples),
// one CTA per FFT block, the transform in shared memory.
//
// Arithmetic contract.  Radix-2 decimation in time, every butterfly  (u, v) -> (u + w v, u - w v)  with
//   w v = (w.x v.x - w.y v.y,  w.x v.y + w.y v.x),   four products and two sums, each rounded on its own,
// w = (float cos a, float sin a), a = -2 pi j / s evaluated in double, products  x * conj_template  formed the same
// way and |z|^2 = z.x z.x + z.y z.y.  That is the std-only structure of the reference's own FFT
// (ALG/fourier/fft.hpp:70-101) and, bit for bit, the INDEPENDENT arithmetic the tests hold the 2048 kernel to for
// indices (oracle FftKind::Radix2, also the FFT under the reference's block code in oracle/_ref/librefblocks.so): on
// this path the metric, the detections and the records are bit-identical to that arithmetic, not merely index-exact.
// The butterflies run in Stockham (autosort) order, two stages per pass — the same butterfly graph on the same operands as the
// bit-reversal form, so the same roundings — which keeps every shared-memory access unit-stride or a two-way split:
//   stage s (sub-transform length m = 2^s, R = N / 2m), butterfly i = r + R k (r < R, k < m):
//       u = src[i + R k], v = src[i + R k + R], w = tw[m - 1 + k]  ->  dst[i] = u + w v, dst[i + N/2] = u - w v.
#include "b200sync_internal.h"
#include "peak_walk.cuh"

namespace b200sync {

namespace {

template <int LOGN>
struct Gen {
    static constexpr int N = 1 << LOGN;
    static constexpr int NT = N / 8 < 32 ? 32 : (N / 8 > 1024 ? 1024 : N / 8);   // threads per CTA
    static constexpr int EPT = N / NT;                                           // elements per thread (2 .. 8)
    static constexpr int BPT = N / 2 / NT;                                       // butterflies per thread and stage
    static constexpr size_t smem_fft = sizeof(float2) * 3 * (size_t)N;           // twiddles + two transform buffers
};

__device__ __forceinline__ float2 cmul_plain(float2 a, float2 b) {
    return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)),
                       __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float norm2_plain(float2 z) { return __fadd_rn(__fmul_rn(z.x, z.x), __fmul_rn(z.y, z.y)); }

// in: natural order in `src`, complete and visible (a barrier behind the last write).  Returns the buffer that holds
// the natural-order transform; a barrier has been passed after its last write.
// Two radix-2 stages per pass (the four butterflies of a radix-4 step, each with the radix-2 arithmetic above — the
// intermediate values never leave the registers), one plain radix-2 pass last when LOGN is odd: half the
// shared-memory traffic and barriers of stage-by-stage passes, the same roundings.
//   pass over stages s, s+1 (m = 2^s, R' = N / 4m), unit q = r' + R' k (r' < R', k < m), base = q + 3 R' k:
//     x_j = src[base + j R'];  stage s: a = x0 +- w x2, b = x1 +- w x3, w = tw[m-1+k];
//     stage s+1: y0, y2 = a0 +- w1 b0, w1 = tw[2m-1+k];  y1, y3 = a1 +- w2 b1, w2 = tw[2m-1+k+m];  dst[q + j N/4] = y_j
template <int LOGN>
__device__ __forceinline__ float2* fft_r2(float2* src, float2* dst, const float2* __restrict__ tw_s, int tid) {
    using G = Gen<LOGN>;
    auto bfly = [](float2 u, float2 v, float2 w, float2& p, float2& q) {
        const float2 t = cmul_plain(w, v);
        p = make_float2(__fadd_rn(u.x, t.x), __fadd_rn(u.y, t.y));
        q = make_float2(__fsub_rn(u.x, t.x), __fsub_rn(u.y, t.y));
    };
    int s = 0;
#pragma unroll 1
    for (; s + 1 < LOGN; s += 2) {
        const int m = 1 << s, lr = LOGN - 2 - s, Rp = 1 << lr;
        constexpr int UPT = G::N / 4 / G::NT > 0 ? G::N / 4 / G::NT : 1;   // units per thread
#pragma unroll
        for (int i = 0; i < UPT; ++i) {
            const int q = tid + i * G::NT;
            if (G::N / 4 >= G::NT || q < G::N / 4) {
                const int k = q >> lr;
                const float2* x = src + q + 3 * Rp * k;
                float2 x0, x1, x2, x3;
                if (lr == 0) {   // last pass of an even log2 N: the four inputs are contiguous — two 128-bit loads
                    const float4 lo4 = reinterpret_cast<const float4*>(x)[0], hi4 = reinterpret_cast<const float4*>(x)[1];
                    x0 = make_float2(lo4.x, lo4.y);
                    x1 = make_float2(lo4.z, lo4.w);
                    x2 = make_float2(hi4.x, hi4.y);
                    x3 = make_float2(hi4.z, hi4.w);
                } else {
                    x0 = x[0];
                    x1 = x[Rp];
                    x2 = x[2 * Rp];
                    x3 = x[3 * Rp];
                }
                const float2 w = tw_s[m - 1 + k], w1 = tw_s[2 * m - 1 + k], w2 = tw_s[3 * m - 1 + k];
                float2 a0, a1, b0, b1, y0, y1, y2, y3;
                bfly(x0, x2, w, a0, a1);
                bfly(x1, x3, w, b0, b1);
                bfly(a0, b0, w1, y0, y2);
                bfly(a1, b1, w2, y1, y3);
                dst[q] = y0;
                dst[q + G::N / 4] = y1;
                dst[q + G::N / 2] = y2;
                dst[q + 3 * G::N / 4] = y3;
            }
        }
        __syncthreads();
        float2* t = src;
        src = dst;
        dst = t;
    }
    if (s < LOGN) {   // LOGN odd: the last stage on its own, m = N/2, R = 1
        constexpr int m = G::N / 2;
#pragma unroll
        for (int i = 0; i < G::BPT; ++i) {
            const int bf = tid + i * G::NT;   // = k
            float2 p, q;
            const float4 uv = reinterpret_cast<const float4*>(src)[bf];   // (src[2 bf], src[2 bf + 1]) as one 128-bit load
            bfly(make_float2(uv.x, uv.y), make_float2(uv.z, uv.w), tw_s[m - 1 + bf], p, q);
            dst[bf] = p;
            dst[bf + G::N / 2] = q;
        }
        __syncthreads();
        float2* t = src;
        src = dst;
        dst = t;
    }
    return src;
}

template <int LOGN>
__device__ __forceinline__ void load_tw(float2* tw_s, const float2* __restrict__ tw_g, int tid) {
    using G = Gen<LOGN>;
    for (int i = tid; i < G::N - 1; i += G::NT) tw_s[i] = tw_g[i];
}

// conj(FFT(zero-padded shifted syncword k)) in natural order: hc[k][f]   (:183-188)
template <int LOGN>
__global__ void __launch_bounds__(Gen<LOGN>::NT)
template_spectra_generic_kernel(const float2* __restrict__ td, float2* __restrict__ hc, const float2* __restrict__ tw_g) {
    using G = Gen<LOGN>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw_s = reinterpret_cast<float2*>(smem_raw);
    float2* a = tw_s + G::N;
    float2* b = a + G::N;
    const int tid = threadIdx.x;
    load_tw<LOGN>(tw_s, tw_g, tid);
    const float2* src = td + (size_t)blockIdx.x * G::N;
#pragma unroll
    for (int j = 0; j < G::EPT; ++j) a[tid + j * G::NT] = src[tid + j * G::NT];
    __syncthreads();
    const float2* y = fft_r2<LOGN>(a, b, tw_s, tid);
#pragma unroll
    for (int j = 0; j < G::EPT; ++j) {
        const float2 z = y[tid + j * G::NT];
        hc[(size_t)blockIdx.x * G::N + tid + j * G::NT] = make_float2(z.x, -z.y);
    }
}

// Persistent: CTA c handles blocks c, c + gridDim.x, ... of [b0, b0 + nb) (nb = channels x nb_chan in batched
// channel mode).  Block b covers absolute samples [b S, b S + N) and produces zpow for [b S, (b+1) S).
template <int LOGN>
__global__ void __launch_bounds__(Gen<LOGN>::NT)
correlate_generic_kernel(const float2* __restrict__ in, long long in_base, float* __restrict__ zpow, long long z_base,
                         const float2* __restrict__ hc, int K, int S, long long b0, long long nb,
                         const float2* __restrict__ tw_g, float2* __restrict__ out_delayed, long long out_base,
                         long long out_lo, long long out_hi, int delay, long long nb_chan, long long in_chan_stride,
                         long long z_chan_stride) {
    using G = Gen<LOGN>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw_s = reinterpret_cast<float2*>(smem_raw);
    float2* a = tw_s + G::N;
    float2* b = a + G::N;
    const int tid = threadIdx.x;
    load_tw<LOGN>(tw_s, tw_g, tid);
    for (long long blk = blockIdx.x; blk < nb; blk += gridDim.x) {
        long long bb = blk, ch = 0;
        if (nb_chan > 0) {
            ch = blk / nb_chan;
            bb = blk - ch * nb_chan;
        }
        const long long s0 = (b0 + bb) * (long long)S;   // absolute first sample of the block
        const float2* src = in + ch * in_chan_stride + (s0 - in_base);
        float2 xs[G::EPT];
#pragma unroll
        for (int j = 0; j < G::EPT; ++j) xs[j] = __ldcs(src + tid + j * G::NT);
#pragma unroll
        for (int j = 0; j < G::EPT; ++j) a[tid + j * G::NT] = xs[j];   // (own elements: the previous block's reads were the owner's)
        if (out_delayed != nullptr) {
            // block contract: out[n] = in[n - delay] (:318-319), stored only inside [out_lo, out_hi): nothing at or
            // past the publish limit (:346), and a time shard stores only the slice it owns
#pragma unroll
            for (int j = 0; j < G::EPT; ++j) {
                const int i = tid + j * G::NT;
                const long long o = s0 + i + delay;
                if (i < S && o >= out_lo && o < out_hi) out_delayed[o - out_base] = xs[j];
            }
        }
        __syncthreads();
        {
            const float2* y = fft_r2<LOGN>(a, b, tw_s, tid);            // :239-241
#pragma unroll
            for (int j = 0; j < G::EPT; ++j) xs[j] = y[tid + j * G::NT];   // the spectrum stays in registers for all K
        }
        float best[G::EPT];
#pragma unroll
        for (int j = 0; j < G::EPT; ++j) best[j] = -1.0f;               // :303
        for (int k = 0; k < K; ++k) {
            const float2* h = hc + (size_t)k * G::N + tid;
#pragma unroll
            for (int j = 0; j < G::EPT; ++j) a[tid + j * G::NT] = cmul_plain(xs[j], __ldg(h + j * G::NT));   // :247-249
            __syncthreads();
            const float2* y = fft_r2<LOGN>(a, b, tw_s, tid);            // :250-251
#pragma unroll
            for (int j = 0; j < G::EPT; ++j)
                best[j] = fmaxf(best[j], norm2_plain(y[tid + j * G::NT]));   // :307-308, the power itself
        }
        // time reversal: lag kk lives at index (N - kk) mod N (:300)
        float* zdst = zpow + ch * z_chan_stride + (s0 - z_base);
#pragma unroll
        for (int j = 0; j < G::EPT; ++j) {
            const int kk = (G::N - (tid + j * G::NT)) & (G::N - 1);
            if (kk < S) zdst[kk] = best[j];
        }
    }
}

// Record fields of the detected samples (:257-265 noise power, :301-313 winner and neighbours, :326-342), with the
// arithmetic of correlate_generic_kernel.  One CTA per detection at a time; full transforms (detections are sparse).
// The in-order peak walk of the streaming path is not fused here (api.cu takes the two-launch walk for this path).
template <int LOGN>
__global__ void __launch_bounds__(Gen<LOGN>::NT)
refine_generic_kernel(const float2* __restrict__ in, long long in_base, const float* __restrict__ zpow, long long z_base,
                      const float2* __restrict__ hc, int K, int S, int min_freq_bin, const float2* __restrict__ tw_g,
                      const unsigned long long* __restrict__ det_idx, const unsigned int* __restrict__ det_count,
                      unsigned int det_cap, DetectionRecord* __restrict__ recs, long long in_chan_stride,
                      long long z_chan_stride, long long det_chan_stride) {
    using G = Gen<LOGN>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    {   // batched channel mode: blockIdx.y is the channel; det_count points into an array of PeakState
        const long long ch = blockIdx.y;
        in += ch * in_chan_stride;
        zpow += ch * z_chan_stride;
        det_idx += ch * det_chan_stride;
        recs += ch * det_chan_stride;
        det_count += ch * (long long)(sizeof(PeakState) / sizeof(unsigned int));
    }
    float2* tw_s = reinterpret_cast<float2*>(smem_raw);
    float2* a = tw_s + G::N;
    float2* b = a + G::N;
    float2* corr_s = b + G::N;                                   // [kMaxHyp + 1]
    float* xpow = reinterpret_cast<float*>(corr_s + kMaxHyp + 1);   // [N / 2]: |X[f]|^2, f in [N/4, 3N/4)
    unsigned int n = *det_count;
    if (n > det_cap) n = det_cap;
    if (blockIdx.x >= n) return;
    const int tid = threadIdx.x;
    load_tw<LOGN>(tw_s, tw_g, tid);
    for (unsigned int d = blockIdx.x; d < n; d += gridDim.x) {
        const long long p = (long long)det_idx[d];
        const long long blk = p / S;
        const int kk = (int)(p - blk * S);
        const int m = (G::N - kk) & (G::N - 1);
        const float2* src = in + (blk * (long long)S - in_base);
        float2 xs[G::EPT];
#pragma unroll
        for (int j = 0; j < G::EPT; ++j) a[tid + j * G::NT] = __ldcs(src + tid + j * G::NT);
        __syncthreads();
        {
            const float2* y = fft_r2<LOGN>(a, b, tw_s, tid);
#pragma unroll
            for (int j = 0; j < G::EPT; ++j) {
                const int f = tid + j * G::NT;
                xs[j] = y[f];
                if (f >= G::N / 4 && f < 3 * G::N / 4) xpow[f - G::N / 4] = norm2_plain(xs[j]);
            }
        }
        for (int k = 0; k < K; ++k) {
            const float2* h = hc + (size_t)k * G::N + tid;
#pragma unroll
            for (int j = 0; j < G::EPT; ++j) a[tid + j * G::NT] = cmul_plain(xs[j], __ldg(h + j * G::NT));
            __syncthreads();
            const float2* y = fft_r2<LOGN>(a, b, tw_s, tid);
            if (tid == (m & (G::NT - 1))) corr_s[k] = y[m];   // by the OWNER of element m: the next product rewrites owners' elements only
        }
        __syncthreads();   // corr_s of the last hypothesis
        if (tid == 0) {
            // noise power: sequential float sum over f = N/4 .. 3N/4-1 in index order (:257-265)
            float acc = 0.0f;
            for (int f = 0; f < G::N / 2; ++f) acc = __fadd_rn(acc, xpow[f]);
            const float noise = __fdiv_rn(acc, __fmul_rn((float)(G::N / 2), (float)G::N));
            int best_freq = 0;  // :301-313
            float2 z = make_float2(0.f, 0.f);
            float zp = -1.0f;
            for (int k = 0; k < K; ++k) {
                const float q = norm2_plain(corr_s[k]);
                if (q > zp) { best_freq = k; z = corr_s[k]; zp = q; }
            }
            DetectionRecord r;
            r.index = (unsigned long long)p;
            r.corr_re = z.x;
            r.corr_im = z.y;
            r.pow = zp;
            r.pow_left = best_freq > 0 ? norm2_plain(corr_s[best_freq - 1]) : 0.0f;
            r.pow_right = best_freq < K - 1 ? norm2_plain(corr_s[best_freq + 1]) : 0.0f;
            r.pow_prev = (p - 1 >= 0 && p - 1 >= z_base) ? zpow[p - 1 - z_base] : 0.0f;
            r.pow_next = zpow[p + 1 - z_base];
            r.noise_power = noise;
            r.freq_bin = min_freq_bin + best_freq;
            r._pad = 0;
            recs[d] = r;
        }
        __syncthreads();
    }
}

template <int LOGN>
size_t refine_smem() {
    return Gen<LOGN>::smem_fft + sizeof(float2) * (kMaxHyp + 1) + sizeof(float) * (Gen<LOGN>::N / 2);
}

template <typename F>
cudaError_t raise_smem(F fn, size_t bytes) {
    return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <int LOGN>
cudaError_t do_template_spectra(const float2* d_td, float2* d_hc, int K, const float2* d_tw, cudaStream_t st) {
    cudaError_t e = raise_smem(template_spectra_generic_kernel<LOGN>, Gen<LOGN>::smem_fft);
    if (e != cudaSuccess) return e;
    template_spectra_generic_kernel<LOGN><<<K, Gen<LOGN>::NT, Gen<LOGN>::smem_fft, st>>>(d_td, d_hc, d_tw);
    count_launch();
    return cudaGetLastError();
}

template <int LOGN>
cudaError_t do_correlate(const float2* d_in, long long in_base, float* d_zpow, long long z_base, const float2* d_hc,
                         int K, int S, long long b0, long long nb, const float2* d_tw, float2* d_out_delayed,
                         long long out_base, long long out_lo, long long out_hi, int delay, int num_sms,
                         cudaStream_t st, long long nb_chan, long long in_chan_stride, long long z_chan_stride) {
    auto kern = correlate_generic_kernel<LOGN>;
    cudaError_t e = raise_smem(kern, Gen<LOGN>::smem_fft);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Gen<LOGN>::NT, Gen<LOGN>::smem_fft);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const long long want = (long long)num_sms * per_sm;
    const int grid = (int)(nb < want ? nb : want);
    kern<<<grid, Gen<LOGN>::NT, Gen<LOGN>::smem_fft, st>>>(d_in, in_base, d_zpow, z_base, d_hc, K, S, b0, nb, d_tw,
                                                          d_out_delayed, out_base, out_lo, out_hi, delay, nb_chan,
                                                          in_chan_stride, z_chan_stride);
    count_launch();
    return cudaGetLastError();
}

template <int LOGN>
cudaError_t do_refine(const float2* d_in, long long in_base, const float* d_zpow, long long z_base, const float2* d_hc,
                      int K, int S, int min_freq_bin, const float2* d_tw, const unsigned long long* d_det_idx,
                      const unsigned int* d_det_count, unsigned int det_cap, DetectionRecord* d_recs, int num_sms,
                      cudaStream_t st, int nch, long long in_chan_stride, long long z_chan_stride,
                      long long det_chan_stride) {
    auto kern = refine_generic_kernel<LOGN>;
    const size_t smem = refine_smem<LOGN>();
    cudaError_t e = raise_smem(kern, smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Gen<LOGN>::NT, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int grid = num_sms * per_sm;
    if (nch > 1) grid = (grid + nch - 1) / nch;
    if ((unsigned)grid > det_cap) grid = (int)det_cap;
    if (grid < 1) grid = 1;
    kern<<<dim3((unsigned)grid, (unsigned)nch), Gen<LOGN>::NT, smem, st>>>(
        d_in, in_base, d_zpow, z_base, d_hc, K, S, min_freq_bin, d_tw, d_det_idx, d_det_count, det_cap, d_recs,
        in_chan_stride, z_chan_stride, det_chan_stride);
    count_launch();
    return cudaGetLastError();
}

#define B200_GEN_DISPATCH(fft, CALL)                   \
    switch (fft) {                                     \
    case 64: return CALL(6);                           \
    case 128: return CALL(7);                          \
    case 256: return CALL(8);                          \
    case 512: return CALL(9);                          \
    case 1024: return CALL(10);                        \
    case 2048: return CALL(11);                        \
    case 4096: return CALL(12);                        \
    case 8192: return CALL(13);                        \
    default: return cudaErrorInvalidValue;             \
    }

}  // namespace

bool generic_fft_supported(unsigned fft) { return fft >= 64 && fft <= 8192 && (fft & (fft - 1)) == 0; }

cudaError_t launch_template_spectra_generic(int fft, const float2* d_td, float2* d_hc, int K, const float2* d_tw,
                                            cudaStream_t st) {
#define CALL(L) do_template_spectra<L>(d_td, d_hc, K, d_tw, st)
    B200_GEN_DISPATCH(fft, CALL)
#undef CALL
}

cudaError_t launch_correlate_generic(int fft, const float2* d_in, long long in_base, float* d_zpow, long long z_base,
                                     const float2* d_hc, int K, int S, long long b0, long long nb, const float2* d_tw,
                                     float2* d_out_delayed, long long out_base, long long out_lo, long long out_hi,
                                     int delay, int num_sms, cudaStream_t st, long long nb_chan,
                                     long long in_chan_stride, long long z_chan_stride) {
    if (nb <= 0) return cudaSuccess;
#define CALL(L)                                                                                                        \
    do_correlate<L>(d_in, in_base, d_zpow, z_base, d_hc, K, S, b0, nb, d_tw, d_out_delayed, out_base, out_lo, out_hi, \
                    delay, num_sms, st, nb_chan, in_chan_stride, z_chan_stride)
    B200_GEN_DISPATCH(fft, CALL)
#undef CALL
}

cudaError_t launch_refine_generic(int fft, const float2* d_in, long long in_base, const float* d_zpow, long long z_base,
                                  const float2* d_hc, int K, int S, int min_freq_bin, const float2* d_tw,
                                  const unsigned long long* d_det_idx, const unsigned int* d_det_count,
                                  unsigned int det_cap, DetectionRecord* d_recs, int num_sms, cudaStream_t st, int nch,
                                  long long in_chan_stride, long long z_chan_stride, long long det_chan_stride) {
#define CALL(L)                                                                                                   \
    do_refine<L>(d_in, in_base, d_zpow, z_base, d_hc, K, S, min_freq_bin, d_tw, d_det_idx, d_det_count, det_cap, \
                 d_recs, num_sms, st, nch, in_chan_stride, z_chan_stride, det_chan_stride)
    B200_GEN_DISPATCH(fft, CALL)
#undef CALL
}

}  // namespace b200sync

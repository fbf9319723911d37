// peaks.cu — parallel form of the running-max / timeout / median-threshold peak
// detector of SyncwordDetection (PM/syncword_detection.hpp:267-298, 314-317).
//
// The reference walks the stream with two scalars (_best, _best_idx):
//   at sample c:  if c - _best_idx > T  -> examine _best (count test), reset at c;
//                 if zpow[c] > _best    -> _best = zpow[c], _best_idx = c.
// Equivalent parallel statement (proved in DESIGN.md §4):
//   * p is a CANDIDATE iff no q in (p, p+T] has zpow[q] > zpow[p]   (forward-isolated);
//   * starting from search position r (r = 0 at stream start), the next EXAMINED
//     peak is the first candidate p >= r, and the search resumes at r' = p + T + 1;
//   * an examined p is a DETECTION iff #{q in [p-T, p+T] : zpow[q] < zpow[p]/thr}
//     (zpow[q<0] = 0, the zero-initialised HistoryBuffer) satisfies 2*count >= 2T+1.
// Kernels:
//   peak_flags_kernel   candidate bitmap (van Herk sliding max in shared memory) and
//                       threshold-test bitmap (warp popcount over the 2T+1 window)
//   chain_tables_kernel the search position only matters modulo pieces of T+1
//                       samples: each piece maps entry offset j in [0,T] to an exit
//                       offset; one warp composes the maps of M consecutive pieces
//                       into a (T+1)-entry table held in registers
//   chain_scan_kernel   composes segment maps for all T+1 entry states at once (constant
//                       segments cost one flag read) -> per-segment entry state of the
//                       real chain, and the whole-range table (multi-GPU stitching)
//   chain_emit_kernel   each warp re-walks its segment from its now-known entry
//                       state and appends examined&&passing peaks to the list
#include "b200sync_internal.h"

namespace b200sync {

constexpr int kFlagsTile = 4096;    // peaks decided per CTA
constexpr int kFlagsThreads = 512;
constexpr int kScanThreads = 1024;  // >= T+1

struct PeakPlan {
    long long range;   // hi - lo
    long long nwords;  // bitmap words (with 2 words of zero padding)
    long long nfr;     // pieces of T+1 samples
    int M;             // pieces per segment
    long long nseg;
    size_t off_cand, off_pass, off_tables, off_jin, off_flag, total;
};

static PeakPlan make_plan(long long range, int T, int num_sms) {
    PeakPlan p{};
    const long long Fr = T + 1;
    p.range = range;
    // every flags tile writes all of its kFlagsTile/32 words; +2 zero words of padding
    p.nwords = ((range + kFlagsTile - 1) / kFlagsTile) * (kFlagsTile / 32) + 2;
    p.nfr = (range + Fr - 1) / Fr;
    long long M = (p.nfr + (long long)num_sms * 32 - 1) / ((long long)num_sms * 32);
    if (M < 4) M = 4;
    p.M = (int)M;
    p.nseg = (p.nfr + M - 1) / M;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) & ~size_t(255); return r; };
    p.off_cand = take(sizeof(uint32_t) * p.nwords);
    p.off_pass = take(sizeof(uint32_t) * p.nwords);
    p.off_tables = take(sizeof(uint16_t) * (size_t)p.nseg * Fr);
    p.off_jin = take(sizeof(uint16_t) * (size_t)p.nseg);
    p.off_flag = take(sizeof(uint32_t) * (size_t)p.nseg);
    p.total = o;
    return p;
}

size_t peak_workspace_bytes_sms(long long max_range, int T, int num_sms) {
    // nseg is bounded by min(nfr/4+1, 32*num_sms+1) and monotone in range up to that bound,
    // so the plan of the largest range, with the segment areas sized by the bound, covers all.
    PeakPlan p = make_plan(max_range, T, num_sms);
    const long long Fr = T + 1;
    long long nseg_ub = p.nfr / 4 + 1;
    if (nseg_ub > 32LL * num_sms + 1) nseg_ub = 32LL * num_sms + 1;
    if (nseg_ub < p.nseg) nseg_ub = p.nseg;
    size_t total = 2 * ((sizeof(uint32_t) * p.nwords + 255) & ~size_t(255));
    total += (sizeof(uint16_t) * (size_t)nseg_ub * Fr + 255) & ~size_t(255);
    total += (sizeof(uint16_t) * (size_t)nseg_ub + 255) & ~size_t(255);
    total += (sizeof(uint32_t) * (size_t)nseg_ub + 255) & ~size_t(255);
    return total + 1024;
}

__device__ __forceinline__ float warp_scan_max(float v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v = fmaxf(v, t);
    }
    return v;
}

// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFlagsThreads)
peak_flags_generic_kernel(const float* __restrict__ zpow, long long z_base, long long z_end, long long lo,
                  long long hi, int T, float thr, uint32_t* __restrict__ cand_bits,
                  uint32_t* __restrict__ pass_bits) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = kFlagsTile + 2 * T;
    float* z = reinterpret_cast<float*>(smem_raw);
    float* Sx = z + n;   // prefix max within W-blocks
    float* Rx = Sx + n;  // suffix max within W-blocks
    uint32_t* passw = reinterpret_cast<uint32_t*>(Rx + n);
    unsigned short* cand_list = reinterpret_cast<unsigned short*>(passw + kFlagsTile / 32);
    __shared__ int ncand;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kFlagsThreads / 32;
    const long long tile_lo = lo + (long long)blockIdx.x * kFlagsTile;
    // stage zpow[tile_lo - T, tile_lo + TILE + T); out-of-stream / not-yet-known -> 0
    for (int i = tid; i < n; i += kFlagsThreads) {
        const long long q = tile_lo - T + i;
        z[i] = (q >= 0 && q < z_end) ? zpow[q - z_base] : 0.0f;
    }
    if (tid < kFlagsTile / 32) passw[tid] = 0u;
    if (tid == 0) ncand = 0;
    __syncthreads();

    if (T > 0) {
        // van Herk / Gil-Werman: blocks of W = T; window (p, p+T] = [a, a+W-1], a = p+1
        const int W = T;
        const int nblk = (n + W - 1) / W;
        const float NEG = -__int_as_float(0x7f800000);
        for (int b = warp; b < nblk; b += nwarps) {
            const int s = b * W, e = min(s + W, n);
            float run = NEG;
            for (int c = s; c < e; c += 32) {
                const int idx = c + lane;
                float v = idx < e ? z[idx] : NEG;
                v = fmaxf(warp_scan_max(v, lane), run);
                if (idx < e) Sx[idx] = v;
                run = __shfl_sync(0xffffffffu, v, 31);
            }
            run = NEG;
            for (int c = e; c > s; c -= 32) {
                const int idx = c - 1 - lane;
                float v = idx >= s ? z[idx] : NEG;
                v = fmaxf(warp_scan_max(v, lane), run);
                if (idx >= s) Rx[idx] = v;
                run = __shfl_sync(0xffffffffu, v, 31);
            }
        }
    }
    __syncthreads();

    // candidate flags
    for (int tp = tid; tp < kFlagsTile; tp += kFlagsThreads) {
        const int i = T + tp;
        const long long p = tile_lo + tp;
        bool cand = false;
        if (p < hi) {
            if (T > 0) {
                const float fwd = fmaxf(Rx[i + 1], Sx[i + T]);
                cand = !(fwd > z[i]);
            } else {
                cand = true;
            }
        }
        const uint32_t w = __ballot_sync(0xffffffffu, cand);
        if (lane == 0) cand_bits[(tile_lo - lo) / 32 + (tp >> 5)] = w;
        if (cand) cand_list[atomicAdd(&ncand, 1)] = (unsigned short)tp;
    }
    __syncthreads();

    // threshold test for every candidate: count history items below best/thr (:273-279)
    const int nc = ncand;
    for (int ci = warp; ci < nc; ci += nwarps) {
        const int tp = cand_list[ci];
        const int i = T + tp;
        const float tv = __fdiv_rn(z[i], thr);
        int cnt = 0;
        if (tv > 0.0f) {  // zpow >= 0: nothing is below a non-positive threshold
            for (int u = i - T + lane; u <= i + T; u += 32) cnt += (z[u] < tv) ? 1 : 0;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        if (lane == 0 && 2 * cnt >= 2 * T + 1) atomicOr(&passw[tp >> 5], 1u << (tp & 31));
    }
    __syncthreads();
    if (tid < kFlagsTile / 32) pass_bits[(tile_lo - lo) / 32 + tid] = passw[tid];
}


// ---------------------------------------------------------------------------------
// Fast flags kernel for T >= 32.  The tile's zpow window is staged in shared memory with a
// 33-word stride per 32-sample group, so one THREAD can run a sequential prefix / suffix max
// over a whole group with no bank conflicts (lane g reads word 33 g + e: bank (g + e) mod 32).
// The forward window (p, p+T] = [a, b] then decomposes into
//     suffix-in-group(a)  |  q whole groups  |  prefix-in-group(b),   q in {m0-1, m0}, m0 = (T-1)/32
// and the whole-group part is a per-group sliding maximum over the group maxima.
// ~6 instructions per sample instead of ~45 for the warp-shuffle scans of the generic kernel.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ int padi(int i) { return i + (i >> 5); }

__global__ void __launch_bounds__(kFlagsThreads)
peak_flags_kernel(const float* __restrict__ zpow, long long z_base, long long z_end, long long lo,
                  long long hi, int T, float thr, uint32_t* __restrict__ cand_bits,
                  uint32_t* __restrict__ pass_bits) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = kFlagsTile + 2 * T;
    const int ng = (n + 31) >> 5;          // groups of 32 (the last may be partial: padded with 0)
    const int npad = ng * 33;
    float* z = reinterpret_cast<float*>(smem_raw);
    float* pre = z + npad;                 // prefix max within group, inclusive
    float* suf = pre + npad;               // suffix max within group, inclusive
    float* gmax = suf + npad;              // [ng + 64] group maxima, -inf beyond ng
    float* gmid = gmax + ng + 64;          // [ng] max of gmax[g+1 .. g+m0-1]
    uint32_t* passw = reinterpret_cast<uint32_t*>(gmid + ng);
    unsigned short* cand_list = reinterpret_cast<unsigned short*>(passw + kFlagsTile / 32);
    __shared__ int ncand;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kFlagsThreads / 32;
    const long long tile_lo = lo + (long long)blockIdx.x * kFlagsTile;
    const float NEG = -__int_as_float(0x7f800000);
    {
        // stage the window: all global loads issued before the first shared store
        // (n <= 4096 + 2*1023 -> at most 13 rounds of 512 threads)
        constexpr int kRounds = (kFlagsTile + 2 * kMaxTimeThreshold + 31 + kFlagsThreads - 1) / kFlagsThreads;
        const long long q0 = tile_lo - T;                       // absolute index of window element 0
        const long long lo_ok = q0 < 0 ? -q0 : 0;               // first element with q >= 0
        const long long hi_ok = min((long long)n, z_end - q0);  // first element with q >= z_end (or n)
        const float* src = zpow + (q0 - z_base);
        float v[kRounds];
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {
            const int i = tid + r * kFlagsThreads;
            v[r] = (i >= lo_ok && i < hi_ok) ? src[i] : 0.0f;
        }
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {
            const int i = tid + r * kFlagsThreads;
            if (i < ng * 32) z[padi(i)] = v[r];
        }
    }
    if (tid < kFlagsTile / 32) passw[tid] = 0u;
    if (tid == 0) ncand = 0;
    for (int g = ng + tid; g < ng + 64; g += kFlagsThreads) gmax[g] = NEG;
    __syncthreads();
    // per-group sequential scans: threads [0, ng) do prefixes, threads [ng, 2 ng) do suffixes
    for (int w = tid; w < 2 * ng; w += kFlagsThreads) {
        if (w < ng) {
            const int base = w * 33;
            float run = NEG;
#pragma unroll 8
            for (int e = 0; e < 32; ++e) {
                run = fmaxf(run, z[base + e]);
                pre[base + e] = run;
            }
            gmax[w] = run;
        } else {
            const int base = (w - ng) * 33;
            float run = NEG;
#pragma unroll 8
            for (int e = 31; e >= 0; --e) {
                run = fmaxf(run, z[base + e]);
                suf[base + e] = run;
            }
        }
    }
    __syncthreads();
    const int m0 = (T - 1) >> 5;
    for (int g = tid; g < ng; g += kFlagsThreads) {
        float m = NEG;
        for (int k = 1; k < m0; ++k) m = fmaxf(m, gmax[g + k]);
        gmid[g] = m;
    }
    __syncthreads();

    {
        // candidate flags.  All per-sample indices advance by constants between rounds
        // (kFlagsThreads is a multiple of 32), so they are strength-reduced by hand.
        const long long rem = hi - tile_lo;
        const int nvalid = rem < (long long)kFlagsTile ? (int)rem : kFlagsTile;  // p < hi
        const int a0 = T + 1 + tid, b0 = a0 + T - 1;
        int ga = a0 >> 5;
        const int dg = (b0 >> 5) - ga;           // gb - ga: the same in every round
        const bool whole = (dg == 0);            // [a, b] is exactly one whole group
        const bool extra = (m0 >= 1) && (dg - 1 == m0);
        int pa = padi(a0), pb = padi(b0), pz = padi(a0 - 1);
        uint32_t* cw = cand_bits + (tile_lo - lo) / 32 + warp;
        constexpr int kStep = kFlagsThreads + kFlagsThreads / 32;  // padded-index stride per round
#pragma unroll
        for (int r = 0; r < kFlagsTile / kFlagsThreads; ++r) {
            const int tp = tid + r * kFlagsThreads;
            float fwd = suf[pa];
            if (!whole) {
                fwd = fmaxf(fmaxf(fwd, pre[pb]), gmid[ga]);
                if (extra) fwd = fmaxf(fwd, gmax[ga + m0]);
            }
            const bool cand = (tp < nvalid) && !(fwd > z[pz]);
            const uint32_t w = __ballot_sync(0xffffffffu, cand);
            if (lane == 0) cw[r * (kFlagsThreads / 32)] = w;
            if (cand) cand_list[atomicAdd(&ncand, 1)] = (unsigned short)tp;
            pa += kStep; pb += kStep; pz += kStep; ga += kFlagsThreads / 32;
        }
    }
    __syncthreads();

    // threshold test for every candidate: count history items below best/thr (:273-279)
    const int nc = ncand;
    for (int ci = warp; ci < nc; ci += nwarps) {
        const int tp = cand_list[ci];
        const int i = T + tp;
        const float tv = __fdiv_rn(z[padi(i)], thr);
        int cnt = 0;
        if (tv > 0.0f) {  // zpow >= 0: nothing is below a non-positive threshold
            for (int u = i - T + lane; u <= i + T; u += 32) cnt += (z[padi(u)] < tv) ? 1 : 0;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        if (lane == 0 && 2 * cnt >= 2 * T + 1) atomicOr(&passw[tp >> 5], 1u << (tp & 31));
    }
    __syncthreads();
    if (tid < kFlagsTile / 32) pass_bits[(tile_lo - lo) / 32 + tid] = passw[tid];
}

// bits [fstart + 32*lane, +32) of a piece of length L starting at bit fstart
__device__ __forceinline__ uint32_t piece_word(const uint32_t* __restrict__ bits, long long fstart, int L,
                                               int lane) {
    const int off = 32 * lane;
    if (off >= L) return 0u;
    const long long o = fstart + off;
    const long long wi = o >> 5;
    const int sh = (int)(o & 31);
    uint32_t w = __funnelshift_r(bits[wi], bits[wi + 1], sh);
    const int valid = L - off;
    if (valid < 32) w &= (1u << valid) - 1u;
    return w;
}

// ---------------------------------------------------------------------------------
// next-candidate lookup inside one piece: first set bit >= x (x < L), or -1
__device__ __forceinline__ int piece_next(uint32_t w, uint32_t nz, int x) {
    const int wq = x >> 5;
    const uint32_t mw = __shfl_sync(0xffffffffu, w, wq) & (0xffffffffu << (x & 31));
    const uint32_t m2 = (wq >= 31) ? 0u : (nz & (0xfffffffeu << wq));
    const int w2 = m2 ? (__ffs(m2) - 1) : 0;
    const uint32_t ww2 = __shfl_sync(0xffffffffu, w, w2);
    if (mw) return 32 * wq + __ffs(mw) - 1;
    if (m2) return 32 * w2 + __ffs(ww2) - 1;
    return -1;
}

// segflag[seg]: bit 31 set -> the segment maps EVERY entry offset to (segflag & 0xffff)
// (its table is not written); 0 -> general table in tables[seg].
__global__ void __launch_bounds__(128)
chain_tables_kernel(const uint32_t* __restrict__ cand_bits, long long range, int T, int M, long long nfr,
                    long long nseg, uint16_t* __restrict__ tables, uint32_t* __restrict__ segflag) {
    const int lane = threadIdx.x & 31;
    const long long seg = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (seg >= nseg) return;
    const int Fr = T + 1;
    int ent[32];  // table entries j = lane + 32 e
#pragma unroll
    for (int e = 0; e < 32; ++e) ent[e] = lane + 32 * e;
    bool uniform = (Fr == 1);
    int uval = 0;
    const long long f_end = min(nfr, (seg + 1) * (long long)M);
    for (long long f = seg * (long long)M; f < f_end; ++f) {
        const long long fstart = f * Fr;
        const int L = (int)min((long long)Fr, range - fstart);
        const uint32_t w = piece_word(cand_bits, fstart, L, lane);
        const uint32_t nz = __ballot_sync(0xffffffffu, w != 0u);
        if (uniform) {
            // the table has collapsed to one value (and stays collapsed): one lookup per piece
            const int x = uval;
            const int a = piece_next(w, nz, x < L ? x : 0);
            uval = (x >= L) ? (x - L) : ((a >= 0) ? (a + Fr - L) : 0);
            continue;
        }
        bool same = true;
        int v0 = 0;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            const int x = ent[e];
            const int a = piece_next(w, nz, x < L ? x : 0);
            const int y = (x >= L) ? (x - L) : ((a >= 0) ? (a + Fr - L) : 0);
            ent[e] = y;
            if (e == 0) v0 = __shfl_sync(0xffffffffu, y, 0);
            if (lane + 32 * e < Fr) same = same && (y == v0);
        }
        if (__all_sync(0xffffffffu, same)) {
            uniform = true;
            uval = v0;
        }
    }
    if (uniform) {
        if (lane == 0) segflag[seg] = 0x80000000u | (uint32_t)uval;
    } else {
        if (lane == 0) segflag[seg] = 0u;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            const int j = lane + 32 * e;
            if (j < Fr) tables[seg * Fr + j] = (uint16_t)ent[e];
        }
    }
}

// ---------------------------------------------------------------------------------
// Composes the segment maps for all T+1 entry states at once (thread t follows state t).
// Constant segments (almost all of them on real data) cost one shared-memory flag read;
// only general segments stage their table.
// mode 0: only the whole-range table; mode 1: walk the real chain (entry j_in or derived from
// state), write per-segment entry states and the new search position.
__global__ void __launch_bounds__(kScanThreads)
chain_scan_kernel(const uint16_t* __restrict__ tables, const uint32_t* __restrict__ segflag, long long nseg_ll,
                  int T, int mode, int j_in_param, PeakState* __restrict__ state, long long lo, long long hi,
                  uint16_t* __restrict__ seg_jin, uint16_t* __restrict__ range_table) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nseg = (int)nseg_ll;
    uint32_t* flags = reinterpret_cast<uint32_t*>(smem_raw);         // [nseg]
    uint16_t* jin_s = reinterpret_cast<uint16_t*>(flags + nseg);     // [nseg + 1] entry state per segment
    uint16_t* tab = jin_s + nseg + 2;                                // [T + 1]
    __shared__ int first_const, final_state;
    const int Fr = T + 1;
    const int t = threadIdx.x;
    int j_in = -1;
    if (mode == 1) {
        if (j_in_param >= 0) j_in = j_in_param;
        else {
            const unsigned long long r = state->r_abs;
            j_in = (r > (unsigned long long)lo) ? (int)(r - (unsigned long long)lo) : 0;
        }
        if (j_in > T) j_in = T;
    }
    if (t == 0) first_const = nseg;
    __syncthreads();
    for (int i = t; i < nseg; i += kScanThreads) {
        const uint32_t f = segflag[i];
        flags[i] = f;
        if (f & 0x80000000u) atomicMin(&first_const, i);
    }
    __syncthreads();
    const int fc = first_const;
    // (1) all T+1 entry states through the leading general segments (none on ordinary data)
    int cur = t < Fr ? t : 0;
    for (int s = 0; s < fc; ++s) {
        if (t == j_in) jin_s[s] = (uint16_t)cur;
        __syncthreads();
        for (int i = t; i < Fr; i += kScanThreads) tab[i] = tables[(long long)s * Fr + i];
        __syncthreads();
        cur = tab[cur];
    }
    if (fc == nseg) {  // no constant segment at all: every state keeps its own image
        if (t < Fr && range_table != nullptr) range_table[t] = (uint16_t)cur;
        if (mode == 1 && t == j_in) state->r_abs = (unsigned long long)hi + (unsigned long long)cur;
        if (mode == 1) {
            __syncthreads();
            for (int i = t; i < nseg; i += kScanThreads) seg_jin[i] = jin_s[i];
        }
        return;
    }
    // (2) from the first constant segment on all states have merged into one chain.  Entry of
    // segment s+1 is known at once wherever segment s is constant ...
    if (t == j_in || (mode == 0 && t == 0)) jin_s[fc] = (uint16_t)cur;
    for (int s = fc + t; s < nseg; s += kScanThreads)
        if (flags[s] & 0x80000000u) jin_s[s + 1] = (uint16_t)(flags[s] & 0xffffu);
    __syncthreads();
    // ... and the (rare) general segments are fixed up in order by one thread
    if (t < 32) {  // warp 0: 32 segments per ballot, general ones handled in order by lane 0
        for (int s0 = fc + 1; s0 < nseg; s0 += 32) {
            const int s = s0 + t;
            uint32_t gen = __ballot_sync(0xffffffffu, s < nseg && !(flags[s] & 0x80000000u));
            if (t == 0) {
                while (gen) {
                    const int sg = s0 + __ffs(gen) - 1;
                    gen &= gen - 1;
                    jin_s[sg + 1] = tables[(long long)sg * Fr + jin_s[sg]];
                }
            }
            __syncwarp();
        }
        if (t == 0) final_state = jin_s[nseg];
    }
    __syncthreads();
    const int fin = final_state;
    if (t < Fr && range_table != nullptr) range_table[t] = (uint16_t)fin;
    if (mode == 1) {
        if (t == 0) state->r_abs = (unsigned long long)hi + (unsigned long long)fin;
        for (int i = t; i < nseg; i += kScanThreads) seg_jin[i] = jin_s[i];
    }
}

// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
chain_emit_kernel(const uint32_t* __restrict__ cand_bits, const uint32_t* __restrict__ pass_bits,
                  long long range, int T, int M, long long nfr, long long nseg,
                  const uint16_t* __restrict__ seg_jin, long long lo, unsigned long long* __restrict__ det_idx,
                  unsigned int det_cap, PeakState* __restrict__ state) {
    const int lane = threadIdx.x & 31;
    const long long seg = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (seg >= nseg) return;
    const int Fr = T + 1;
    int j = seg_jin[seg];
    const long long f_end = min(nfr, (seg + 1) * (long long)M);
    for (long long f = seg * (long long)M; f < f_end; ++f) {
        const long long fstart = f * Fr;
        const int L = (int)min((long long)Fr, range - fstart);
        uint32_t w = piece_word(cand_bits, fstart, L, lane);
        const int off = 32 * lane;
        if (off + 31 < j) w = 0u;
        else if (off < j) w &= 0xffffffffu << (j - off);
        const uint32_t m = __ballot_sync(0xffffffffu, w != 0u);
        if (j < L && m != 0u) {
            const int l0 = __ffs(m) - 1;
            const uint32_t ww = __shfl_sync(0xffffffffu, w, l0);
            const int a = 32 * l0 + __ffs(ww) - 1;
            if (lane == 0) {
                const long long bit = fstart + a;
                if ((pass_bits[bit >> 5] >> (bit & 31)) & 1u) {
                    const unsigned int slot = atomicAdd(&state->det_count, 1u);
                    if (slot < det_cap) det_idx[slot] = (unsigned long long)(lo + bit);
                }
            }
            j = a + Fr - L;
        } else {
            j = (j >= L) ? (j - L) : 0;
        }
    }
}

// ---------------------------------------------------------------------------------
static cudaError_t set_smem_attr(const void* fn, size_t bytes) {
    return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t launch_peak_phase1(const float* d_zpow, long long z_base, long long z_end, long long lo,
                               long long hi, int T, float power_threshold, void* d_ws, size_t ws_bytes,
                               uint16_t* d_range_table, int num_sms, cudaStream_t st) {
    const long long range = hi - lo;
    if (range <= 0) return cudaSuccess;
    const PeakPlan pl = make_plan(range, T, num_sms);
    if (pl.total > ws_bytes) return cudaErrorInvalidValue;
    unsigned char* ws = static_cast<unsigned char*>(d_ws);
    uint32_t* cand = reinterpret_cast<uint32_t*>(ws + pl.off_cand);
    uint32_t* pass = reinterpret_cast<uint32_t*>(ws + pl.off_pass);
    uint16_t* tables = reinterpret_cast<uint16_t*>(ws + pl.off_tables);
    cudaError_t e;
    const long long ntiles = (range + kFlagsTile - 1) / kFlagsTile;
    // zero the padding words past the last tile (tiles write every word they own)
    const long long written = ntiles * (kFlagsTile / 32);
    if (written < pl.nwords) {
        e = cudaMemsetAsync(cand + written, 0, sizeof(uint32_t) * (pl.nwords - written), st);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(pass + written, 0, sizeof(uint32_t) * (pl.nwords - written), st);
        if (e != cudaSuccess) return e;
    }
    uint32_t* segflag = reinterpret_cast<uint32_t*>(ws + pl.off_flag);
    const int n = kFlagsTile + 2 * T;
    if (T >= 32) {
        const int ng = (n + 31) / 32;
        const size_t smem = sizeof(float) * (3 * (size_t)ng * 33 + (ng + 64) + ng) +
                            sizeof(uint32_t) * (kFlagsTile / 32) + sizeof(unsigned short) * kFlagsTile;
        e = set_smem_attr((const void*)peak_flags_kernel, smem);
        if (e != cudaSuccess) return e;
        peak_flags_kernel<<<(unsigned)ntiles, kFlagsThreads, smem, st>>>(d_zpow, z_base, z_end, lo, hi, T,
                                                                        power_threshold, cand, pass);
    } else {
        const size_t smem = sizeof(float) * 3 * n + sizeof(uint32_t) * (kFlagsTile / 32) +
                            sizeof(unsigned short) * kFlagsTile;
        e = set_smem_attr((const void*)peak_flags_generic_kernel, smem);
        if (e != cudaSuccess) return e;
        peak_flags_generic_kernel<<<(unsigned)ntiles, kFlagsThreads, smem, st>>>(d_zpow, z_base, z_end, lo, hi,
                                                                                T, power_threshold, cand, pass);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    chain_tables_kernel<<<(unsigned)((pl.nseg + 3) / 4), 128, 0, st>>>(cand, range, T, pl.M, pl.nfr,
                                                                       pl.nseg, tables, segflag);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (d_range_table != nullptr) {
        const size_t ssm = sizeof(uint32_t) * (size_t)pl.nseg + sizeof(uint16_t) * (size_t)(pl.nseg + 2 + T + 2);
        e = set_smem_attr((const void*)chain_scan_kernel, ssm);
        if (e != cudaSuccess) return e;
        chain_scan_kernel<<<1, kScanThreads, ssm, st>>>(tables, segflag, pl.nseg, T, 0, -1, nullptr, lo, hi,
                                                        nullptr, d_range_table);
        e = cudaGetLastError();
    }
    return e;
}

cudaError_t launch_peak_phase2(long long lo, long long hi, int T, void* d_ws, size_t ws_bytes, int j_in,
                               PeakState* d_state, unsigned long long* d_det_idx, unsigned int det_cap,
                               int num_sms, cudaStream_t st) {
    const long long range = hi - lo;
    if (range <= 0) return cudaSuccess;
    const PeakPlan pl = make_plan(range, T, num_sms);
    if (pl.total > ws_bytes) return cudaErrorInvalidValue;
    unsigned char* ws = static_cast<unsigned char*>(d_ws);
    uint32_t* cand = reinterpret_cast<uint32_t*>(ws + pl.off_cand);
    uint32_t* pass = reinterpret_cast<uint32_t*>(ws + pl.off_pass);
    uint16_t* tables = reinterpret_cast<uint16_t*>(ws + pl.off_tables);
    uint16_t* jin = reinterpret_cast<uint16_t*>(ws + pl.off_jin);
    const uint32_t* segflag = reinterpret_cast<const uint32_t*>(ws + pl.off_flag);
    const size_t ssm = sizeof(uint32_t) * (size_t)pl.nseg + sizeof(uint16_t) * (size_t)(pl.nseg + 2 + T + 2);
    cudaError_t e = set_smem_attr((const void*)chain_scan_kernel, ssm);
    if (e != cudaSuccess) return e;
    chain_scan_kernel<<<1, kScanThreads, ssm, st>>>(tables, segflag, pl.nseg, T, 1, j_in, d_state, lo, hi, jin,
                                                    nullptr);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    chain_emit_kernel<<<(unsigned)((pl.nseg + 3) / 4), 128, 0, st>>>(cand, pass, range, T, pl.M, pl.nfr,
                                                                     pl.nseg, jin, lo, d_det_idx, det_cap,
                                                                     d_state);
    return cudaGetLastError();
}

}  // namespace b200sync

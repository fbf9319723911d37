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
#include <climits>
#include <type_traits>

#include "b200sync_internal.h"
#include "peak_walk.cuh"

namespace b200sync {

constexpr int kFlagsTile = 4096;    // peaks decided per CTA
constexpr int kFlagsThreads = 512;
constexpr int kScanThreads = 1024;  // >= T+1
constexpr int kPlanTile = 8192;     // == kFastTile (a multiple of kFlagsTile)

// Batched channel mode: blockIdx.y is the channel; every per-channel array of a kernel lies `stride` bytes (or
// elements, as named) after the previous channel's.  Single-stream launches pass zeros and gridDim.y = 1.
struct PeakBatch {
    long long z_stride;      // floats between the metric arrays of consecutive channels
    size_t ws_stride;        // bytes between the workspaces of consecutive channels
    size_t det_stride;       // entries between the detection lists of consecutive channels
};
template <typename T>
__device__ __forceinline__ T* ws_at(T* p, size_t stride_bytes) {
    return reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(const_cast<typename std::remove_const<T>::type*>(p)) +
                                (size_t)blockIdx.y * stride_bytes);
}

struct PeakPlan {
    long long range;   // hi - lo
    long long nwords;  // bitmap words (with 2 words of zero padding)
    long long nfr;     // pieces of T+1 samples
    int M;             // pieces per segment
    long long nseg;
    size_t off_cand, off_pass, off_tables, off_jin, off_flag, off_slots, off_segcnt, off_segoff, total;
};

static PeakPlan make_plan(long long range, int T, int num_sms) {
    PeakPlan p{};
    const long long Fr = T + 1;
    p.range = range;
    // every flags tile writes all of its words (the larger tile of the two flags kernels bounds
    // both); +2 zero words of padding
    p.nwords = ((range + kPlanTile - 1) / kPlanTile) * (kPlanTile / 32) + 2;
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
    p.off_slots = take(sizeof(unsigned long long) * (size_t)p.nseg * (size_t)p.M);  // <= one detection per piece
    p.off_segcnt = take(sizeof(uint32_t) * (size_t)p.nseg);
    p.off_segoff = take(sizeof(uint32_t) * (size_t)(p.nseg + 1));
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
    // detection slots: nseg * M <= nfr + M, both monotone in the range
    total += (sizeof(unsigned long long) * (size_t)(p.nfr + p.M + 4) + 255) & ~size_t(255);
    total += 2 * ((sizeof(uint32_t) * (size_t)(nseg_ub + 1) + 255) & ~size_t(255));
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
                  uint32_t* __restrict__ pass_bits, PeakBatch pb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    zpow += (long long)blockIdx.y * pb.z_stride;
    cand_bits = ws_at(cand_bits, pb.ws_stride);
    pass_bits = ws_at(pass_bits, pb.ws_stride);
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
// Fast flags kernel for T >= 32: no per-sample scans, and no per-sample work at all for most groups.
//
// The tile's zpow window is staged once in shared memory (128-bit loads/stores) together with
// the maximum of every aligned 32-sample group.  For a sample p of group G the forward window (p, p+T] is
//     rest of group G after p  |  q whole groups G+1..G+q  |  a remainder of <= 62 samples,
// q = (T-31)/32.  A group whose maximum is below the maximum M_G of its q whole groups holds no candidate: ONE
// compare per GROUP rejects about q in q+1 of them; only the others compute the two partial maxima, with
// warp-shuffle scans, and decide their 32 samples exactly.  The threshold test of a candidate (count of window
// samples below zpow[p]/thr, :273-279) consults the group maxima first: groups entirely below the threshold are
// settled by one lane each (a real peak: nearly all of them, the few left are counted with masks); a noise
// maximum, whose threshold lies near the median of its window, counts the 2T+1 samples directly (3 instructions
// per 32 samples).  The minimum of the whole staged window settles the degenerate capture (constant input: every
// sample is a candidate, nothing is below the threshold) and bounds the cost there.
//
// Ordering is done on the int32 bit patterns of zpow, which is the float ordering because zpow is
// a squared magnitude (>= +0; not NaN for finite input); the direct count takes the sign bit of the difference
// of two such patterns (both in [0, 2^31): no overflow).
// ---------------------------------------------------------------------------------
constexpr int kFastTile = 8192;   // peaks decided per CTA (captures)
constexpr int kFastTileSmall = 2048;  // ... for streaming-sized ranges: more CTAs, a quarter of the latency each
constexpr int kFastThreads = 512;

__device__ __forceinline__ int warp_incl_scan_max(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v = max(v, t);
    }
    return v;
}

struct FastGeom {
    int Tpad, ng, nrow;
};
__host__ __device__ inline FastGeom fast_geom(int T, int tile) {
    FastGeom g;
    g.Tpad = (T + 31) & ~31;                          // window origin is group aligned with the tile
    const int n = g.Tpad + tile + T + 1;              // window elements that can be referenced
    g.ng = (n + 31) >> 5;
    g.nrow = (g.ng + 4 + 3) >> 2;                     // rows of 128 samples, >= 4 groups of slack
    return g;
}

template <int TILE>
__global__ void __launch_bounds__(kFastThreads, 4)   // 4 CTAs of 512 threads = every warp slot of the SM: <= 32 registers
peak_flags_kernel(const float* __restrict__ zpow, long long z_base, long long z_end, long long lo,
                  long long hi, int T, float thr, uint32_t* __restrict__ cand_bits,
                  uint32_t* __restrict__ pass_bits, PeakBatch pb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    zpow += (long long)blockIdx.y * pb.z_stride;
    cand_bits = ws_at(cand_bits, pb.ws_stride);
    pass_bits = ws_at(pass_bits, pb.ws_stride);
    const FastGeom geo = fast_geom(T, TILE);
    const int Tpad = geo.Tpad, nrow = geo.nrow;
    constexpr int kWarps = kFastThreads / 32;
    int* z = reinterpret_cast<int*>(smem_raw);             // [nrow * 128] bit patterns of zpow
    int* gmax = z + nrow * 128;                            // [nrow * 4]
    int* gM = gmax + nrow * 4;                             // [TILE / 32] max of the q whole groups
    int* wmin = gM + TILE / 32;                            // [kWarps] minimum of a warp's share (degenerate captures)
    int* nposs_s = wmin + kWarps;                          // [1] possible groups
    unsigned short* poss_list = reinterpret_cast<unsigned short*>(nposs_s + 1);  // [TILE / 32]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_launch_dependents();   // streaming: the walk + refine launch may move in behind this grid
    pdl_wait();                // ... and this one behind the correlator: zpow is complete and visible from here on
    const long long tile_lo = lo + (long long)blockIdx.x * TILE;
    const long long q0 = tile_lo - Tpad;                   // absolute index of window element 0
    const int* src = reinterpret_cast<const int*>(zpow) + (q0 - z_base);
    // first element that is both inside the stream and inside the caller's metric buffer: the window origin is
    // rounded down to a group boundary (Tpad >= T), which can reach below z_base when a caller keeps exactly
    // the 2T+2 samples of history the decisions need (streaming, short shard halos).  Those pad elements are
    // never part of a decision window — they only enter the maximum of a group that straddles the window edge,
    // where a zero can only send the count to the exact per-sample path — so they read as zeros.
    const long long first_ok = z_base > 0 ? z_base : 0;
    const int lo_ok = q0 < first_ok ? (int)(first_ok - q0) : 0;
    const long long hi_ll = z_end - q0;                    // first element not yet known
    const int hi_ok = hi_ll < (long long)(nrow * 128) ? (int)hi_ll : nrow * 128;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    // ---- stage rows of 128 samples; out-of-stream / not-yet-known samples read as 0 (the
    //      zero-initialised HistoryBuffer before the stream start)
    int* const gmax_lane = gmax + (lane >> 3);
    auto put_row = [&](int r, const int4& v) {             // the row into shared memory + its four groups' maxima
        *reinterpret_cast<int4*>(z + r * 128 + 4 * lane) = v;
        int mx = max(max(v.x, v.y), max(v.z, v.w));
#pragma unroll
        for (int dd = 1; dd < 8; dd <<= 1)                 // 8 lanes hold one 32-sample group
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, dd));
        if ((lane & 7) == 0) gmax_lane[r * 4] = mx;
    };
    if (vec_ok && lo_ok == 0 && hi_ok == nrow * 128) {
        // interior tile (all but the first and last few of a capture): no per-element bounds, four rows in flight
        const int4* src4 = reinterpret_cast<const int4*>(src) + lane;
        for (int r0 = warp; r0 < nrow; r0 += 4 * kWarps) {
            int4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = r0 + u * kWarps;
                if (r < nrow) v[u] = __ldg(src4 + r * 32);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = r0 + u * kWarps;
                if (r < nrow) put_row(r, v[u]);
            }
        }
    } else {
        for (int r0 = warp; r0 < nrow; r0 += 2 * kWarps) {
            int4 v[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int r = r0 + u * kWarps;
                const int e0 = r * 128 + 4 * lane;
                v[u] = make_int4(0, 0, 0, 0);
                if (r < nrow) {
                    if (vec_ok && e0 >= lo_ok && e0 + 3 < hi_ok) {
                        v[u] = __ldg(reinterpret_cast<const int4*>(src + e0));
                    } else {
                        if (e0 + 0 >= lo_ok && e0 + 0 < hi_ok) v[u].x = __ldg(src + e0 + 0);
                        if (e0 + 1 >= lo_ok && e0 + 1 < hi_ok) v[u].y = __ldg(src + e0 + 1);
                        if (e0 + 2 >= lo_ok && e0 + 2 < hi_ok) v[u].z = __ldg(src + e0 + 2);
                        if (e0 + 3 >= lo_ok && e0 + 3 < hi_ok) v[u].w = __ldg(src + e0 + 3);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int r = r0 + u * kWarps;
                if (r < nrow) put_row(r, v[u]);
            }
        }
    }
    if (tid == 0) *nposs_s = 0;
    __syncthreads();
    const int q = (T - 31) >> 5;         // whole groups inside every forward window of a group
    const int d = T - 32 * q - 32;       // remainder reaches element (lane + d) of group G+q+1, d in [-1, 30]
    const int G0 = Tpad >> 5;
    const long long rem_ll = hi - tile_lo;
    const int nvalid = rem_ll < (long long)TILE ? (int)rem_ll : TILE;  // p < hi
    const int tv_need = 2 * T + 1;
    uint32_t* cw = cand_bits + (tile_lo - lo) / 32;
    uint32_t* pw = pass_bits + (tile_lo - lo) / 32;
    // ---- one thread per tile group: the maximum M of the q whole groups inside the forward window of every
    //      sample of the group.  A group whose own maximum is below M holds no candidate (about q in q+1 of them):
    //      settled here by ONE compare per group — both its bitmap words are zero; the others are queued.
    static_assert((TILE / 32) % 32 == 0 && TILE / 32 <= kFastThreads, "whole warps of tile groups");
    if (tid < TILE / 32) {
        // sliding maximum over q <= 31 consecutive group maxima by doubling: the warp's 32 outputs need the 32 + q - 1
        // entries after G0 + tid, held as (a, b) = elements lane and lane + 32; after the levels 1, 2, .. p/2 an
        // element is the maximum of p = 2^floor(log2 q) consecutive entries and two of those cover q
        const int* gsrc = gmax + G0 + tid + 1;
        int a = gsrc[0];
        int b = G0 + tid + 33 < nrow * 4 ? gsrc[32] : INT_MIN;   // beyond the staged groups: never part of a used window
        int m = INT_MIN;
        if (q > 0) {
            int p = 1;
            for (; 2 * p <= q; p *= 2) {
                const int ta = __shfl_down_sync(0xffffffffu, a, p);
                const int tb = __shfl_up_sync(0xffffffffu, b, 32 - p);
                const int tc = __shfl_down_sync(0xffffffffu, b, p);
                a = max(a, lane + p < 32 ? ta : tb);
                b = max(b, lane + p < 32 ? tc : b);
            }
            const int sft = q - p;                          // in [0, p)
            const int ta = __shfl_down_sync(0xffffffffu, a, sft);
            const int tb = __shfl_up_sync(0xffffffffu, b, (32 - sft) & 31);
            m = max(a, (sft == 0 || lane + sft < 32) ? ta : tb);
        }
        gM[tid] = m;
        const bool poss = 32 * tid < nvalid && !(m > gmax[G0 + tid]);
        const uint32_t possw = __ballot_sync(0xffffffffu, poss);
        int base = 0;
        if (lane == 0 && possw != 0u) base = atomicAdd(nposs_s, __popc(possw));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (poss) {
            poss_list[base + __popc(possw & ((1u << lane) - 1u))] = (unsigned short)tid;
        } else {
            cw[tid] = 0u;
            pw[tid] = 0u;
        }
    }
    __syncthreads();
    const int nposs = *nposs_s;
    // The minimum of the whole staged window settles the degenerate capture (constant input: every sample is a
    // candidate and nothing is below its threshold) and bounds the cost there; a plausible capture queues about
    // TILE / (T+1) groups and never takes this branch (CTA-uniform).
    int tmin = INT_MIN;
    if (nposs > TILE / 128) {
        int rmin = INT_MAX;
        for (int i = tid; i < nrow * 32; i += kFastThreads) {
            const int4 v = reinterpret_cast<const int4*>(z)[i];
            rmin = min(rmin, min(min(v.x, v.y), min(v.z, v.w)));
        }
        rmin = __reduce_min_sync(0xffffffffu, rmin);
        if (lane == 0) wmin[warp] = rmin;
        __syncthreads();
        tmin = __reduce_min_sync(0xffffffffu, lane < kWarps ? wmin[lane] : INT_MAX);
    }
    // ---- a warp per queued group: its candidates decided exactly, then the threshold test of each (:273-279).
    //      The warp owns both bitmap words of the group: no atomics, no second queue.
    for (int pi = warp; pi < nposs; pi += kWarps) {
        const int w = poss_list[pi];
        const int G = G0 + w;
        const int zi = z[32 * G + lane];
        const int M = gM[w];
        const bool valid = (32 * w + lane) < nvalid;
        // exclusive suffix maximum inside the group
        int s = zi;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const int t = __shfl_down_sync(0xffffffffu, s, dd);
            if (lane + dd < 32) s = max(s, t);
        }
        s = __shfl_down_sync(0xffffffffu, s, 1);
        if (lane == 31) s = INT_MIN;
        // remainder: elements 0..e of group G+q+1 (continuing into G+q+2), e = lane + d
        const int A = G + q + 1;
        const int pa = warp_incl_scan_max(z[32 * A + lane], lane);
        const int pbm = warp_incl_scan_max(z[32 * A + 32 + lane], lane);
        const int e = lane + d;
        const int ra = __shfl_sync(0xffffffffu, pa, e & 31);
        const int rb = __shfl_sync(0xffffffffu, pbm, e & 31);
        const int ga = __shfl_sync(0xffffffffu, pa, 31);
        const int rem = e < 0 ? INT_MIN : (e < 32 ? ra : max(ga, rb));
        const int fwd = max(max(s, M), rem);
        const bool cand = valid && !(fwd > zi);
        const uint32_t candw = __ballot_sync(0xffffffffu, cand);
        uint32_t passw = 0u;
        for (uint32_t left = candw; left != 0u; left &= left - 1u) {
            const int lp = __ffs(left) - 1;
            const int ic = 32 * G + lp;
            const float tvf = __fdiv_rn(__int_as_float(__shfl_sync(0xffffffffu, zi, lp)), thr);
            if (!(tvf > 0.0f)) continue;        // zpow >= 0: nothing is below a non-positive threshold
            const int tv = __float_as_int(tvf);
            if (tv <= tmin) continue;           // nothing in the window is below the threshold
            const int w_lo = ic - T, w_hi = ic + T;
            const int g_first = w_lo >> 5, g_last = w_hi >> 5;   // at most 65 groups
            int cnt = 0;
            uint32_t need[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int gi = g_first + 32 * r + lane;
                bool nd = false;
                if (gi <= g_last) {
                    if (gmax[gi] < tv) cnt += min(32 * gi + 31, w_hi) - max(32 * gi, w_lo) + 1;
                    else nd = true;
                }
                need[r] = __ballot_sync(0xffffffffu, nd);
            }
            if (__popc(need[0]) + __popc(need[1]) + __popc(need[2]) > 6) {
                // many groups reach the threshold — the usual noise maximum, whose threshold sits near the median of
                // the window: count the 2T+1 samples directly, lane-strided (conflict-free, no alignment, no masks; a
                // 128-bit variant with masked end chunks measured slower).
                // Sign bit of the difference of two bit patterns = the compare (no overflow: see the header).
                const int* zw = z + w_lo + lane;
                const int nfull = tv_need >> 5;
                unsigned c0 = 0u, c1 = 0u;
                const unsigned utv = (unsigned)tv;   // unsigned wrap-around arithmetic keeps the compiler from turning it back into compares
                int k = 0;
#pragma unroll 4
                for (; k + 1 < nfull; k += 2) {
                    c0 += ((unsigned)zw[32 * k] - utv) >> 31;
                    c1 += ((unsigned)zw[32 * k + 32] - utv) >> 31;
                }
                if (k < nfull) c0 += ((unsigned)zw[32 * k] - utv) >> 31;
                if (lane < (tv_need & 31)) c1 += ((unsigned)zw[32 * nfull] - utv) >> 31;
                cnt = (int)(c0 + c1);
            } else {
                const int nchunk = ((g_last - g_first) >> 2) + 1;   // 4 groups = 128 samples per step
                const unsigned span = 2u * (unsigned)T;
                const int* zc0 = z + 32 * g_first + 4 * lane;
                const int o0 = 32 * g_first + 4 * lane - w_lo;      // window offset of this lane's first sample
                const uint32_t mybit = 1u << (lane >> 3);
                for (int ch = 0; ch < nchunk; ++ch) {
                    const uint32_t word = (ch >> 3) == 0 ? need[0] : ((ch >> 3) == 1 ? need[1] : need[2]);
                    const uint32_t nib = (word >> (4 * (ch & 7))) & 0xfu;
                    if (nib & mybit) {
                        const int4 x = *reinterpret_cast<const int4*>(zc0 + 128 * ch);
                        const int o = o0 + 128 * ch;
                        cnt += ((unsigned)(o + 0) <= span && x.x < tv) ? 1 : 0;
                        cnt += ((unsigned)(o + 1) <= span && x.y < tv) ? 1 : 0;
                        cnt += ((unsigned)(o + 2) <= span && x.z < tv) ? 1 : 0;
                        cnt += ((unsigned)(o + 3) <= span && x.w < tv) ? 1 : 0;
                    }
                }
            }
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (2 * cnt >= tv_need) passw |= 1u << lp;
        }
        if (lane == 0) {
            cw[w] = candw;
            pw[w] = passw;
        }
    }
}

// ---------------------------------------------------------------------------------
// Flags from the correlator's group extrema (T in [32, 1023], stride S in [64, 2016]).
//
// The correlator leaves, per FFT block, (max, min) of every GROUP of the metric: group 0 = the block's first
// sample, group q >= 1 = samples 32q-31 .. 32q of the block (clipped to S) — b200sync_internal.h.  Groups tile the
// stream without gaps, in order, with lengths 1, 32, ..., 32, (S-1) % 32.  This kernel decides the candidate and
// threshold bitmaps of [lo, hi) reading those 8 bytes per group, and the samples themselves only where the
// extrema cannot decide:
//   * a group can hold a candidate only if its maximum is >= the maximum M of the groups that lie entirely
//     inside the forward window (p, p+T] of EVERY sample of the group (about T/32 - 1 groups): one compare per group
//     rejects ~ (1 - 32/T) of them.  The others load their <= 32 samples and the <= 62 samples between the last whole
//     group and p+T, and decide every sample exactly (suffix maximum | M | prefix maximum of the remainder);
//   * the threshold test of a candidate (count of window samples below zpow[p]/thr, :273-279) gets a lower bound
//     from the groups entirely below the threshold and an upper bound from the groups entirely not below; almost
//     every candidate is decided by the bounds (a real peak: nearly everything is below; a noise maximum: nearly
//     nothing is), only the rest counts the undecided groups sample by sample.
// DRAM traffic: 0.25 B/sample of extrema + the samples of ~1 group in T/32, instead of 4 B/sample.
// The bitmaps are pre-zeroed; bits are set with atomicOr (groups are not aligned to bitmap words).
// ---------------------------------------------------------------------------------
constexpr int kGmThreads = 256;
constexpr int kGmGroups = 512;    // groups decided per CTA
constexpr int kGmAhead = 40;      // whole groups a forward window can hold: <= 1023/32 + 2

struct GmGeom {
    const float2* gm;        // rows of the first channel: kGmF2PerGroup float2 per group, block b at (b - b0) * ng groups
    long long b0;            // block of row 0
    long long n_blocks;      // blocks available
    long long chan_stride;   // float2 between channels
    int S, ng;
    unsigned inv_s, inv_ng;  // ceil(2^32 / S), ceil(2^32 / ng): exact quotients for arguments below 2^16
};
// positions and groups RELATIVE to the CTA's base block (all below 2^16): 32-bit arithmetic, no divisions
__device__ __forceinline__ int gm_group_of(int x, int S, int ng, unsigned inv_s) {
    const int b = (int)__umulhi((unsigned)x, inv_s);
    const int kk = x - b * S;
    return b * ng + ((kk + 31) >> 5);
}
__device__ __forceinline__ void gm_span(int g, int S, int ng, unsigned inv_ng, int& start, int& len, int& q) {
    const int b = (int)__umulhi((unsigned)g, inv_ng);
    q = g - b * ng;
    const int k0 = q ? 32 * q - 31 : 0;
    start = b * S + k0;
    len = q ? min(32, S - k0) : 1;
}

__global__ void __launch_bounds__(kGmThreads)
peak_flags_gm_kernel(const float* __restrict__ zpow, long long z_base, long long z_end, GmGeom gg, long long lo,
                     long long hi, int T, float thr, uint32_t* __restrict__ cand_bits,
                     uint32_t* __restrict__ pass_bits, PeakBatch pb) {
    __shared__ int gmax_s[kGmGroups + kGmAhead];
    zpow += (long long)blockIdx.y * pb.z_stride;
    cand_bits = ws_at(cand_bits, pb.ws_stride);
    pass_bits = ws_at(pass_bits, pb.ws_stride);
    const float4* gm4 = reinterpret_cast<const float4*>(gg.gm + (long long)blockIdx.y * gg.chan_stride);   // 2 float4 per group
    const int S = gg.S, ng = gg.ng;
    const unsigned inv_s = gg.inv_s, inv_ng = gg.inv_ng;
    // ---- the CTA's frame: base block B0 such that every sample it can reference lies at or after B0 * S
    const long long b_lo = lo / S;
    const long long G_lo = b_lo * ng + (((int)(lo - b_lo * S) + 31) >> 5);
    const long long Gc0 = G_lo + (long long)blockIdx.x * kGmGroups;      // first group of the CTA (absolute)
    const long long bc = Gc0 / ng;                                        // its block
    const long long B0 = bc - (T + S - 1) / S > 0 ? bc - (T + S - 1) / S : 0;
    const long long P0 = B0 * S, G0 = B0 * ng;
    const int g_c0 = (int)(Gc0 - G0);                                     // relative group of the CTA's first group
    const int x_lo = (int)max(lo - P0, -1000000LL), x_hi = (int)min(hi - P0, 1000000LL);   // [lo, hi) in the frame
    const int x_zend = (int)min(z_end - P0, 1000000LL);
    const long long bit0 = P0 - lo;                                       // bitmap bit of frame position x: x + bit0
    const int* zb = reinterpret_cast<const int*>(zpow) + (P0 - z_base);   // zb[x] = bits of zpow at frame position x
    const long long g_first = gg.b0 * ng - G0, g_last = (gg.b0 + gg.n_blocks) * ng - 1 - G0;   // relative groups with a row
    const float4* row0 = gm4 + (G0 - gg.b0 * ng) * 2;                     // row of relative group 0
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kGmGroups + kGmAhead; i += kGmThreads) {
        const long long g = (long long)g_c0 + i;
        gmax_s[i] = (g >= g_first && g <= g_last) ? __float_as_int(__ldg(&row0[2 * g].x)) : INT_MIN;
    }
    __syncthreads();
    const int total = 2 * T + 1;
    const int g_hi = gm_group_of(x_hi - 1, S, ng, inv_s);                 // last group that touches [lo, hi)
    for (int r = 0; r < kGmGroups / (kGmThreads / 32) / 32; ++r) {
        const int i = (warp * (kGmGroups / (kGmThreads / 32))) + r * 32 + lane;   // this lane's group inside the CTA
        const int g = g_c0 + i;
        int st = 0, ln = 0, q = 0, M = INT_MIN, r0 = 0;
        bool poss = false;
        if (g <= g_hi) {
            gm_span(g, S, ng, inv_ng, st, ln, q);
            // last group that ends at or before st + T (inside the window of the group's FIRST sample, hence of all)
            const int pe = st + T;
            int ge = gm_group_of(pe, S, ng, inv_s), se, le, qe;
            gm_span(ge, S, ng, inv_ng, se, le, qe);
            if (se + le - 1 != pe) {
                --ge;
                gm_span(ge, S, ng, inv_ng, se, le, qe);
            }
            r0 = se + le;                          // first sample after the last whole group
            for (int k = g + 1; k <= ge; ++k) M = max(M, gmax_s[k - g_c0]);
            poss = st < x_hi && st + ln > x_lo && !(M > gmax_s[i]);
        }
        uint32_t possb = __ballot_sync(0xffffffffu, poss);
        while (possb != 0u) {                      // warp-uniform: one possible group at a time, a lane per sample
            const int src = __ffs(possb) - 1;
            possb &= possb - 1;
            const int gst = __shfl_sync(0xffffffffu, st, src);
            const int gln = __shfl_sync(0xffffffffu, ln, src);
            const int gM = __shfl_sync(0xffffffffu, M, src);
            const int gr0 = __shfl_sync(0xffffffffu, r0, src);
            const int x = gst + lane;
            const bool in_group = lane < gln;
            const int zi = in_group ? __ldg(zb + x) : INT_MIN;
            // remainder: samples gr0 .. x+T (at most 62 of them)
            const int ia = gr0 + lane, ib = gr0 + 32 + lane;
            const int za = ia < x_zend ? __ldg(zb + ia) : INT_MIN;
            const int zc = ib < x_zend ? __ldg(zb + ib) : INT_MIN;
            // exclusive suffix maximum inside the group
            int sfx = zi;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {
                const int t = __shfl_down_sync(0xffffffffu, sfx, dd);
                if (lane + dd < 32) sfx = max(sfx, t);
            }
            sfx = __shfl_down_sync(0xffffffffu, sfx, 1);
            if (lane == 31) sfx = INT_MIN;
            const int pa = warp_incl_scan_max(za, lane);
            const int pb2 = warp_incl_scan_max(zc, lane);
            const int e = x + T - gr0;
            const int ra = __shfl_sync(0xffffffffu, pa, e & 31);
            const int rb = __shfl_sync(0xffffffffu, pb2, e & 31);
            const int ga = __shfl_sync(0xffffffffu, pa, 31);
            const int rem = e < 0 ? INT_MIN : (e < 32 ? ra : max(ga, rb));
            const bool cand = in_group && x >= x_lo && x < x_hi && !(max(max(sfx, gM), rem) > zi);
            if (cand) atomicOr(cand_bits + ((x + bit0) >> 5), 1u << ((x + bit0) & 31));
            uint32_t candw = __ballot_sync(0xffffffffu, cand);
            // ---- threshold test of every candidate of the group (:273-279)
            while (candw != 0u) {
                const int cl = __ffs(candw) - 1;
                candw &= candw - 1;
                const int xc = gst + cl;
                const float tvf = __fdiv_rn(__int_as_float(__shfl_sync(0xffffffffu, zi, cl)), thr);
                if (!(tvf > 0.0f)) continue;       // zpow >= 0: nothing is below a non-positive threshold
                const int tv = __float_as_int(tvf);
                const int w_hi = xc + T;
                int w_lo = xc - T;
                int below = 0, notbelow = 0;
                if (w_lo < 0) {                    // only with P0 = 0: the zero-initialised history before the stream, 0 < tv
                    if (lane == 0) below = -w_lo;
                    w_lo = 0;
                }
                const int gwa = gm_group_of(w_lo, S, ng, inv_s), gwb = gm_group_of(w_hi, S, ng, inv_s);
                uint32_t need[3];
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) {
                    const int gi = gwa + lane + 32 * rr;
                    bool nd = false;
                    if (gi <= gwb) {
                        int si, li, qi;
                        gm_span(gi, S, ng, inv_ng, si, li, qi);
                        if (si >= w_lo && si + li - 1 <= w_hi && gi >= g_first && gi <= g_last) {
                            const float4 mx = __ldg(row0 + 2 * gi);
                            if (__float_as_int(mx.x) < tv) {
                                below += li;
                            } else {
                                // quarter minima, ascending: quarter j covers samples 8j .. 8j+7 of the group (group
                                // 0, one sample, sits in quarter 3)
                                const float4 mn = __ldg(row0 + 2 * gi + 1);
                                const int m4[4] = {__float_as_int(mn.x), __float_as_int(mn.y), __float_as_int(mn.z),
                                                   __float_as_int(mn.w)};
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const int sl = qi ? max(0, min(8, li - 8 * j)) : (j == 3 ? 1 : 0);
                                    if (sl > 0) {
                                        if (!(m4[j] < tv)) notbelow += sl;
                                        else nd = true;
                                    }
                                }
                            }
                        } else {
                            nd = true;
                        }
                    }
                    need[rr] = __ballot_sync(0xffffffffu, nd);
                }
                below = __reduce_add_sync(0xffffffffu, below);
                notbelow = __reduce_add_sync(0xffffffffu, notbelow);
                bool pass;
                if (2 * below >= total) pass = true;
                else if (2 * (total - notbelow) < total) pass = false;
                else {                             // the bounds do not decide: count the undecided groups exactly
                    int cnt = 0;
#pragma unroll
                    for (int rr = 0; rr < 3; ++rr) {
                        uint32_t nb = need[rr];
                        while (nb != 0u) {
                            const int gl = __ffs(nb) - 1;
                            nb &= nb - 1;
                            int si, li, qi;
                            gm_span(gwa + gl + 32 * rr, S, ng, inv_ng, si, li, qi);
                            const int xq = si + lane;
                            if (lane < li && xq >= w_lo && xq <= w_hi && xq < x_zend && __ldg(zb + xq) < tv) ++cnt;
                        }
                    }
                    cnt = __reduce_add_sync(0xffffffffu, cnt);
                    pass = 2 * (below + cnt) >= total;
                }
                if (pass && lane == 0) atomicOr(pass_bits + ((xc + bit0) >> 5), 1u << ((xc + bit0) & 31));
            }
        }
    }
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
                    long long nseg, uint16_t* __restrict__ tables, uint32_t* __restrict__ segflag, PeakBatch pb) {
    cand_bits = ws_at(cand_bits, pb.ws_stride);
    tables = ws_at(tables, pb.ws_stride);
    segflag = ws_at(segflag, pb.ws_stride);
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
    // the walk is a chain of dependent shuffles per piece; the bitmap words of the next TWO pieces are already on
    // their way (a piece per iteration would expose one L2/DRAM latency each: 0.37 -> 0.2 ms at 2^30 samples)
    auto piece_len = [&](long long f) { return (int)max(0LL, min((long long)Fr, range - f * Fr)); };
    const long long f_first = seg * (long long)M;
    uint32_t w1 = piece_word(cand_bits, f_first * Fr, piece_len(f_first), lane);
    uint32_t w2 = f_first + 1 < f_end ? piece_word(cand_bits, (f_first + 1) * Fr, piece_len(f_first + 1), lane) : 0u;
    for (long long f = f_first; f < f_end; ++f) {
        const int L = piece_len(f);
        const uint32_t w = w1;
        w1 = w2;
        w2 = f + 2 < f_end ? piece_word(cand_bits, (f + 2) * Fr, piece_len(f + 2), lane) : 0u;
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
                  uint16_t* __restrict__ seg_jin, uint16_t* __restrict__ range_table, PeakBatch pb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    tables = ws_at(tables, pb.ws_stride);
    segflag = ws_at(segflag, pb.ws_stride);
    if (seg_jin != nullptr) seg_jin = ws_at(seg_jin, pb.ws_stride);
    if (state != nullptr) state += blockIdx.y;
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
// Each warp re-walks its segment from the now-known entry state.  A piece examines at most one
// candidate, so segment `seg` owns M slots; detections land there in increasing order and
// det_offsets/det_gather compact the slots into one SORTED list (no host-side sort).
__global__ void __launch_bounds__(128)
chain_emit_kernel(const uint32_t* __restrict__ cand_bits, const uint32_t* __restrict__ pass_bits,
                  long long range, int T, int M, long long nfr, long long nseg,
                  const uint16_t* __restrict__ seg_jin, long long lo, unsigned long long* __restrict__ slots,
                  uint32_t* __restrict__ seg_count, PeakBatch pb) {
    cand_bits = ws_at(cand_bits, pb.ws_stride);
    pass_bits = ws_at(pass_bits, pb.ws_stride);
    seg_jin = ws_at(seg_jin, pb.ws_stride);
    slots = ws_at(slots, pb.ws_stride);
    seg_count = ws_at(seg_count, pb.ws_stride);
    const int lane = threadIdx.x & 31;
    const long long seg = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (seg >= nseg) return;
    const int Fr = T + 1;
    int j = seg_jin[seg];
    unsigned int cnt = 0;
    const long long f0 = seg * (long long)M;
    const long long f_end = min(nfr, (seg + 1) * (long long)M);
    int L = (int)min((long long)Fr, range - f0 * Fr);
    uint32_t w = piece_word(cand_bits, f0 * Fr, L, lane);
    uint32_t pw = piece_word(pass_bits, f0 * Fr, L, lane);   // the threshold bits of the piece travel with its candidates
    for (long long f = f0; f < f_end; ++f) {
        const long long fstart = f * Fr;
        const int Lc = L;
        uint32_t wc = w;
        const uint32_t pc = pw;
        if (f + 1 < f_end) {  // prefetch the next piece: the walk itself is a chain of dependent shuffles
            L = (int)min((long long)Fr, range - (fstart + Fr));
            w = piece_word(cand_bits, fstart + Fr, L, lane);
            pw = piece_word(pass_bits, fstart + Fr, L, lane);
        }
        const int off = 32 * lane;
        if (off + 31 < j) wc = 0u;
        else if (off < j) wc &= 0xffffffffu << (j - off);
        const uint32_t m = __ballot_sync(0xffffffffu, wc != 0u);
        if (j < Lc && m != 0u) {
            const int l0 = __ffs(m) - 1;
            const uint32_t ww = __shfl_sync(0xffffffffu, wc, l0);
            const int a = 32 * l0 + __ffs(ww) - 1;
            const long long bit = fstart + a;
            if ((__shfl_sync(0xffffffffu, pc, l0) >> (a & 31)) & 1u) {  // warp-uniform
                if (lane == 0) slots[seg * (long long)M + cnt] = (unsigned long long)(lo + bit);
                ++cnt;
            }
            j = a + Fr - Lc;
        } else {
            j = (j >= Lc) ? (j - Lc) : 0;
        }
    }
    if (lane == 0) seg_count[seg] = cnt;
}

// exclusive scan of the per-segment detection counts (one CTA; nseg <= 32 * num_sms + 1)
__global__ void __launch_bounds__(1024)
det_offsets_kernel(const uint32_t* __restrict__ seg_count, int nseg, uint32_t* __restrict__ seg_off,
                   PeakState* __restrict__ state, PeakBatch pb) {
    seg_count = ws_at(seg_count, pb.ws_stride);
    seg_off = ws_at(seg_off, pb.ws_stride);
    state += blockIdx.y;
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) carry_s = 0u;
    __syncthreads();
    for (int base = 0; base < nseg; base += 1024) {
        const int i = base + t;
        const uint32_t v = i < nseg ? seg_count[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t wt = warp_tot[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, wt, d);
                if (lane >= d) wt += y;
            }
            warp_tot[lane] = wt;  // inclusive
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t before = carry + (warp ? warp_tot[warp - 1] : 0u) + (x - v);
        if (i < nseg) seg_off[i] = before;
        __syncthreads();
        if (t == 1023) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    if (t == 0) {
        seg_off[nseg] = carry_s;
        state->det_count = carry_s;
    }
}

__global__ void __launch_bounds__(128)
det_gather_kernel(const unsigned long long* __restrict__ slots, const uint32_t* __restrict__ seg_off, int M,
                  long long nseg, unsigned long long* __restrict__ det_idx, unsigned int det_cap, PeakBatch pb) {
    slots = ws_at(slots, pb.ws_stride);
    seg_off = ws_at(seg_off, pb.ws_stride);
    det_idx += (size_t)blockIdx.y * pb.det_stride;
    const int lane = threadIdx.x & 31;
    const long long seg = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (seg >= nseg) return;
    const uint32_t o = seg_off[seg], n = seg_off[seg + 1] - o;
    for (uint32_t i = lane; i < n; i += 32)
        if (o + i < det_cap) det_idx[o + i] = slots[seg * (long long)M + i];
}

// ---------------------------------------------------------------------------------
// Streaming-sized ranges (one processBulk span: tens of thousands of samples, at most a few dozen examined peaks):
// the walk of PM/syncword_detection.hpp:267-298 over the candidate / threshold bitmaps by ONE warp, in order — 32
// bitmap words per probe, a jump of T+1 after every examined peak.  Replaces chain_tables + chain_scan + chain_emit +
// det_offsets + det_gather (five launches whose parallelism only pays on captures of many millions of samples).
constexpr int kSmallThreads = 256;
constexpr long long kWalkPiece = 1LL << 19;   // bitmap bits staged in shared memory at a time (2 x 64 KiB)
__global__ void __launch_bounds__(kSmallThreads)
chain_small_kernel(const uint32_t* __restrict__ cand_bits, const uint32_t* __restrict__ pass_bits, long long range,
                   int T, long long lo, long long hi, PeakState* __restrict__ state,
                   unsigned long long* __restrict__ det_idx, unsigned int det_cap, PeakBatch pb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cand_bits = ws_at(cand_bits, pb.ws_stride);
    pass_bits = ws_at(pass_bits, pb.ws_stride);
    state += blockIdx.y;
    det_idx += (size_t)blockIdx.y * pb.det_stride;
    // The walk is a chain of DEPENDENT probes: from global memory every probe would cost an L2 round trip, so the
    // bitmaps go through shared memory, a piece of 2^19 bits at a time (coalesced, all threads), and one warp walks
    // each piece.  Longer ranges than one piece only occur where the parallel chain kernels do not apply
    // (time_threshold > 1023).
    const int lane = threadIdx.x & 31;
    const unsigned long long r = state->r_abs;
    long long j = (r > (unsigned long long)lo) ? (long long)(r - (unsigned long long)lo) : 0;  // search offset in the range
    unsigned int cnt = 0;
    const int piece_words = (int)(min(range, kWalkPiece) + 31) >> 5;
    uint32_t* cand_s = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* pass_s = cand_s + piece_words;
    for (long long p0 = 0; p0 < range; p0 += kWalkPiece) {
        const long long plen = min(kWalkPiece, range - p0);
        const int nwords = (int)((plen + 31) >> 5);
        const long long w0 = p0 >> 5;   // kWalkPiece is a multiple of 32
        __syncthreads();
        for (int i = threadIdx.x; i < nwords; i += kSmallThreads) {
            cand_s[i] = cand_bits[w0 + i];
            pass_s[i] = pass_bits[w0 + i];
        }
        __syncthreads();
        if (threadIdx.x < 32 && j < p0 + plen) {
            const long long jr = j > p0 ? j - p0 : 0;
            const long long je = peak_walk_warp(cand_s, pass_s, nwords, plen, T, jr, [&](long long found) {
                if (lane == 0 && cnt < det_cap) det_idx[cnt] = (unsigned long long)(lo + p0 + found);
                ++cnt;
            });
            j = p0 + je;
        }
        // (j and cnt live in warp 0 only; the other warps just stage)
    }
    if (threadIdx.x == 0) {
        const long long r_end = lo + j;
        state->r_abs = (unsigned long long)(r_end > hi ? r_end : hi);
        state->det_count = cnt;
    }
}

// ---------------------------------------------------------------------------------
// cudaFuncSetAttribute costs microseconds per call and the streaming path launches these kernels per span: remember
// the largest size already granted per (function, device)
static cudaError_t set_smem_attr(const void* fn, size_t bytes) {
    struct Slot { const void* fn; int dev; size_t bytes; };
    static Slot slots[64] = {};
    static int nslots = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    for (int i = 0; i < nslots; ++i)
        if (slots[i].fn == fn && slots[i].dev == dev) {
            if (bytes <= slots[i].bytes) return cudaSuccess;
            e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            if (e == cudaSuccess) slots[i].bytes = bytes;
            return e;
        }
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess && nslots < 64) slots[nslots++] = Slot{fn, dev, bytes};  // benign race: worst case a repeated call
    return e;
}

// candidate + threshold bitmaps of [lo, hi) (the per-sample part of a6); nch channels side by side (blockIdx.y)
static cudaError_t launch_flags(const float* d_zpow, long long z_base, long long z_end, long long lo, long long hi,
                                int T, float power_threshold, const PeakPlan& pl, unsigned char* ws, int nch,
                                const PeakBatch& pb, cudaStream_t st, bool zero_padding = true,
                                const GmGeom* gg = nullptr, bool pdl = false) {
    const long long range = hi - lo;
    cudaError_t e;
    if (gg != nullptr && gg->gm != nullptr) {
        // from the correlator's group extrema: bitmaps zeroed, bits set atomically
        for (int c = 0; c < nch; ++c) {
            e = cudaMemsetAsync(ws + (size_t)c * pb.ws_stride + pl.off_cand, 0, sizeof(uint32_t) * pl.nwords, st);
            if (e != cudaSuccess) return e;
            e = cudaMemsetAsync(ws + (size_t)c * pb.ws_stride + pl.off_pass, 0, sizeof(uint32_t) * pl.nwords, st);
            if (e != cudaSuccess) return e;
        }
        const long long b_lo = lo / gg->S, b_hi = (hi - 1) / gg->S;
        const long long ngroups = (b_hi - b_lo + 1) * gg->ng;   // upper bound of the groups touching [lo, hi)
        const dim3 grid((unsigned)((ngroups + kGmGroups - 1) / kGmGroups), (unsigned)nch);
        peak_flags_gm_kernel<<<grid, kGmThreads, 0, st>>>(d_zpow, z_base, z_end, *gg, lo, hi, T, power_threshold,
                                                         reinterpret_cast<uint32_t*>(ws + pl.off_cand),
                                                         reinterpret_cast<uint32_t*>(ws + pl.off_pass), pb);
        count_launch();
        return cudaGetLastError();
    }
    const bool fast = T >= 32 && T <= kMaxTimeThreshold;   // beyond: the generic (van Herk) kernel, any T
    const bool small = fast && range <= (1LL << 18);
    const int tile = fast ? (small ? kFastTileSmall : kFastTile) : kFlagsTile;
    const long long ntiles = (range + tile - 1) / tile;
    // zero the padding words past the last tile (tiles write every word they own)
    const long long written = ntiles * (tile / 32);
    for (int c = 0; zero_padding && c < nch && written < pl.nwords; ++c) {
        uint32_t* cand = reinterpret_cast<uint32_t*>(ws + (size_t)c * pb.ws_stride + pl.off_cand);
        uint32_t* pass = reinterpret_cast<uint32_t*>(ws + (size_t)c * pb.ws_stride + pl.off_pass);
        e = cudaMemsetAsync(cand + written, 0, sizeof(uint32_t) * (pl.nwords - written), st);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(pass + written, 0, sizeof(uint32_t) * (pl.nwords - written), st);
        if (e != cudaSuccess) return e;
    }
    uint32_t* cand = reinterpret_cast<uint32_t*>(ws + pl.off_cand);
    uint32_t* pass = reinterpret_cast<uint32_t*>(ws + pl.off_pass);
    const dim3 grid((unsigned)ntiles, (unsigned)nch);
    if (fast) {
        const FastGeom geo = fast_geom(T, tile);
        const size_t smem = sizeof(int) * ((size_t)geo.nrow * 128 + (size_t)geo.nrow * 4 + tile / 32 + kFastThreads / 32 + 1) +
                            sizeof(unsigned short) * (tile / 32);
        auto kern = small ? peak_flags_kernel<kFastTileSmall> : peak_flags_kernel<kFastTile>;
        e = set_smem_attr((const void*)kern, smem);
        if (e != cudaSuccess) return e;
        if (pdl) {
            e = launch_pdl(kern, grid, dim3(kFastThreads), smem, st, d_zpow, z_base, z_end, lo, hi, T, power_threshold, cand,
                           pass, pb);
            if (e != cudaSuccess) return e;
        } else {
            kern<<<grid, kFastThreads, smem, st>>>(d_zpow, z_base, z_end, lo, hi, T, power_threshold, cand, pass, pb);
        }
        count_launch();
    } else {
        const int n = kFlagsTile + 2 * T;
        const size_t smem = sizeof(float) * 3 * n + sizeof(uint32_t) * (kFlagsTile / 32) +
                            sizeof(unsigned short) * kFlagsTile;
        e = set_smem_attr((const void*)peak_flags_generic_kernel, smem);
        if (e != cudaSuccess) return e;
        peak_flags_generic_kernel<<<grid, kFlagsThreads, smem, st>>>(d_zpow, z_base, z_end, lo, hi, T,
                                                                     power_threshold, cand, pass, pb);
        count_launch();
    }
    return cudaGetLastError();
}

size_t peak_plan_bytes(long long range, int T, int num_sms) { return make_plan(range, T, num_sms).total; }

// nch > 1: batched channel mode — d_zpow, d_ws, d_state, d_det_idx are the first channel's; the others follow at
// z_stride floats / ws_stride bytes / one PeakState / det_stride entries
cudaError_t launch_peak_phase1(const float* d_zpow, long long z_base, long long z_end, long long lo,
                               long long hi, int T, float power_threshold, void* d_ws, size_t ws_bytes,
                               uint16_t* d_range_table, int num_sms, cudaStream_t st, int nch, long long z_stride,
                               size_t ws_stride, const float2* d_gm, long long gm_b0, long long gm_blocks, int S,
                               long long gm_chan_stride) {
    const long long range = hi - lo;
    if (range <= 0) return cudaSuccess;
    const PeakPlan pl = make_plan(range, T, num_sms);
    if (pl.total > ws_bytes) return cudaErrorInvalidValue;
    if (nch > 1 && (d_range_table != nullptr || ws_stride < pl.total)) return cudaErrorInvalidValue;
    const PeakBatch pb{z_stride, ws_stride, 0};
    unsigned char* ws = static_cast<unsigned char*>(d_ws);
    uint32_t* cand = reinterpret_cast<uint32_t*>(ws + pl.off_cand);
    uint16_t* tables = reinterpret_cast<uint16_t*>(ws + pl.off_tables);
    uint32_t* segflag = reinterpret_cast<uint32_t*>(ws + pl.off_flag);
    const int ngr = S > 0 ? gm_groups_per_block(S) : 1;
    GmGeom gg{d_gm, gm_b0, gm_blocks, gm_chan_stride, S, ngr,
              S > 0 ? (unsigned)((0x100000000ULL + S - 1) / S) : 0u, (unsigned)((0x100000000ULL + ngr - 1) / ngr)};
    const bool use_gm = d_gm != nullptr && gm_supported(S, T);
    cudaError_t e = launch_flags(d_zpow, z_base, z_end, lo, hi, T, power_threshold, pl, ws, nch, pb, st, true,
                                 use_gm ? &gg : nullptr);
    if (e != cudaSuccess) return e;
    chain_tables_kernel<<<dim3((unsigned)((pl.nseg + 3) / 4), (unsigned)nch), 128, 0, st>>>(cand, range, T, pl.M, pl.nfr,
                                                                                           pl.nseg, tables, segflag, pb);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (d_range_table != nullptr) {
        const size_t ssm = sizeof(uint32_t) * (size_t)pl.nseg + sizeof(uint16_t) * (size_t)(pl.nseg + 2 + T + 2);
        e = set_smem_attr((const void*)chain_scan_kernel, ssm);
        if (e != cudaSuccess) return e;
        chain_scan_kernel<<<1, kScanThreads, ssm, st>>>(tables, segflag, pl.nseg, T, 0, -1, nullptr, lo, hi,
                                                        nullptr, d_range_table, pb);
        count_launch();
        e = cudaGetLastError();
    }
    return e;
}

// both phases for a streaming-sized range: two launches (flags, in-order walk) instead of seven
cudaError_t launch_peak_stream(const float* d_zpow, long long z_base, long long z_end, long long lo, long long hi,
                               int T, float power_threshold, void* d_ws, size_t ws_bytes, PeakState* d_state,
                               unsigned long long* d_det_idx, unsigned int det_cap, int num_sms, cudaStream_t st) {
    const long long range = hi - lo;
    if (range <= 0) return cudaSuccess;
    const PeakPlan pl = make_plan(range, T, num_sms);
    if (pl.total > ws_bytes) return cudaErrorInvalidValue;
    const PeakBatch pb{0, 0, 0};
    unsigned char* ws = static_cast<unsigned char*>(d_ws);
    uint32_t* cand = reinterpret_cast<uint32_t*>(ws + pl.off_cand);
    uint32_t* pass = reinterpret_cast<uint32_t*>(ws + pl.off_pass);
    cudaError_t e = launch_flags(d_zpow, z_base, z_end, lo, hi, T, power_threshold, pl, ws, 1, pb, st);
    if (e != cudaSuccess) return e;
    const size_t ssm = 2 * sizeof(uint32_t) * (size_t)(((range < kWalkPiece ? range : kWalkPiece) + 31) >> 5);
    e = set_smem_attr((const void*)chain_small_kernel, ssm);
    if (e != cudaSuccess) return e;
    chain_small_kernel<<<1, kSmallThreads, ssm, st>>>(cand, pass, range, T, lo, hi, d_state, d_det_idx, det_cap, pb);
    count_launch();
    return cudaGetLastError();
}

// flags only; the walk runs fused in front of the refine stage (it reads words < (range+31)/32 only: no padding needed)
cudaError_t launch_peak_flags_stream(const float* d_zpow, long long z_base, long long z_end, long long lo, long long hi,
                                     int T, float power_threshold, void* d_ws, size_t ws_bytes, int num_sms,
                                     cudaStream_t st, StreamWalk* walk) {
    const long long range = hi - lo;
    if (range <= 0) return cudaErrorInvalidValue;
    const PeakPlan pl = make_plan(range, T, num_sms);
    if (pl.total > ws_bytes) return cudaErrorInvalidValue;
    const PeakBatch pb{0, 0, 0};
    unsigned char* ws = static_cast<unsigned char*>(d_ws);
    walk->cand = reinterpret_cast<uint32_t*>(ws + pl.off_cand);
    walk->pass = reinterpret_cast<uint32_t*>(ws + pl.off_pass);
    walk->range = range;
    walk->lo = lo;
    walk->hi = hi;
    walk->T = T;
    return launch_flags(d_zpow, z_base, z_end, lo, hi, T, power_threshold, pl, ws, 1, pb, st, false, nullptr, true);
}

cudaError_t launch_peak_phase2(long long lo, long long hi, int T, void* d_ws, size_t ws_bytes, int j_in,
                               PeakState* d_state, unsigned long long* d_det_idx, unsigned int det_cap,
                               int num_sms, cudaStream_t st, int nch, size_t ws_stride, size_t det_stride) {
    const long long range = hi - lo;
    if (range <= 0) return cudaSuccess;
    const PeakPlan pl = make_plan(range, T, num_sms);
    if (pl.total > ws_bytes) return cudaErrorInvalidValue;
    if (nch > 1 && ws_stride < pl.total) return cudaErrorInvalidValue;
    const PeakBatch pb{0, ws_stride, det_stride};
    unsigned char* ws = static_cast<unsigned char*>(d_ws);
    uint32_t* cand = reinterpret_cast<uint32_t*>(ws + pl.off_cand);
    uint32_t* pass = reinterpret_cast<uint32_t*>(ws + pl.off_pass);
    uint16_t* tables = reinterpret_cast<uint16_t*>(ws + pl.off_tables);
    uint16_t* jin = reinterpret_cast<uint16_t*>(ws + pl.off_jin);
    const uint32_t* segflag = reinterpret_cast<const uint32_t*>(ws + pl.off_flag);
    const size_t ssm = sizeof(uint32_t) * (size_t)pl.nseg + sizeof(uint16_t) * (size_t)(pl.nseg + 2 + T + 2);
    cudaError_t e = set_smem_attr((const void*)chain_scan_kernel, ssm);
    if (e != cudaSuccess) return e;
    chain_scan_kernel<<<dim3(1, (unsigned)nch), kScanThreads, ssm, st>>>(tables, segflag, pl.nseg, T, 1, j_in, d_state,
                                                                        lo, hi, jin, nullptr, pb);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    unsigned long long* slots = reinterpret_cast<unsigned long long*>(ws + pl.off_slots);
    uint32_t* segcnt = reinterpret_cast<uint32_t*>(ws + pl.off_segcnt);
    uint32_t* segoff = reinterpret_cast<uint32_t*>(ws + pl.off_segoff);
    const dim3 gseg((unsigned)((pl.nseg + 3) / 4), (unsigned)nch);
    chain_emit_kernel<<<gseg, 128, 0, st>>>(cand, pass, range, T, pl.M, pl.nfr, pl.nseg, jin, lo, slots, segcnt, pb);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    det_offsets_kernel<<<dim3(1, (unsigned)nch), 1024, 0, st>>>(segcnt, (int)pl.nseg, segoff, d_state, pb);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    det_gather_kernel<<<gseg, 128, 0, st>>>(slots, segoff, pl.M, pl.nseg, d_det_idx, det_cap, pb);
    count_launch();
    return cudaGetLastError();
}

}  // namespace b200sync

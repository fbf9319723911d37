// peak_walk.cuh — the in-order walk of the peak detector (PM/syncword_detection.hpp:267-298) over the candidate /
// threshold bitmaps of a streaming-sized range, by ONE warp, from shared memory.
//   candidate bit p: no larger power within the next T samples;  pass bit p: the median count test holds for p.
//   From search offset j: the next examined peak is the first candidate >= j; it is a detection iff its pass bit
//   is set; the search resumes T+1 after it.
// Used by chain_small_kernel (peaks.cu) and, fused in front of the refine stage, by stream_tail (correlator.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200sync {

// all 32 lanes of a warp call this with the same arguments; emit(p) is called (by all lanes) for every detection,
// in increasing order.  Returns the search offset after the range (>= range, or where the search gave up).
template <class Emit>
__device__ __forceinline__ long long peak_walk_warp(const uint32_t* cand_s, const uint32_t* pass_s, int nwords,
                                                     long long range, int T, long long j, Emit&& emit) {
    const int lane = threadIdx.x & 31;
    while (j < range) {
        long long found = -1;
        const int w0 = (int)(j >> 5);
        for (int base = w0; base < nwords; base += 32) {
            const int wi = base + lane;
            uint32_t word = wi < nwords ? cand_s[wi] : 0u;
            if (wi == w0) word &= 0xffffffffu << (j & 31);
            const uint32_t m = __ballot_sync(0xffffffffu, word != 0u);
            if (m != 0u) {
                const int l = __ffs(m) - 1;
                const uint32_t ww = __shfl_sync(0xffffffffu, word, l);
                found = (long long)(base + l) * 32 + (__ffs(ww) - 1);
                break;
            }
        }
        if (found < 0 || found >= range) break;      // no candidate left: the search resumes at the range end
        if ((pass_s[found >> 5] >> (found & 31)) & 1u) emit(found);
        j = found + T + 1;                           // :296-297: the search restarts after the timeout
    }
    return j;
}

}  // namespace b200sync

// tmem.cuh — Tensor Memory (TMEM, 256 KiB per SM: 128 lanes x 512 columns x 32 bit) used as a per-thread
// scratch / constant store next to a shared-memory-bound kernel.
//
// No tensor-core instruction is involved: tcgen05.st / tcgen05.ld (SASS STTM / LDTM) move 32-bit words
// between a thread's registers and "its" TMEM lane.  With the 32x32b shape, thread t of warp w addresses lane
// 32*(w % 4) + t, and .xN moves N consecutive columns of that lane.  Measured on B200
// (profiles/r2_ubench_pipes.txt): tcgen05.ld delivers ~470 B/clk/SM and runs BESIDE the LSU pipe (LDS.64 is
// 127 B/clk/SM), so whatever is parked in TMEM no longer costs shared-memory wavefronts.
//
// Rules (PTX ISA, tcgen05): allocation is per CTA, by ONE warp, a power of two >= 32 columns; the address
// lands in shared memory; the same warp deallocates before the CTA exits; registers written by tcgen05.ld may
// only be read after tcgen05.wait::ld; data written by tcgen05.st may only be loaded after tcgen05.wait::st.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200sync {

// one warp: allocate `cols` columns (power of two, 32..512), write the base address to *slot_smem
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     static_cast<uint32_t>(__cvta_generic_to_shared(slot_smem))),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// address of column `col` in the calling thread's lane quarter (the hardware adds the lane within the quarter)
__device__ __forceinline__ uint32_t tmem_addr(uint32_t tbase, int warp_in_cta, int col) {
    return tbase + (static_cast<uint32_t>(32 * (warp_in_cta & 3)) << 16) + static_cast<uint32_t>(col);
}

// 8 complex values = 16 columns of the thread's lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float2 (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=f"(r[0].x), "=f"(r[0].y), "=f"(r[1].x), "=f"(r[1].y), "=f"(r[2].x), "=f"(r[2].y), "=f"(r[3].x), "=f"(r[3].y),
          "=f"(r[4].x), "=f"(r[4].y), "=f"(r[5].x), "=f"(r[5].y), "=f"(r[6].x), "=f"(r[6].y), "=f"(r[7].x), "=f"(r[7].y)
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float2 (&r)[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};" ::
            "f"(r[0].x), "f"(r[0].y), "f"(r[1].x), "f"(r[1].y), "f"(r[2].x), "f"(r[2].y), "f"(r[3].x), "f"(r[3].y),
        "f"(r[4].x), "f"(r[4].y), "f"(r[5].x), "f"(r[5].y), "f"(r[6].x), "f"(r[6].y), "f"(r[7].x), "f"(r[7].y), "r"(taddr)
        : "memory");
}
// compiler-level dependency: nothing that reads r may be scheduled before this point (placed right after
// tmem_wait_ld(), so no use of a loaded register can move above the wait).  Emits no instruction.
__device__ __forceinline__ void tmem_use(float2 (&r)[8]) {
    asm volatile("" : "+f"(r[0].x), "+f"(r[0].y), "+f"(r[1].x), "+f"(r[1].y), "+f"(r[2].x), "+f"(r[2].y), "+f"(r[3].x),
                      "+f"(r[3].y), "+f"(r[4].x), "+f"(r[4].y), "+f"(r[5].x), "+f"(r[5].y), "+f"(r[6].x), "+f"(r[6].y),
                      "+f"(r[7].x), "+f"(r[7].y)::"memory");
}

}  // namespace b200sync

// tma.cuh — 1-D bulk copies global -> shared memory by the Tensor Memory Accelerator (cp.async.bulk,
// SASS UBLKCP) completing on an mbarrier.  Used to stage per-CTA constant tables (filter taps, FFT
// twiddles): one elected thread issues the copy, the copy engine moves the bytes while all threads go on
// staging samples, everyone waits on the barrier's phase before the first read.
// Sizes must be multiples of 16 bytes, both addresses 16-byte aligned.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200sync {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // visible to the async proxy
}

// arm the barrier with the byte count and start the copy (call from ONE thread, after mbar_init + a CTA barrier)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

}  // namespace b200sync

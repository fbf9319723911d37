// Development aid: is TMEM usable as per-thread constant storage next to a shared-memory-bound kernel?
// Measures, for one CTA of 768 threads per SM (the correlator's shape):
//   (a) tcgen05.ld.32x32b.x16 throughput alone,  (b) LDS.64 throughput alone (same bytes per thread),
//   (c) both interleaved — if the paths are independent, (c) ~ max(a, b), not a + b.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_tmem scripts/ubench_tmem.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld16(unsigned taddr, unsigned (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(unsigned taddr, const unsigned (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};\n" ::
            "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(taddr));
}

template <int MODE>
__global__ void __launch_bounds__(768, 1) k(unsigned* out, long long* cycles, int iters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned tbase_s;
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 16 * 768; i += 768) sm[i] = make_float2((float)i, 1.0f);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"l"(
            (unsigned long long)__cvta_generic_to_shared(&tbase_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const unsigned tbase = tbase_s;
    // warp w owns TMEM lanes 32*(w%4)..+31 and, among the 6 warps sharing that quarter, columns 64*(w/4)..+63
    const unsigned taddr = tbase + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)(64 * (warp >> 2));
    unsigned v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = tid * 16 + j;
    for (int c = 0; c < 64; c += 16) tmem_st16(taddr + c, v);
    asm volatile("tcgen05.wait::st.sync.aligned;\n");
    __syncthreads();
    unsigned acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int c = 0; c < 64; c += 16) {
                unsigned r[16];
                tmem_ld16(taddr + c, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;\n");
#pragma unroll
                for (int j = 0; j < 16; ++j) acc += r[j];
            }
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {  // 32 x LDS.64 = 256 B per thread, like the 64 words above
                unsigned fx, fy;
                const unsigned sa = (unsigned)__cvta_generic_to_shared(&sm[((j + it) & 15) * 768 + tid]);
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];\n" : "=r"(fx), "=r"(fy) : "r"(sa));
                acc += fx ^ fy;
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    // correctness of the TMEM round trip
    unsigned r[16];
    tmem_ld16(taddr + 16, r);
    asm volatile("tcgen05.wait::ld.sync.aligned;\n");
    unsigned bad = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) bad |= (r[j] != (unsigned)(tid * 16 + j));
    out[blockIdx.x * 768 + tid] = acc + (bad << 31);
    if (bad) atomicAdd((unsigned long long*)&cycles[1], 1ull);
    if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tbase));
    (void)lane;
}

int main() {
    unsigned* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 768 * 4);
    cudaMallocManaged(&cyc, 16);
    const int iters = 2000;
    const size_t smem = 16 * 768 * 8;
    auto run = [&](auto kern, const char* name) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cyc[0] = cyc[1] = 0;
        kern<<<148, 768, smem>>>(out, cyc, iters);
        cudaError_t e = cudaDeviceSynchronize();
        // per iteration each thread moved 256 B through each active path: 768 * 256 B = 192 KiB per SM
        std::printf("%s: %s  %.1f clk/iter  -> %.1f B/clk/SM per path  (round-trip errors: %lld)\n", name,
                    cudaGetErrorString(e), (double)cyc[0] / iters, 768.0 * 256.0 * iters / (double)cyc[0], cyc[1]);
    };
    run(k<0>, "tcgen05.ld x16 only ");
    run(k<1>, "LDS.64 only         ");
    run(k<2>, "both interleaved    ");
    return 0;
}

// Development aid: which on-chip data paths run beside the shared-memory pipe of an LSU-bound kernel?
// One CTA of 768 threads per SM (the correlator's shape).  Each path moves 256 B per thread per iteration:
//   T  tcgen05.ld 32x32b (TMEM -> registers), shapes .x16 / .x32 / .x64, one wait::ld per instruction
//   L  32 x LDS.64, conflict-free (256 B per thread)
//   S  64 x SHFL.BFLY (32-bit)
//   F  512 FFMAs per thread (8 independent chains)
// all interleaved instruction by instruction inside every warp (a first version ran the paths as separate
// phases of the loop body: the 24 warps then move through the phases in lock step and every combination
// reads as the SUM of its parts, which says nothing about the hardware)
// and combinations.  If two paths are independent, the combined time is ~ max, not the sum.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_pipes scripts/ubench_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

#define LD_ARGS16(o) "=r"(r[o+0]), "=r"(r[o+1]), "=r"(r[o+2]), "=r"(r[o+3]), "=r"(r[o+4]), "=r"(r[o+5]), "=r"(r[o+6]), "=r"(r[o+7]), \
                     "=r"(r[o+8]), "=r"(r[o+9]), "=r"(r[o+10]), "=r"(r[o+11]), "=r"(r[o+12]), "=r"(r[o+13]), "=r"(r[o+14]), "=r"(r[o+15])

__device__ __forceinline__ void tmem_ld16(unsigned taddr, unsigned* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : LD_ARGS16(0) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, unsigned* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : LD_ARGS16(0), LD_ARGS16(16) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld64(unsigned taddr, unsigned* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
        "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,"
        "%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];\n"
        : LD_ARGS16(0), LD_ARGS16(16), LD_ARGS16(32), LD_ARGS16(48) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(unsigned taddr, const unsigned* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};\n" ::
            "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// TSHAPE: 0 none, 16 / 32 / 64;  SAMEADDR: all six warps of a lane quarter read the same 64 columns (the
// correlator's case: every FFT group reads the same constants) instead of private column ranges
template <int TSHAPE, bool L, bool S, bool F, bool SAMEADDR>
__global__ void __launch_bounds__(768, 1) k(unsigned* out, long long* cycles, int iters, int zero) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned tbase_s;
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5;
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
    const unsigned taddr_own = tbase + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)(64 * (warp >> 2));
    const unsigned taddr = SAMEADDR ? tbase + ((unsigned)(32 * (warp & 3)) << 16) : taddr_own;
    unsigned v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = tid * 16 + j;
    for (int c = 0; c < 64; c += 16) tmem_st16(taddr_own + c, v);
    asm volatile("tcgen05.wait::st.sync.aligned;\n");
    __syncthreads();
    unsigned acc = tid * 2654435761u;
    float f0 = (float)tid, f1 = 1.0f, f2 = 2.0f, f3 = 3.0f, f4 = 0.5f, f5 = 0.25f, f6 = 4.f, f7 = 5.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        // finely interleaved, like a real kernel: per quarter, one TMEM load (64 B/thread) is issued, then
        // 8 x (LDS.64 | 2 SHFL | 16 FFMA) run while it is in flight, then it is waited for and consumed
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            unsigned r[32];
            if (TSHAPE == 16) tmem_ld16(taddr + 16 * qd, r);
            if (TSHAPE == 32 && (qd & 1) == 0) tmem_ld32(taddr + 16 * qd, r);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (L) {
                    unsigned fx, fy;
                    const unsigned sa = (unsigned)__cvta_generic_to_shared(&sm[((8 * qd + j + it) & 15) * 768 + tid + (qd >> 1) * zero]);
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];\n" : "=r"(fx), "=r"(fy) : "r"(sa) : "memory");
                    acc += fx ^ fy;
                }
                if (S) {
                    acc += __shfl_xor_sync(0xffffffffu, acc + j, 1 + (j & 15));
                    acc += __shfl_xor_sync(0xffffffffu, acc + j, 2 + (j & 15));
                }
                if (F) {
                    f0 = fmaf(f0, f4, f5); f1 = fmaf(f1, f4, f5); f2 = fmaf(f2, f4, f5); f3 = fmaf(f3, f4, f5);
                    f6 = fmaf(f6, f4, f5); f7 = fmaf(f7, f4, f5); f0 = fmaf(f0, f5, f4); f1 = fmaf(f1, f5, f4);
                    f2 = fmaf(f2, f5, f4); f3 = fmaf(f3, f5, f4); f6 = fmaf(f6, f5, f4); f7 = fmaf(f7, f5, f4);
                    f0 = fmaf(f0, f4, f5); f1 = fmaf(f1, f4, f5); f2 = fmaf(f2, f4, f5); f3 = fmaf(f3, f4, f5);
                }
            }
            if (TSHAPE == 16) {
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) acc += r[j];
            }
            if (TSHAPE == 32 && (qd & 1) == 1) {
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc += r[j];
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    unsigned r[16];
    tmem_ld16(taddr_own + 16, r);
    tmem_wait_ld();
    unsigned bad = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) bad |= (r[j] != (unsigned)(tid * 16 + j));
    out[blockIdx.x * 768 + tid] = acc + (bad << 31) + (unsigned)(f0 + f1 + f2 + f3 + f6 + f7);
    if (bad) atomicAdd((unsigned long long*)&cycles[1], 1ull);
    if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tbase));
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
        kern<<<148, 768, smem>>>(out, cyc, iters, 0);
        cudaError_t e = cudaDeviceSynchronize();
        std::printf("%-44s %s  %8.1f clk/iter  (192 KiB per path per iter -> %6.1f B/clk/SM if one path)  errors %lld\n", name,
                    cudaGetErrorString(e), (double)cyc[0] / iters, 768.0 * 256.0 * iters / (double)cyc[0], cyc[1]);
    };
    run(k<16, false, false, false, false>, "T x16 private columns");
    run(k<32, false, false, false, false>, "T x32 private columns");
    run(k<16, false, false, false, true>, "T x16 same columns (6 warps)");
    run(k<32, false, false, false, true>, "T x32 same columns (6 warps)");
    run(k<0, true, false, false, false>, "L (32 LDS.64 = 256 B/thread)");
    run(k<0, false, true, false, false>, "S (64 SHFL)");
    run(k<0, false, false, true, false>, "F (512 FFMA)");
    run(k<32, true, false, false, true>, "T x32 same + L");
    run(k<16, true, false, false, true>, "T x16 same + L");
    run(k<0, true, true, false, false>, "L + S");
    run(k<32, false, false, true, true>, "T x32 same + F");
    run(k<0, true, false, true, false>, "L + F");
    run(k<32, true, false, true, true>, "T x32 same + L + F");
    return 0;
}

// Microbenchmark: issue rate of scalar FP32 (FFMA/FADD) vs packed FP32x2 (FFMA2/FADD2) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_f32x2 scripts/ubench_f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096, ACC = 8;
template <int MODE>
__global__ void k(float2* out, float2 w, float2 z) {
    float2 a[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) {
            if (MODE == 0) { a[i].x = __fmaf_rn(a[i].x, w.x, z.x); a[i].y = __fmaf_rn(a[i].y, w.y, z.y); }
            if (MODE == 1) { a[i] = __ffma2_rn(a[i], w, z); }
            if (MODE == 2) { a[i].x = __fadd_rn(a[i].x, z.x); a[i].y = __fadd_rn(a[i].y, z.y); }
            if (MODE == 3) { a[i] = __fadd2_rn(a[i], z); }
            if (MODE == 4) { a[i].x = __fmul_rn(a[i].x, w.x); a[i].y = __fmul_rn(a[i].y, w.y); }
            if (MODE == 5) { a[i] = __fmul2_rn(a[i], w); }
        }
    }
    float2 s = a[0];
#pragma unroll
    for (int i = 1; i < ACC; ++i) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, float2* d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, block = 256;
    k<MODE><<<grid, block>>>(d, make_float2(1.0001f, 0.9999f), make_float2(1e-3f, -1e-3f));
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<grid, block>>>(d, make_float2(1.0001f, 0.9999f), make_float2(1e-3f, -1e-3f));
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double ops = (double)grid * block * ITERS * ACC * 2;  // scalar FP32 ops (an FMA counts once)
    printf("%-8s %.3f ms  %.2f T scalar-op/s  (%.1f op/clk/SM at 1.965 GHz)\n", name, ms, ops / ms / 1e9,
           ops / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    float2* d; cudaMalloc(&d, sizeof(float2) * 148 * 8 * 256);
    run<0>("FFMA", d); run<1>("FFMA2", d); run<2>("FADD", d); run<3>("FADD2", d); run<4>("FMUL", d); run<5>("FMUL2", d);
    return 0;
}

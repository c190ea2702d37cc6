// Microbenchmark: issue / pipe throughput of FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/ffma2_bench tools/micro/ffma2_bench.cu ; run it on a B200 (gpurun)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
    float2 a[8], b = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }   // 2 FFMA
            else a[i] = __ffma2_rn(a[i], b, c);                                                       // 1 FFMA2
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 1024 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode)
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters); else k<1><<<148 * 8, 256>>>(d, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flops = 2.0 * 16 * (double)iters * 148 * 8 * 256;
            printf("mode %d (%s): %.3f ms  %.1f TFLOP/s\n", mode, mode ? "FFMA2" : "FFMA", ms, flops / ms / 1e9);
        }
    return 0;
}

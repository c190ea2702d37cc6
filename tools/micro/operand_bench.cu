// Microbenchmark: does the rate of the packed FFMA2 depend on how many distinct register pairs it reads?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/operand_bench tools/micro/operand_bench.cu
// run on a B200:  tools/micro/operand_bench [warps per SM sub-partition]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define NACC 16
// MODE 0: acc = acc * b + c            (one varying register pair; b, c loop constants)
// MODE 1: acc = x[i] * y[j] + acc      (three distinct register pairs)
// MODE 2: acc = x[i] * x[i] + acc      (two distinct register pairs)
// MODE 3: acc = x[i] * s + acc, s a 32-bit scalar broadcast (the i-atom operand form of the pair kernel)
// MODE 4: like 1 with scalar FFMA (two per packed operation)
template <int MODE>
__global__ void __launch_bounds__(128) k(float* out, const float* in, int iters) {
    float2 acc[NACC], x[8], y[8];
    float sc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[i] = make_float2(in[threadIdx.x + 2 * i], in[threadIdx.x + 2 * i + 1]);
        y[i] = make_float2(in[threadIdx.x + 16 + 2 * i], in[threadIdx.x + 17 + 2 * i]);
        sc[i] = in[threadIdx.x + 40 + i];
    }
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = make_float2(i, -i);
    const float2 b = make_float2(in[0], in[1]), c = make_float2(in[2], in[3]);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < NACC; ++i) {
                const int xi = (i + u) & 7, yi = (i * 3 + u + 1) & 7;
                if (MODE == 0) acc[i] = __ffma2_rn(acc[i], b, c);
                if (MODE == 1) acc[i] = __ffma2_rn(x[xi], y[yi], acc[i]);
                if (MODE == 2) acc[i] = __ffma2_rn(x[xi], x[xi], acc[i]);
                if (MODE == 3) acc[i] = __ffma2_rn(x[xi], make_float2(sc[yi], sc[yi]), acc[i]);
                if (MODE == 4) { acc[i].x = fmaf(x[xi].x, y[yi].x, acc[i].x); acc[i].y = fmaf(x[xi].y, y[yi].y, acc[i].y); }
            }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float *d, *din;
static int g_ctas = 3;
template <int MODE> void run(const char* name) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 5000;
    const size_t smem = (200 * 1024) / g_ctas - 2048;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<148 * g_ctas, 128, smem>>>(d, din, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double n_packed = 4.0 * NACC * iters;             // per warp
    const double cyc = best * 1e-3 * 1.965e9 / (n_packed * g_ctas);
    printf("%-44s %.3f ms  %.2f cycles per packed operation per sub-partition  (%.1f TFLOP/s)\n", name, best, cyc,
           4.0 * n_packed * 148 * g_ctas * 128 / best / 1e9);
}
int main(int argc, char** argv) {
    if (argc > 1) g_ctas = atoi(argv[1]);
    cudaMalloc(&d, 148 * 16 * 128 * sizeof(float));
    cudaMalloc(&din, 1024 * sizeof(float)); cudaMemset(din, 0, 1024 * sizeof(float));
    printf("warps per sub-partition: %d\n", g_ctas);
    run<0>("FFMA2 acc = acc*b + c (1 varying pair)");
    run<1>("FFMA2 acc = x*y + acc (3 distinct pairs)");
    run<2>("FFMA2 acc = x*x + acc (2 distinct pairs)");
    run<3>("FFMA2 acc = x*s + acc (pair, scalar, pair)");
    run<4>("2 x FFMA acc = x*y + acc (scalar)");
    return 0;
}

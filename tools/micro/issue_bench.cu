// Microbenchmark: does a packed FFMA2 leave its second pipe cycle free for the issue of another instruction?
// Each mode runs the same 16 independent FFMA2 chains per thread and adds K "other" instructions (FMNMX, MUFU,
// SHFL, LDS, IADD3) between them.  If the others issue in the shadow of the FFMA2s the time does not change until
// K exceeds the FFMA2 count; if an FFMA2 holds the issue port for both cycles the time grows by K / 32.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/issue_bench tools/micro/issue_bench.cu
// run on a B200:  tools/micro/issue_bench [warps_per_sm_subpartition]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define NCH 16
enum { NONE = 0, FMNMX = 1, MUFU = 2, SHFL = 3, LDS = 4, IADD = 5, FFMA1 = 6, FMNMX3 = 7 };

template <int OTHER, int K, int SCALAR>
__global__ void __launch_bounds__(128) k(float* out, int iters) {
    __shared__ float sm[128];
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    float2 a[NCH];
    const float2 b = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
    float o[8];
    int io[8];
#pragma unroll
    for (int i = 0; i < NCH; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i] = 1.5f + i + threadIdx.x; io[i] = threadIdx.x + i; }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            if (SCALAR) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
            else a[i] = __ffma2_rn(a[i], b, c);
            if (i < K) {
                const int j = i & 7;
                if (OTHER == FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(o[j]) : "f"(a[(i + 8) % NCH].y));
                if (OTHER == FMNMX3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(o[j]) : "f"(a[(i + 8) % NCH].y), "f"(a[(i + 9) % NCH].x));
                if (OTHER == MUFU) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(o[j]));
                if (OTHER == SHFL) o[j] = __shfl_sync(0xffffffffu, o[j], (threadIdx.x + 1) & 31);
                if (OTHER == LDS) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o[j]) : "r"((unsigned)(__cvta_generic_to_shared(sm) + 4 * ((threadIdx.x + j) & 127))));
                if (OTHER == IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(io[j]) : "r"(it));
                if (OTHER == FFMA1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(o[j]) : "f"(b.x), "f"(c.x));
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += a[i].x + a[i].y;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += o[i] + io[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float* d;
static int g_ctas_per_sm = 3;
template <int OTHER, int K, int SCALAR>
void run(const char* name) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    float best = 1e9f;
    // one 128-thread CTA = one warp per SM sub-partition; g_ctas_per_sm CTAs per SM (limited through dynamic smem)
    const size_t smem = (200 * 1024) / g_ctas_per_sm - 2048;
    cudaFuncSetAttribute(k<OTHER, K, SCALAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<OTHER, K, SCALAR><<<148 * g_ctas_per_sm, 128, smem>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double flops = 2.0 * 2 * NCH * (double)iters * 148 * g_ctas_per_sm * 128;
    // cycles per iteration per sub-partition at 1.965 GHz; the FFMA2s alone need 2 * NCH * warps
    const double cyc = best * 1e-3 * 1.965e9 / iters;
    printf("%-28s K=%2d  %.3f ms  %.1f TFLOP/s  %.1f cycles/iter (fma floor %d, +others %d)\n", name, K, best, flops / best / 1e9, cyc,
           2 * NCH * g_ctas_per_sm, (2 * NCH + K) * g_ctas_per_sm);
}

int main(int argc, char** argv) {
    if (argc > 1) g_ctas_per_sm = atoi(argv[1]);
    cudaMalloc(&d, 148 * 16 * 128 * sizeof(float));
    printf("warps per sub-partition: %d\n", g_ctas_per_sm);
    run<NONE, 0, 0>("FFMA2 only");
    run<NONE, 0, 1>("FFMA only (scalar)");
    run<FMNMX, 8, 0>("FFMA2 + FMNMX");
    run<FMNMX, 16, 0>("FFMA2 + FMNMX");
    run<FMNMX3, 8, 0>("FFMA2 + FMNMX3");
    run<MUFU, 4, 0>("FFMA2 + MUFU");
    run<MUFU, 8, 0>("FFMA2 + MUFU");
    run<SHFL, 8, 0>("FFMA2 + SHFL");
    run<LDS, 8, 0>("FFMA2 + LDS");
    run<IADD, 8, 0>("FFMA2 + IADD");
    run<IADD, 16, 0>("FFMA2 + IADD");
    run<FFMA1, 8, 0>("FFMA2 + scalar FFMA");
    run<FMNMX, 8, 1>("FFMA(scalar) + FMNMX");
    run<MUFU, 8, 1>("FFMA(scalar) + MUFU");
    return 0;
}

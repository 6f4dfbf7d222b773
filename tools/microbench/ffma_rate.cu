// Microbenchmark (GPU box): issue rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma_rate tools/microbench/ffma_rate.cu && /tmp/ffma_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
    // 32 independent accumulators (64 for FFMA2), outer-product style: acc[i][j] += a[i] * b[j]
    float a[4], b[8];
    for (int i = 0; i < 4; ++i) a[i] = seed + i + threadIdx.x * 1e-3f;
    for (int j = 0; j < 8; ++j) b[j] = seed * 0.5f + j;
    if (MODE == 0) {
        float acc[4][8];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            for (int i = 0; i < 4; ++i) a[i] += 1e-7f;
        }
        float s = 0.f;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 8; ++j) s += acc[i][j];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        float2 acc[4][8];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 8; ++j) acc[i][j] = make_float2(0.f, 0.f);
        float2 b2[8];
        for (int j = 0; j < 8; ++j) b2[j] = make_float2(b[j], b[j] + 0.25f);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 aa = make_float2(a[i], a[i]);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = ffma2(aa, b2[j], acc[i][j]);
            }
            for (int i = 0; i < 4; ++i) a[i] += 1e-7f;
        }
        float s = 0.f;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 8; ++j) s += acc[i][j].x + acc[i][j].y;
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters, 1.f);
            else k<1><<<148 * 8, 256>>>(out, iters, 1.f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double fma = (double)148 * 8 * 256 * iters * 32 * (mode == 0 ? 1 : 2);
            printf("%s: %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", mode == 0 ? "FFMA " : "FFMA2", ms, 2 * fma / ms / 1e9,
                   fma / (ms * 1e-3) / 148 / 1.965e9);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

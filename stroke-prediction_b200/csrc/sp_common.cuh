// sp_common.cuh — shared device helpers and host-side error plumbing for libstroke_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/stroke_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libstroke_b200 is written for sm_100a (B200) only"
#endif

// ---- host-side error handling -------------------------------------------------------------------------------
void sp_set_error(const char* fmt, ...);

#define SP_REQUIRE(cond, ...)                       \
    do {                                            \
        if (!(cond)) {                              \
            sp_set_error(__VA_ARGS__);              \
            return -1;                              \
        }                                           \
    } while (0)

#define SP_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            sp_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return (int)e__;                                                                  \
        }                                                                                     \
    } while (0)

#define SP_LAUNCH_OK(name)                                                             \
    do {                                                                               \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) {                                                      \
            sp_set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));    \
            return (int)e__;                                                           \
        }                                                                              \
    } while (0)

static inline cudaStream_t sp_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int sp_num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

static inline int64_t sp_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- activations ----------------------------------------------------------------------------
__device__ __forceinline__ float sp_act_fwd(float v, int act, float alpha) {
    switch (act) {
        case SP_ACT_ELU:     return v > 0.f ? v : alpha * expm1f(v);
        case SP_ACT_LEAKY:   return v > 0.f ? v : alpha * v;
        case SP_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        default:             return v;
    }
}
// derivative expressed through the activation OUTPUT y (in-place activations in the reference keep only y)
__device__ __forceinline__ float sp_act_bwd(float y, int act, float alpha) {
    switch (act) {
        case SP_ACT_ELU:     return y > 0.f ? 1.f : (y + alpha);
        case SP_ACT_LEAKY:   return y > 0.f ? 1.f : alpha;
        case SP_ACT_SIGMOID: return y * (1.f - y);
        default:             return 1.f;
    }
}

// ---- packed fp32 FMA (Blackwell FFMA2): {acc.x, acc.y} += x * {w.x, w.y} -------------------------------------------
// One instruction, two FMAs per lane; ptxas folds the (x, x) pair into a scalar-broadcast operand (`Rx.F32`), so the
// FMA pipe still retires 128 FMA/clk/SM while only every second issue slot is an FFMA and every operand is one
// aligned 64-bit register read (no even/odd bank conflicts).
__device__ __forceinline__ float2 sp_ffma2(float x, float2 w, float2 acc) {
    float2 xx = make_float2(x, x);
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&xx), rb = *reinterpret_cast<unsigned long long*>(&w),
                       rc = *reinterpret_cast<unsigned long long*>(&acc), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

// ---- reductions ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double sp_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float sp_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 128-bit accessors (activations are touched once per kernel; keep them out of L1)
__device__ __forceinline__ float4 sp_ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

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
// expm1 on v <= 0 for the ELU epilogues (16 .. 24 evaluations per output voxel in the tensor-core tiers' epilogue warps; 64 per
// thread in the thin first-layer kernel, where the previous all-FMA form — Cody-Waite exponential + Taylor series, ~28
// instructions per value — was more than half of the kernel's instruction stream), branch free, ~14 instructions:
//   -1 < v <= 0: v * q(v), q = degree-6 Chebyshev fit of expm1(v) / v on [-1, 0] (fit error 2.4e-8; fp32 Horner: 1.5e-7 max relative)
//   v <= -1:     ex2.approx(v * log2 e) - 1 on the MUFU pipe: |result| >= 0.63 and exp(v) <= 0.37, so the 2^-22 relative error
//                bound of ex2.approx and the rounding of v * log2 e stay below 2.4e-7 relative to the result.
// (libm's expm1f in fp32: 1.9e-7.)  tools/microbench/elu_err.cu measures it on the device.
__device__ __forceinline__ float sp_ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#define SP_EM1_C1 0.4999997913837433f
#define SP_EM1_C2 0.16666345298290253f
#define SP_EM1_C3 0.0416470430791378f
#define SP_EM1_C4 0.008275879546999931f
#define SP_EM1_C5 0.0013011073460802436f
#define SP_EM1_C6 0.0001291868684347719f
__device__ __forceinline__ float sp_expm1_neg(float v) {
    const float q = fmaf(v, fmaf(v, fmaf(v, fmaf(v, fmaf(v, fmaf(v, SP_EM1_C6, SP_EM1_C5), SP_EM1_C4), SP_EM1_C3), SP_EM1_C2), SP_EM1_C1), 1.f);
    const float e = sp_ex2_approx(v * 1.442695041f) - 1.f;
    return v > -1.f ? v * q : e;
}
// ELU of N values at once, written level by level so that the N independent dependency chains are interleaved in the instruction
// stream.  Evaluated one value after the other (what the compiler emits for a loop over sp_act_fwd) the epilogue warps of the
// tensor-core tiers issued one instruction every ~4 cycles: ncu stall_wait 3 : 1 selected.
template <int N>
__device__ __forceinline__ void sp_elu_n(float* v, float alpha) {
    float x[N], e[N], q[N];
#pragma unroll
    for (int j = 0; j < N; ++j) x[j] = fminf(v[j], 0.f);
#pragma unroll
    for (int j = 0; j < N; ++j) { e[j] = sp_ex2_approx(x[j] * 1.442695041f); q[j] = fmaf(x[j], SP_EM1_C6, SP_EM1_C5); }
#pragma unroll
    for (int j = 0; j < N; ++j) q[j] = fmaf(x[j], q[j], SP_EM1_C4);
#pragma unroll
    for (int j = 0; j < N; ++j) q[j] = fmaf(x[j], q[j], SP_EM1_C3);
#pragma unroll
    for (int j = 0; j < N; ++j) q[j] = fmaf(x[j], q[j], SP_EM1_C2);
#pragma unroll
    for (int j = 0; j < N; ++j) q[j] = fmaf(x[j], q[j], SP_EM1_C1);
#pragma unroll
    for (int j = 0; j < N; ++j) q[j] = fmaf(x[j], q[j], 1.f);
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const float m = x[j] > -1.f ? x[j] * q[j] : e[j] - 1.f;                                          // expm1(x), x <= 0
        v[j] = fmaf(alpha, m, fmaxf(v[j], 0.f));
    }
}

__device__ __forceinline__ float sp_act_fwd(float v, int act, float alpha) {
    switch (act) {
        // (two-sided form on purpose: it keeps this switch a real branch on the uniform `act`.  Written as straight-line code the
        // whole switch is if-converted and every LeakyReLU / linear layer pays for the exponential: U-Net step +4 %.  Epilogues that
        // evaluate many ELUs per thread dispatch on `act` once and call sp_elu_n.)
        case SP_ACT_ELU:     return v > 0.f ? v : alpha * sp_expm1_neg(v);
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

// asr_b200 -- shared host/device helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/asr_b200.h"

#define ASRB_CUDA_OK(expr)                                          \
    do {                                                            \
        cudaError_t _e = (expr);                                    \
        if (_e != cudaSuccess) return (int)_e;                      \
    } while (0)

#define ASRB_LAUNCH_OK()                                            \
    do {                                                            \
        cudaError_t _e = cudaGetLastError();                        \
        if (_e != cudaSuccess) return (int)_e;                      \
    } while (0)

#define ASRB_REQUIRE(cond, code)                                    \
    do {                                                            \
        if (!(cond)) return (code);                                 \
    } while (0)

namespace asrb {

constexpr int kNumSMs = 148;  // B200

extern unsigned g_debug_flags;  // see asrb_set_debug_flags

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div64(long long a, long long b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
// tanh via exp keeps ~1e-7 abs accuracy (tanh.approx is only ~1e-3)
__device__ __forceinline__ float tanhf_(float x) {
    float ax = fabsf(x);
    float e = __expf(-2.0f * ax);
    float t = (1.0f - e) / (1.0f + e);
    return copysignf(t, x);
}

// log(exp(a)+exp(b)) safe for -inf operands
__device__ __forceinline__ float log_add(float a, float b) {
    float m = fmaxf(a, b);
    if (m == -INFINITY) return -INFINITY;
    return m + log1pf(expf(-fabsf(a - b)));
}

}  // namespace asrb

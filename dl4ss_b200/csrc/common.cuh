// Shared host/device helpers for the dl4ss_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <mutex>
#include "../../include/dl4ss_b200.h"

namespace dl4ss {

constexpr int DL4SS_MAX_DEVICES = 64;

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define DL4SS_CHECK_ARG(cond, ...)                                   \
    do {                                                             \
        if (!(cond)) {                                               \
            dl4ss::set_error(__VA_ARGS__);                           \
            return DL4SS_EINVAL;                                     \
        }                                                            \
    } while (0)

#define DL4SS_CUDA(call)                                                                  \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            dl4ss::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),     \
                             __FILE__, __LINE__);                                         \
            return DL4SS_ECUDA;                                                           \
        }                                                                                 \
    } while (0)

#define DL4SS_LAUNCH_CHECK(name)                                                          \
    do {                                                                                  \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess) {                                                         \
            dl4ss::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));   \
            return DL4SS_ECUDA;                                                           \
        }                                                                                 \
        dl4ss::count_launch();                                                            \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline long long cdivll(long long a, long long b) { return (a + b - 1) / b; }

int sm_count();   // SMs of the current device (cached)

#ifdef __CUDACC__
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// accurate-enough fp32 transcendental helpers (abs err ~1e-7, far below the 1e-4 parity bar)
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }   // MUFU.EX2 + MUFU.RCP, ~2 ulp
__device__ __forceinline__ float tanh_f(float x) {
    // tanh(x) = 1 - 2/(exp(2x)+1); saturates cleanly for |x| large
    float e = __expf(2.0f * x);
    return 1.0f - __fdividef(2.0f, e + 1.0f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

}  // namespace dl4ss

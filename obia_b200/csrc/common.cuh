// Shared helpers for the obia_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/obia_b200.h"

namespace obia {

// thread-local error text returned by obia_b200_last_error()
char *err_buf();
int set_err(int code, const char *fmt, ...);

#define OBIA_CUDA_CHECK(expr)                                                        \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess)                                                       \
            return obia::set_err(OBIA_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);        \
    } while (0)

// every kernel launch of the library goes through OBIA_LAUNCH_CHECK, which also
// counts it (obia_b200_launch_count: the "gpu_launches" figure of bench.py)
void count_launch();

#define OBIA_LAUNCH_CHECK()                                                          \
    do {                                                                             \
        obia::count_launch();                                                        \
        cudaError_t _e = cudaGetLastError();                                         \
        if (_e != cudaSuccess)                                                       \
            return obia::set_err(OBIA_B200_ERR_CUDA, "kernel launch failed: %s (%s:%d)", \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);        \
    } while (0)

// optional CUDA-event timing of one kernel class (the SLIC assign+update kernel)
void prof_begin(cudaStream_t st);
void prof_end(cudaStream_t st);

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// order-preserving float <-> uint32 key (for atomicMin/Max on floats)
__host__ __device__ __forceinline__ uint32_t float_to_key(float f)
{
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    uint32_t b;
    memcpy(&b, &f, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float key_to_float(uint32_t k)
{
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

// streaming 128-bit load that does not allocate in L1 (data read once)
__device__ __forceinline__ float4 ldg_stream_f4(const float4 *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

}  // namespace obia

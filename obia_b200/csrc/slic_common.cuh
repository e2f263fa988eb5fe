// Shared pieces of the SLIC kernels (slic.cu: exact mode, slic_fast.cu: tolerance mode).
#pragma once
#include "common.cuh"

#ifndef OBIA_NS
#define OBIA_NS 2   // strip phases per CTA tile (32 x (32*NS) pixels for 4 px / lane)
#endif

namespace obia {

constexpr int kWarps = 8;

// workspace layout (all 16-byte aligned)
struct SlicWs {
    unsigned long long *acc;  // [n][3+Cf] count, sum y, sum x, fixed-point colour sums
    int32_t *head;            // [ncy*ncx] cell -> first centre
    int32_t *next;            // [n]
    float *maxdc;             // [n] SLICO: largest colour distance seen per centre (slic_zero)
    int64_t ncy, ncx;
    int64_t bytes;
    float sp_y, sp_x;         // voxel spacing of skimage's `spacing=(sy, sx)` (1, 1 unless the caller says otherwise)
};

static SlicWs slic_ws_layout(void *base, int64_t H, int64_t W, int Cf, int64_t n, int step_y, int step_x)
{
    SlicWs w;
    w.sp_y = w.sp_x = 1.0f;
    w.ncy = ceil_div(H, step_y);
    w.ncx = ceil_div(W, step_x);
    char *p = (char *)base;
    int64_t off = 0;
    w.acc = (unsigned long long *)(p + off);
    off += round_up(n * (3 + Cf) * 8, 256);
    w.head = (int32_t *)(p + off);
    off += round_up(w.ncy * w.ncx * 4, 256);
    w.next = (int32_t *)(p + off);
    off += round_up(n * 4, 256);
    w.maxdc = (float *)(p + off);
    off += round_up(n * 4, 256);
    w.bytes = off;
    return w;
}


__device__ __forceinline__ int floordiv_i(int a, int b)
{
    int q = a / b;
    return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// C cast float -> integer as the Cython code does (`<Py_ssize_t>`): truncation
__device__ __forceinline__ int trunc_i(float v) { return (int)v; }

}  // namespace obia

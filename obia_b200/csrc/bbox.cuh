// Per-label bounding boxes and pixel counts of a label raster: one pass (4 B/pixel), one set of
// atomics per run of equal labels inside a warp.  Shared by K4 (zonal.cu) and the texture
// kernel (texture.cu); `static` so every translation unit gets its own copy.
#pragma once
#include "common.cuh"

namespace obia {

struct ZonalWs {
    int32_t *xmin, *xmax, *ymin, *ymax, *count;
    int64_t bytes;
};

static inline ZonalWs zonal_ws_layout(void *base, int64_t max_label)
{
    ZonalWs w;
    const int64_t n = max_label + 1;
    char *p = (char *)base;
    const int64_t stride = round_up(n * 4, 256);
    w.xmin = (int32_t *)(p);
    w.xmax = (int32_t *)(p + stride);
    w.ymin = (int32_t *)(p + 2 * stride);
    w.ymax = (int32_t *)(p + 3 * stride);
    w.count = (int32_t *)(p + 4 * stride);
    w.bytes = 5 * stride;
    return w;
}

static __global__ void zonal_init_kernel(ZonalWs w, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    w.xmin[i] = 0x7fffffff;
    w.ymin[i] = 0x7fffffff;
    w.xmax[i] = -1;
    w.ymax[i] = -1;
    w.count[i] = 0;
}

// grid: (ceil(W / 256), rows) -- a CTA covers 256 columns of the rows blockIdx.y, blockIdx.y + gridDim.y, ...
// (no per-thread division of a 64-bit pixel index); launch through zonal_bbox_launch.
static __global__ void __launch_bounds__(256)
zonal_bbox_kernel(const int32_t *__restrict__ labels, ZonalWs w, int H, int W, int64_t max_label,
                  int32_t label_lo = 0, int32_t zero_row = 0)
{
    // table row = label - label_lo (label_lo != 0: the rank-local label range of a sharded raster);
    // with zero_row the table starts with one extra row for label 0: row = label - label_lo + 1
    const int lane = threadIdx.x & 31;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const int64_t i = (int64_t)y * W + x;
        int32_t l = -1;
        if (x < W) {
            l = labels[i];
            const bool is_zero = zero_row && l == 0;
            if (!is_zero && (l < 0 || l < label_lo || (int64_t)l - label_lo + zero_row > max_label)) l = -1;
        }
        const int32_t prev = __shfl_up_sync(0xffffffffu, l, 1);
        const bool is_head = (lane == 0) || (prev != l);
        const unsigned heads = __ballot_sync(0xffffffffu, is_head);
        if (is_head && l >= 0) {
            const unsigned later = (lane == 31) ? 0u : (heads & ~((2u << lane) - 1u));
            const int end = later ? (__ffs(later) - 1) : 32;
            const int len = end - lane;
            const int32_t r = (zero_row && l == 0) ? 0 : l - label_lo + zero_row;
            atomicMin(w.xmin + r, x);
            atomicMax(w.xmax + r, x + len - 1);
            // a row can only be the label's first / last one if the pixel above / below the run head
            // carries another label (otherwise a smaller / larger y is reported by that row)
            if (y == 0 || labels[i - W] != l) atomicMin(w.ymin + r, y);
            if (y + 1 >= H || labels[i + W] != l) atomicMax(w.ymax + r, y);
            atomicAdd(w.count + r, len);
        }
    }
}

static inline void zonal_bbox_launch(const int32_t *labels, const ZonalWs &w, int64_t H, int64_t W, int64_t max_label,
                                     int32_t label_lo, int32_t zero_row, cudaStream_t st)
{
    dim3 grid((unsigned)ceil_div(W, 256), (unsigned)std::min<int64_t>(H, 32768));
    zonal_bbox_kernel<<<grid, 256, 0, st>>>(labels, w, (int)H, (int)W, max_label, label_lo, zero_row);
}

}  // namespace obia

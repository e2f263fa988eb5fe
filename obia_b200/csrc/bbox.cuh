// Per-label bounding boxes and pixel counts of a label raster: one pass (4 B/pixel), runs of equal labels
// folded per CTA tile in shared memory, one set of global atomics per distinct label of a tile.  Shared by K4 (zonal.cu) and the texture
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

// A CTA covers a tile of 256 columns x kBoxRows rows.  Every thread first loads the kBoxRows labels of its column
// (independent loads, all in flight), then the tile's runs of equal labels inside a warp row are folded into a
// shared-memory hash table (label -> slot: box and count, shared atomics), and one set of global atomics leaves
// the CTA per DISTINCT label of the tile instead of one per run (a fragmented SLIC raster has a run every three
// pixels: 39 M RED sectors per 10^8 pixels before, profiles/r02_all_kernels_ncu_raw.csv).  Runs that find the
// table full go to global memory directly.
constexpr int kBoxRows = 16;
constexpr int kBoxSlots = 512;

static __global__ void __launch_bounds__(256)
zonal_bbox_kernel(const int32_t *__restrict__ labels, ZonalWs w, int H, int W, int64_t max_label,
                  int32_t label_lo = 0, int32_t zero_row = 0, int y_base = 0)
{
    // (`labels` / H = the rows [y_base, y_base + H) of the raster: a grid has at most 65535 rows of tiles)
    // table row = label - label_lo (label_lo != 0: the rank-local label range of a sharded raster);
    // with zero_row the table starts with one extra row for label 0: row = label - label_lo + 1
    __shared__ int32_t s_key[kBoxSlots], s_x0[kBoxSlots], s_x1[kBoxSlots], s_y0[kBoxSlots], s_y1[kBoxSlots],
        s_n[kBoxSlots];
    const int lane = threadIdx.x & 31;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int yb = blockIdx.y * kBoxRows;
    for (int i = threadIdx.x; i < kBoxSlots; i += blockDim.x) {
        s_key[i] = -1;
        s_x0[i] = 0x7fffffff;
        s_y0[i] = 0x7fffffff;
        s_x1[i] = -1;
        s_y1[i] = -1;
        s_n[i] = 0;
    }
    int32_t lab[kBoxRows];
#pragma unroll
    for (int r = 0; r < kBoxRows; ++r) {
        const int y = yb + r;
        int32_t l = -1;
        if (x < W && y < H) {
            l = labels[(int64_t)y * W + x];
            const bool is_zero = zero_row && l == 0;
            if (!is_zero && (l < 0 || l < label_lo || (int64_t)l - label_lo + zero_row > max_label)) l = -1;
        }
        lab[r] = l;
    }
    __syncthreads();
    int32_t last_l = -1, last_slot = -1;
#pragma unroll
    for (int r = 0; r < kBoxRows; ++r) {
        const int y = y_base + yb + r;
        const int32_t l = lab[r];
        const int32_t prev = __shfl_up_sync(0xffffffffu, l, 1);
        const bool is_head = (lane == 0) || (prev != l);
        const unsigned heads = __ballot_sync(0xffffffffu, is_head);
        if (is_head && l >= 0) {
            const unsigned later = (lane == 31) ? 0u : (heads & ~((2u << lane) - 1u));
            const int end = later ? (__ffs(later) - 1) : 32;
            const int len = end - lane;
            int slot = (l == last_l) ? last_slot : -1;
            if (slot < 0) {
                int s = (int)(((uint32_t)l * 2654435761u) >> 23) & (kBoxSlots - 1);
                for (int probe = 0; probe < 64; ++probe) {
                    int32_t k = s_key[s];
                    if (k == -1) k = atomicCAS(&s_key[s], -1, l);
                    if (k == -1 || k == l) {
                        slot = s;
                        break;
                    }
                    s = (s + 1) & (kBoxSlots - 1);
                }
                last_l = l;
                last_slot = slot;
            }
            if (slot >= 0) {
                atomicMin(&s_x0[slot], x);
                atomicMax(&s_x1[slot], x + len - 1);
                atomicMin(&s_y0[slot], y);
                atomicMax(&s_y1[slot], y);
                atomicAdd(&s_n[slot], len);
            } else {
                const int32_t row = (zero_row && l == 0) ? 0 : l - label_lo + zero_row;
                atomicMin(w.xmin + row, x);
                atomicMax(w.xmax + row, x + len - 1);
                atomicMin(w.ymin + row, y);
                atomicMax(w.ymax + row, y);
                atomicAdd(w.count + row, len);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kBoxSlots; i += blockDim.x) {
        const int32_t l = s_key[i];
        if (l < 0) continue;
        const int32_t row = (zero_row && l == 0) ? 0 : l - label_lo + zero_row;
        atomicMin(w.xmin + row, s_x0[i]);
        atomicMax(w.xmax + row, s_x1[i]);
        atomicMin(w.ymin + row, s_y0[i]);
        atomicMax(w.ymax + row, s_y1[i]);
        atomicAdd(w.count + row, s_n[i]);
    }
}

static inline void zonal_bbox_launch(const int32_t *labels, const ZonalWs &w, int64_t H, int64_t W, int64_t max_label,
                                     int32_t label_lo, int32_t zero_row, cudaStream_t st)
{
    // (H * W < 2^31 is checked by the callers; rasters taller than 65535 tiles go in several launches)
    const int64_t rows_per_launch = (int64_t)65535 * kBoxRows;
    for (int64_t y0 = 0; y0 < H; y0 += rows_per_launch) {
        const int64_t h = std::min<int64_t>(rows_per_launch, H - y0);
        dim3 grid((unsigned)ceil_div(W, 256), (unsigned)ceil_div(h, kBoxRows));
        zonal_bbox_kernel<<<grid, 256, 0, st>>>(labels + y0 * W, w, (int)h, (int)W, max_label, label_lo, zero_row, (int)y0);
    }
}

}  // namespace obia

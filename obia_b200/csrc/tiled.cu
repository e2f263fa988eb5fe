// Tiled driver on the device (replaces the per-tile Python / shapely / GDAL logic of
// obia/utils/tiling.py:103-287): the seam bookkeeping of `create_tiled_segments` over a global HANDLE raster.
//
// A segment is its set of pixels: G[y][x] = handle of the segment covering the pixel (-1 none); per handle the
// pixel count `sizes`, a `live` flag and the 64-bit creation key (pass | tile row | tile column | local label)
// that orders the final 1..N numbering (tiling.py:289-290).  All windows of a batch (batch.cuh) are handled by
// one launch each:
//   paint   (tiling.py:145-147, :282)  the labels of freshly segmented windows become new handles in G;
//   prepare (tiling.py:182-260)        for every white window of a tile-row: count each earlier segment's pixels
//           inside the window polygon (window minus the two bottom corner squares); count == size <=> `within`
//           -> deleted (re-segmented), 0 < count < size <=> `overlaps` -> frozen (masked out); the window's
//           SLIC mask is written straight into the slab.
#include "batch.cuh"
#include "common.cuh"

namespace obia {

// labels: K3 output on the slab (mask label / -1 on masked pixels).  Window pixel with label L >= start_label
// -> handle hbase + L - start_label; the leftover label 0 of start_label 1 (obia treats it as one more segment
// per tile, SURVEY.md section 8 defect 7) -> handle hzero + window.
__global__ void __launch_bounds__(256)
win_paint_kernel(const int32_t *__restrict__ labels, const uint8_t *__restrict__ mask_slab, int slab_w,
                 const WinDesc *__restrict__ batch, const uint8_t *__restrict__ usable,
                 const int32_t *__restrict__ first_label, int start_label, int32_t hbase, int32_t hzero,
                 int32_t *__restrict__ G, int64_t GW, int32_t *sizes, int hw_max)
{
    const WinDesc &d = batch[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (!usable[blockIdx.y] || i >= d.h * d.w || i >= hw_max) return;
    const int y = i / d.w, x = i - y * d.w;
    const int64_t sp = (int64_t)(d.row0 + y) * slab_w + x;
    if (mask_slab && !mask_slab[sp]) return;
    int32_t L = labels[sp];
    // start_label 0: -2 = merged into "label 0", the first kept piece of this window (its own segment when the
    // window kept none: first_label < 0)
    if (L == -2) L = (first_label[blockIdx.y] >= 0) ? first_label[blockIdx.y] : start_label - 1;
    else if (L < 0) return;
    const int32_t h = (L >= start_label) ? hbase + (L - start_label) : hzero + (int32_t)blockIdx.y;
    G[(int64_t)(d.y0 + y) * GW + d.x0 + x] = h;
    atomicAdd(sizes + h, 1);
}

__device__ __forceinline__ bool in_polygon(int y, int x, int h, int w, int corner)
{
    return !(y >= h - corner && (x < corner || x >= w - corner));
}

__global__ void __launch_bounds__(256)
white_count_kernel(const int32_t *__restrict__ G, int64_t GW, const WinDesc *__restrict__ batch,
                   const uint8_t *__restrict__ parity, int corner, int32_t *cnt, int64_t cap, int32_t *any, int hw_max)
{
    const WinDesc &d = batch[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.h * d.w || i >= hw_max) return;
    const int y = i / d.w, x = i - y * d.w;
    if (!in_polygon(y, x, d.h, d.w, corner)) return;
    const int32_t g = G[(int64_t)(d.y0 + y) * GW + d.x0 + x];
    if (g < 0) return;
    atomicAdd(cnt + (int64_t)parity[blockIdx.y] * cap + g, 1);
    if (any[blockIdx.y] == 0) any[blockIdx.y] = 1;
}

// deleted pixels are left as -(handle + 2) for the reset pass
__global__ void __launch_bounds__(256)
white_apply_kernel(int32_t *__restrict__ G, int64_t GW, const WinDesc *__restrict__ batch,
                   const uint8_t *__restrict__ parity, int corner, const int32_t *__restrict__ cnt, int64_t cap,
                   const int32_t *__restrict__ sizes, uint8_t *live, const int32_t *__restrict__ any,
                   const uint8_t *__restrict__ user_mask, int64_t Wm, uint8_t *__restrict__ mask_slab, int slab_w,
                   int hw_max)
{
    const WinDesc &d = batch[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.h * d.w || i >= hw_max) return;
    const int y = i / d.w, x = i - y * d.w;
    const int64_t gp = (int64_t)(d.y0 + y) * GW + d.x0 + x;
    bool m = user_mask ? user_mask[(int64_t)(d.y0 + y) * Wm + d.x0 + x] != 0 : true;
    if (any[blockIdx.y]) {
        bool excluded = !in_polygon(y, x, d.h, d.w, corner);     // corner squares (tiling.py:245-246)
        const int32_t g = G[gp];
        if (g >= 0) {
            const int32_t c = cnt[(int64_t)parity[blockIdx.y] * cap + g];
            if (c > 0) {
                if (c == sizes[g]) {          // within the polygon: deleted, re-segmented (tiling.py:220-231)
                    G[gp] = -(g + 2);
                    live[g] = 0;
                } else {
                    excluded = true;          // overlaps its edge: frozen (tiling.py:213-218, :257-258)
                }
            }
        }
        m = m && !excluded;
    }
    mask_slab[(int64_t)(d.row0 + y) * slab_w + x] = m ? 1 : 0;
}

__global__ void __launch_bounds__(256)
white_reset_kernel(int32_t *__restrict__ G, int64_t GW, const WinDesc *__restrict__ batch,
                   const uint8_t *__restrict__ parity, int32_t *cnt, int64_t cap, int hw_max)
{
    const WinDesc &d = batch[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.h * d.w || i >= hw_max) return;
    const int y = i / d.w, x = i - y * d.w;
    const int64_t gp = (int64_t)(d.y0 + y) * GW + d.x0 + x;
    int32_t g = G[gp];
    if (g == -1) return;
    if (g < -1) {
        g = -g - 2;
        G[gp] = -1;
    }
    cnt[(int64_t)parity[blockIdx.y] * cap + g] = 0;
}

// ---- seam exchange (multi-GPU: the raster is sharded into column blocks, obia_b200/utils/tiling.py) -------------
// A neighbour's version of the 2*buffer-wide band on a block boundary replaces this rank's.  The band arrives as
// three int64 planes per pixel: creation key (-1 = no segment), segment size, and the segment's home
// (owner rank << 32 | handle on that rank).  Segments homed here keep their handle; segments homed on the
// neighbour are looked up in `mirror` (neighbour handle -> local handle, -1 unknown) and get a fresh local handle
// on first sight.  Three launches, no host round trip:
//   retire  live = 0 for every handle currently in the band (white rows: what does not come back was deleted)
//   claim   every pixel of an unknown neighbour segment draws a slot; one of them wins the mirror entry
//   import  G, keys, sizes, live, home from the planes through the (now complete) mirror
__global__ void __launch_bounds__(256)
seam_retire_kernel(const int32_t *__restrict__ G, int64_t GW, int rows, int cols, uint8_t *live)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int32_t g = G[(int64_t)(i / cols) * GW + i % cols];
    if (g >= 0) live[g] = 0;
}

__global__ void __launch_bounds__(256)
seam_claim_kernel(const long long *__restrict__ planes, int n, int me, int32_t *mirror, int64_t mirror_cap,
                  int32_t slot_base, int32_t *counter, int32_t *err)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || planes[i] < 0) return;
    const long long home = planes[2 * (int64_t)n + i];
    if ((int)(home >> 32) == me) return;
    const int64_t fh = home & 0xffffffffLL;
    if (fh >= mirror_cap) {
        *err = 1;
        return;
    }
    if (mirror[fh] == -1) atomicCAS(mirror + fh, -1, slot_base + atomicAdd(counter, 1));
}

__global__ void __launch_bounds__(256)
seam_import_kernel(const long long *__restrict__ planes, int rows, int cols, int me, const int32_t *__restrict__ mirror,
                   int64_t mirror_cap, int32_t *__restrict__ G, int64_t GW, long long *keys, int32_t *sizes,
                   uint8_t *live, long long *homes)
{
    const int n = rows * cols;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long key = planes[i];
    int32_t h = -1;
    if (key >= 0) {
        const long long home = planes[2 * (int64_t)n + i];
        const int64_t fh = home & 0xffffffffLL;
        if ((int)(home >> 32) == me) {
            h = (int32_t)fh;
        } else if (fh < mirror_cap) {
            h = mirror[fh];
            if (h >= 0) {
                keys[h] = key;
                sizes[h] = (int32_t)planes[(int64_t)n + i];
                homes[h] = home;
            }
        }
        if (h >= 0) live[h] = 1;
    }
    G[(int64_t)(i / cols) * GW + i % cols] = h;
}

}  // namespace obia

using namespace obia;

extern "C" int obia_b200_tiled_paint(const int32_t *labels_slab, const uint8_t *mask_slab, int32_t slab_w,
                                     const void *descs, const uint8_t *usable, const int32_t *first_label, int64_t B,
                                     int32_t hmax, int32_t wmax, int32_t start_label, int32_t handle_base,
                                     int32_t handle_zero, int32_t *G, int64_t GW, int32_t *sizes, void *stream)
{
    if (!labels_slab || !descs || !usable || !first_label || !G || !sizes || B <= 0 || B > 65535 || hmax <= 0 || wmax <= 0 ||
        slab_w < wmax || GW <= 0)
        return set_err(OBIA_B200_ERR_ARG, "tiled_paint: bad argument");
    dim3 grid((unsigned)ceil_div((int64_t)hmax * wmax, 256), (unsigned)B);
    win_paint_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(labels_slab, mask_slab, slab_w, (const WinDesc *)descs,
                                                              usable, first_label, start_label, handle_base, handle_zero, G, GW,
                                                              sizes,
                                                              hmax * wmax);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

// counts: (2, capacity) int32 scratch, all zero on entry and on return; any_segment: (B) int32 scratch;
// parity: (B) 0 / 1, different for neighbouring windows of the row; corner = pixels whose centre lies inside a
// corner square of side buffer / 2; user_mask: (H, Wm) or NULL.
extern "C" int obia_b200_tiled_white_prepare(int32_t *G, int64_t GW, const void *descs, const uint8_t *parity,
                                             int64_t B, int32_t hmax, int32_t wmax, int32_t corner, int32_t *counts,
                                             int64_t capacity, const int32_t *sizes, uint8_t *live,
                                             int32_t *any_segment, const uint8_t *user_mask, int64_t Wm,
                                             uint8_t *mask_slab, int32_t slab_w, void *stream)
{
    if (!G || !descs || !parity || !counts || !sizes || !live || !any_segment || !mask_slab || B <= 0 || B > 65535 ||
        hmax <= 0 || wmax <= 0 || slab_w < wmax || GW <= 0 || capacity <= 0 || corner < 0)
        return set_err(OBIA_B200_ERR_ARG, "tiled_white_prepare: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const WinDesc *batch = (const WinDesc *)descs;
    dim3 grid((unsigned)ceil_div((int64_t)hmax * wmax, 256), (unsigned)B);
    const int hw = hmax * wmax;
    OBIA_CUDA_CHECK(cudaMemsetAsync(any_segment, 0, (size_t)B * 4, st));
    white_count_kernel<<<grid, 256, 0, st>>>(G, GW, batch, parity, corner, counts, capacity, any_segment, hw);
    OBIA_LAUNCH_CHECK();
    white_apply_kernel<<<grid, 256, 0, st>>>(G, GW, batch, parity, corner, counts, capacity, sizes, live, any_segment,
                                             user_mask, Wm, mask_slab, slab_w, hw);
    OBIA_LAUNCH_CHECK();
    white_reset_kernel<<<grid, 256, 0, st>>>(G, GW, batch, parity, counts, capacity, hw);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

// band: the rows x cols block of G at `G_band` (row stride GW); planes: (3, rows, cols) int64 as received;
// mirror: (mirror_cap) int32 for this neighbour, -1 = unknown; new local handles are slot_base + [0, rows * cols):
// the caller reserves that many table slots (counter: device int32, zeroed here; err: device int32, set to 1 when a
// neighbour handle exceeds mirror_cap).  retire != 0: white-row exchange (segments that vanish were deleted).
extern "C" int obia_b200_tiled_seam_import(int32_t *G_band, int64_t GW, int32_t rows, int32_t cols,
                                           const int64_t *planes, int32_t my_rank, int32_t retire, int32_t *mirror,
                                           int64_t mirror_cap, int32_t slot_base, int32_t *counter, int32_t *err,
                                           int64_t *keys, int32_t *sizes, uint8_t *live, int64_t *homes, void *stream)
{
    if (!G_band || !planes || !mirror || !counter || !err || !keys || !sizes || !live || !homes || rows <= 0 ||
        cols <= 0 || GW < cols || mirror_cap <= 0 || slot_base < 0)
        return set_err(OBIA_B200_ERR_ARG, "tiled_seam_import: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int n = rows * cols;
    const unsigned grid = (unsigned)ceil_div(n, 256);
    OBIA_CUDA_CHECK(cudaMemsetAsync(counter, 0, 4, st));
    if (retire) {
        seam_retire_kernel<<<grid, 256, 0, st>>>(G_band, GW, rows, cols, live);
        OBIA_LAUNCH_CHECK();
    }
    seam_claim_kernel<<<grid, 256, 0, st>>>((const long long *)planes, n, my_rank, mirror, mirror_cap, slot_base, counter,
                                            err);
    OBIA_LAUNCH_CHECK();
    seam_import_kernel<<<grid, 256, 0, st>>>((const long long *)planes, rows, cols, my_rank, mirror, mirror_cap, G_band, GW,
                                             (long long *)keys, sizes, live, (long long *)homes);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

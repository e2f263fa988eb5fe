// maskSLIC initialisation, the k-means part: replaces
// `scipy.cluster.vq.kmeans2(coord[idx_dense], coord[idx], iter=5)` called by scikit-image's
// `_get_mask_centroids` (reached from obia/segmentation/segment_boundaries.py:51 whenever a mask
// is passed, i.e. for every tile of obia/utils/tiling.py:137-143, :275-281).
//
// Same arithmetic as scipy's `_vq.vq` for fewer than 5 features: squared distance accumulated in
// float64 in feature order (z, y, x; z is 0 for 2-D rasters), strict `<` so the lowest code index
// wins ties; `update_cluster_means`: per-cluster sums of the (integer-valued) coordinates divided
// by the member count, empty clusters keep their previous centroid.  The sums are exact integers,
// so 64-bit integer atomics reproduce scipy's sequential float64 sums bit for bit.
#include "batch.cuh"
#include "common.cuh"

namespace obia {

constexpr int kKmChunk = 1024;  // centroids staged in shared memory at a time

__global__ void __launch_bounds__(256)
kmeans_assign_kernel(const int32_t *__restrict__ pts, int64_t m, const double *__restrict__ cent, int n,
                     unsigned long long *__restrict__ sums /* [n][3]: count, sum y, sum x */)
{
    __shared__ double s_cy[kKmChunk], s_cx[kKmChunk];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double py = 0.0, px = 0.0;
    int iy = 0, ix = 0;
    if (i < m) {
        iy = pts[2 * i];
        ix = pts[2 * i + 1];
        py = (double)iy;
        px = (double)ix;
    }
    double best = __longlong_as_double(0x7ff0000000000000LL);  // +inf
    int bestj = 0;
    for (int c0 = 0; c0 < n; c0 += kKmChunk) {
        const int nc = min(kKmChunk, n - c0);
        __syncthreads();
        for (int j = threadIdx.x; j < nc; j += blockDim.x) {
            s_cy[j] = cent[2 * (c0 + j)];
            s_cx[j] = cent[2 * (c0 + j) + 1];
        }
        __syncthreads();
        if (i < m) {
            for (int j = 0; j < nc; ++j) {
                const double dy = __dsub_rn(py, s_cy[j]);
                const double dx = __dsub_rn(px, s_cx[j]);
                const double d = __dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dx, dx));
                if (d < best) {
                    best = d;
                    bestj = c0 + j;
                }
            }
        }
    }
    if (i < m) {
        atomicAdd(&sums[3 * (int64_t)bestj + 0], 1ull);
        atomicAdd(&sums[3 * (int64_t)bestj + 1], (unsigned long long)(long long)iy);
        atomicAdd(&sums[3 * (int64_t)bestj + 2], (unsigned long long)(long long)ix);
    }
}

// ---- uniform grid over the centroids: exact nearest-centroid search for large n --------------
// Brute force is n distance evaluations per point (a 2000 x 2000 tile has 2e6 points and 2e4
// centroids: 2e11 float64 evaluations per sweep set).  Centroids are binned into square cells of
// about one centroid each (linked lists, one atomicExch per centroid); a point visits the cells
// ring by ring around its own and stops once the best squared distance is strictly below the
// squared distance to the border of the block already visited -- every unvisited centroid is then
// strictly farther, so the result (including the lowest-index tie rule) is the brute-force one.
struct KmGrid {
    double cs;          // cell size
    int ncy, ncx;
    int32_t *head;      // [ncy * ncx] first centroid of the cell or -1
    int32_t *next;      // [n] next centroid in the same cell or -1
};

static KmGrid km_grid_layout(void *base_after_sums, int64_t n, int64_t ey, int64_t ex)
{
    KmGrid g;
    g.cs = sqrt((double)ey * (double)ex / (double)n);
    if (g.cs < 1.0) g.cs = 1.0;
    g.ncy = (int)((double)ey / g.cs) + 1;
    g.ncx = (int)((double)ex / g.cs) + 1;
    char *p = (char *)base_after_sums;
    g.head = (int32_t *)p;
    g.next = (int32_t *)(p + round_up((int64_t)g.ncy * g.ncx * 4, 256));
    return g;
}

__device__ __forceinline__ int km_cell(double v, double cs, int nc)
{
    int c = (int)floor(v / cs);
    return c < 0 ? 0 : (c >= nc ? nc - 1 : c);
}

__global__ void km_bin_kernel(const double *__restrict__ cent, int n, KmGrid g)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int gy = km_cell(cent[2 * k], g.cs, g.ncy), gx = km_cell(cent[2 * k + 1], g.cs, g.ncx);
    g.next[k] = atomicExch(&g.head[gy * g.ncx + gx], k);
}

// Nearest centroid of (py, px) in scipy's arithmetic.  SQRT: compare Euclidean distances like
// pdist (sqrt of the float64 sum) instead of squared ones; `self` is excluded (-1: nothing is).
template <bool SQRT>
__device__ __forceinline__ int km_nearest(double py, double px, int self, const double *__restrict__ cent,
                                          const KmGrid &g)
{
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const int gy = km_cell(py, g.cs, g.ncy), gx = km_cell(px, g.cs, g.ncx);
    double best = INF, best2 = INF;   // compared value, and its squared distance for the bound
    int bestk = 0x7fffffff;
    auto visit = [&](int cy, int cx) {
        for (int k = g.head[cy * g.ncx + cx]; k >= 0; k = g.next[k]) {
            if (k == self) continue;
            const double dy = __dsub_rn(py, cent[2 * k]);
            const double dx = __dsub_rn(px, cent[2 * k + 1]);
            const double d2 = __dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dx, dx));
            const double d = SQRT ? __dsqrt_rn(d2) : d2;
            if (d < best || (d == best && k < bestk)) {
                best = d;
                best2 = d2;
                bestk = k;
            }
        }
    };
    const int rmax = max(g.ncy, g.ncx);
    for (int r = 0; r <= rmax; ++r) {
        const int y0 = gy - r, y1 = gy + r, x0 = gx - r, x1 = gx + r;
        if (r == 0) {
            visit(gy, gx);
        } else {
            for (int cx = max(x0, 0); cx <= min(x1, g.ncx - 1); ++cx) {
                if (y0 >= 0) visit(y0, cx);
                if (y1 < g.ncy) visit(y1, cx);
            }
            for (int cy = max(y0 + 1, 0); cy <= min(y1 - 1, g.ncy - 1); ++cy) {
                if (x0 >= 0) visit(cy, x0);
                if (x1 < g.ncx) visit(cy, x1);
            }
        }
        // distance from the point to the border of the visited block (sides beyond the grid: none)
        double lb = INF;
        if (y0 > 0) lb = fmin(lb, py - (double)y0 * g.cs);
        if (y1 < g.ncy - 1) lb = fmin(lb, (double)(y1 + 1) * g.cs - py);
        if (x0 > 0) lb = fmin(lb, px - (double)x0 * g.cs);
        if (x1 < g.ncx - 1) lb = fmin(lb, (double)(x1 + 1) * g.cs - px);
        if (lb == INF) break;          // the whole grid has been visited
        lb -= 1e-6;                    // slack for the rounding of cell indices / borders
        if (lb > 0.0 && best2 < lb * lb * (1.0 - 1e-9)) break;
    }
    return bestk;
}

__global__ void __launch_bounds__(256)
kmeans_assign_grid_kernel(const int32_t *__restrict__ pts, int64_t m, const double *__restrict__ cent, KmGrid g,
                          unsigned long long *__restrict__ sums)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int iy = pts[2 * i], ix = pts[2 * i + 1];
    const int bestj = km_nearest<false>((double)iy, (double)ix, -1, cent, g);
    atomicAdd(&sums[3 * (int64_t)bestj + 0], 1ull);
    atomicAdd(&sums[3 * (int64_t)bestj + 1], (unsigned long long)(long long)iy);
    atomicAdd(&sums[3 * (int64_t)bestj + 2], (unsigned long long)(long long)ix);
}

// nearest OTHER centroid of every centroid: `squareform(pdist(c)); fill_diagonal(inf); argmin(-1)`
__global__ void __launch_bounds__(256)
km_closest_kernel(const double *__restrict__ cent, int n, KmGrid g, int32_t *__restrict__ closest)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int j = km_nearest<true>(cent[2 * k], cent[2 * k + 1], k, cent, g);
    closest[k] = (j == 0x7fffffff) ? 0 : j;   // n == 1: argmin of [[inf]] is 0
}

__global__ void kmeans_update_kernel(double *cent, unsigned long long *sums, int n)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const long long cnt = (long long)sums[3 * j];
    if (cnt > 0) {
        cent[2 * j] = (double)(long long)sums[3 * j + 1] / (double)cnt;
        cent[2 * j + 1] = (double)(long long)sums[3 * j + 2] / (double)cnt;
    }  // empty cluster: keep the previous position (kmeans2 missing='warn')
    sums[3 * j] = sums[3 * j + 1] = sums[3 * j + 2] = 0ull;
}

// ---- batched form (tiled driver, batch.cuh): every window of a slab at once ------------------------------
// Points are slab pixel positions (row * slab_w + column; window = row / win_rows, its rows start at
// win * win_rows); centroids are window-local float64 (y, x), indexed batch-wide; `head` / `next` hold
// batch-wide centroid indices, each window owning the cells [km_cell0, km_cell0 + km_ncy * km_ncx).
__device__ __forceinline__ KmGrid km_grid_of(const WinDesc &d, int32_t *head, int32_t *next)
{
    KmGrid g;
    g.cs = d.km_cs;
    g.ncy = d.km_ncy;
    g.ncx = d.km_ncx;
    g.head = head + d.km_cell0;
    g.next = next;
    return g;
}

__global__ void __launch_bounds__(256)
km_batch_seed_kernel(const int32_t *__restrict__ seed_pos, const int32_t *__restrict__ cwin,
                     const WinDesc *__restrict__ batch, int64_t n_total, int slab_w, double *__restrict__ cent)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_total) return;
    const WinDesc &d = batch[cwin[k]];
    const int32_t pos = seed_pos[k];
    cent[2 * k] = (double)(pos / slab_w - d.row0);
    cent[2 * k + 1] = (double)(pos % slab_w);
}

__global__ void __launch_bounds__(256)
km_batch_bin_kernel(const double *__restrict__ cent, const int32_t *__restrict__ cwin, const WinDesc *__restrict__ batch,
                    int64_t n_total, int32_t *head, int32_t *next)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_total) return;
    const WinDesc &d = batch[cwin[k]];
    const int gy = km_cell(cent[2 * k], d.km_cs, d.km_ncy), gx = km_cell(cent[2 * k + 1], d.km_cs, d.km_ncx);
    next[k] = atomicExch(&head[d.km_cell0 + gy * d.km_ncx + gx], (int32_t)k);
}

__global__ void __launch_bounds__(256)
km_batch_assign_kernel(const int32_t *__restrict__ pts_pos, int64_t m_total, const double *__restrict__ cent,
                       const WinDesc *__restrict__ batch, int slab_w, int win_rows, int32_t *head, int32_t *next,
                       unsigned long long *__restrict__ sums)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m_total) return;
    const int32_t pos = pts_pos[i];
    const int row = pos / slab_w, ix = pos - row * slab_w;
    const WinDesc &d = batch[row / win_rows];
    if (!d.valid || d.n <= 0) return;
    const int iy = row - d.row0;
    const KmGrid g = km_grid_of(d, head, next);
    const int bestj = km_nearest<false>((double)iy, (double)ix, -1, cent, g);
    atomicAdd(&sums[3 * (int64_t)bestj + 0], 1ull);
    atomicAdd(&sums[3 * (int64_t)bestj + 1], (unsigned long long)(long long)iy);
    atomicAdd(&sums[3 * (int64_t)bestj + 2], (unsigned long long)(long long)ix);
}

// The same assignment for DENSE windows (coord[idx_dense] = every mask pixel, the usual case: 100 * n_segments
// >= mask pixels): one CTA per 16 x 16 pixel tile of a window.  The centroids of the cells around the tile are
// staged in shared memory once and every pixel searches them by brute force; the result is accepted only when
// the best distance is strictly inside the staged block (the bound km_nearest uses ring by ring), otherwise the
// pixel falls back to km_nearest -- both give the exact nearest centroid with the lowest-index tie rule.
// Sums are folded per tile in shared memory (integers) before they reach the global table.
constexpr int kKmTile = 16, kKmStage = 192;

__global__ void __launch_bounds__(256)
km_batch_assign_tile_kernel(const uint8_t *__restrict__ mask_slab, int slab_w, const double *__restrict__ cent,
                            const WinDesc *__restrict__ batch, int tiles_x, int32_t *head, int32_t *next,
                            unsigned long long *__restrict__ sums)
{
    __shared__ double s_cy[kKmStage], s_cx[kKmStage];
    __shared__ int s_k[kKmStage];
    __shared__ int s_sum[kKmStage][3];
    __shared__ int s_n;
    const WinDesc &d = batch[blockIdx.y];
    const int ty0 = (blockIdx.x / tiles_x) * kKmTile, tx0 = (blockIdx.x % tiles_x) * kKmTile;
    if (!d.valid || d.n <= 0 || ty0 >= d.h || tx0 >= d.w) return;
    const KmGrid g = km_grid_of(d, head, next);
    const int ty1 = min(ty0 + kKmTile, d.h) - 1, tx1 = min(tx0 + kKmTile, d.w) - 1;
    const int cy0 = max(0, km_cell((double)ty0, g.cs, g.ncy) - 1), cy1 = min(g.ncy - 1, km_cell((double)ty1, g.cs, g.ncy) + 1);
    const int cx0 = max(0, km_cell((double)tx0, g.cs, g.ncx) - 1), cx1 = min(g.ncx - 1, km_cell((double)tx1, g.cs, g.ncx) + 1);
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const int ncell = (cy1 - cy0 + 1) * (cx1 - cx0 + 1), cw = cx1 - cx0 + 1;
    for (int c = threadIdx.x; c < ncell; c += blockDim.x) {
        const int cy = cy0 + c / cw, cx = cx0 + c % cw;
        for (int k = g.head[cy * g.ncx + cx]; k >= 0; k = g.next[k]) {
            const int slot = atomicAdd(&s_n, 1);
            if (slot < kKmStage) {
                s_k[slot] = k;
                s_cy[slot] = cent[2 * k];
                s_cx[slot] = cent[2 * k + 1];
                s_sum[slot][0] = s_sum[slot][1] = s_sum[slot][2] = 0;
            }
        }
    }
    __syncthreads();
    const int ns = s_n;
    const int iy = ty0 + (int)threadIdx.x / kKmTile, ix = tx0 + (int)threadIdx.x % kKmTile;
    if (iy < d.h && ix < d.w && mask_slab[(int64_t)(d.row0 + iy) * slab_w + ix]) {
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
        const double py = (double)iy, px = (double)ix;
        int bestk = 0x7fffffff, bests = -1;
        bool exact = false;
        if (ns <= kKmStage) {
            double best = INF;
            for (int j = 0; j < ns; ++j) {
                const double dy = __dsub_rn(py, s_cy[j]);
                const double dx = __dsub_rn(px, s_cx[j]);
                const double d2 = __dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dx, dx));
                const int k = s_k[j];
                if (d2 < best || (d2 == best && k < bestk)) {
                    best = d2;
                    bestk = k;
                    bests = j;
                }
            }
            double lb = INF;
            if (cy0 > 0) lb = fmin(lb, py - (double)cy0 * g.cs);
            if (cy1 < g.ncy - 1) lb = fmin(lb, (double)(cy1 + 1) * g.cs - py);
            if (cx0 > 0) lb = fmin(lb, px - (double)cx0 * g.cs);
            if (cx1 < g.ncx - 1) lb = fmin(lb, (double)(cx1 + 1) * g.cs - px);
            if (lb == INF) {
                exact = bests >= 0;                    // the whole grid was staged
            } else {
                lb -= 1e-6;
                exact = bests >= 0 && lb > 0.0 && best < lb * lb * (1.0 - 1e-9);
            }
        }
        if (exact) {
            atomicAdd(&s_sum[bests][0], 1);
            atomicAdd(&s_sum[bests][1], iy);
            atomicAdd(&s_sum[bests][2], ix);
        } else {
            const int k = km_nearest<false>(py, px, -1, cent, g);
            atomicAdd(&sums[3 * (int64_t)k + 0], 1ull);
            atomicAdd(&sums[3 * (int64_t)k + 1], (unsigned long long)(long long)iy);
            atomicAdd(&sums[3 * (int64_t)k + 2], (unsigned long long)(long long)ix);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < min(ns, kKmStage); j += blockDim.x) {
        if (s_sum[j][0] == 0) continue;
        const int64_t k = s_k[j];
        atomicAdd(&sums[3 * k + 0], (unsigned long long)s_sum[j][0]);
        atomicAdd(&sums[3 * k + 1], (unsigned long long)s_sum[j][1]);
        atomicAdd(&sums[3 * k + 2], (unsigned long long)s_sum[j][2]);
    }
}

__global__ void __launch_bounds__(256)
km_batch_closest_kernel(const double *__restrict__ cent, const int32_t *__restrict__ cwin,
                        const WinDesc *__restrict__ batch, int64_t n_total, int32_t *head, int32_t *next,
                        int32_t *__restrict__ closest)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_total) return;
    const WinDesc &d = batch[cwin[k]];
    const KmGrid g = km_grid_of(d, head, next);
    const int j = km_nearest<true>(cent[2 * k], cent[2 * k + 1], (int)k, cent, g);
    closest[k] = (j == 0x7fffffff) ? d.c0 : j;   // n == 1: argmin of [[inf]] is 0
}

// `steps = np.abs(centroids - centroids[closest]).mean(0)` (row-sequential float64 sums), `step = max(steps)`,
// the spatial weight `1.0 / step ** 2` as slic.cu derives it; a window whose step is not positive is dropped
// (the reference raises ValueError there: an "empty tile").  One WARP per window: the lanes fetch 32 terms at a
// time, the sum itself runs in index order (every lane adds the same sequence, so the rounding is numpy's).
__global__ void __launch_bounds__(128)
km_batch_steps_kernel(const double *__restrict__ cent, const int32_t *__restrict__ closest, WinDesc *batch, int64_t B)
{
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= B) return;
    WinDesc &d = batch[w];
    if (!d.valid) return;
    double sy = 0.0, sx = 0.0;
    for (int i0 = 0; i0 < d.n; i0 += 32) {
        double ay = 0.0, ax = 0.0;
        if (i0 + lane < d.n) {
            const int64_t k = (int64_t)d.c0 + i0 + lane, j = closest[k];
            ay = fabs(__dsub_rn(cent[2 * k], cent[2 * j]));
            ax = fabs(__dsub_rn(cent[2 * k + 1], cent[2 * j + 1]));
        }
        const int cnt = min(32, d.n - i0);
        for (int l = 0; l < cnt; ++l) {
            sy = __dadd_rn(sy, __shfl_sync(0xffffffffu, ay, l));
            sx = __dadd_rn(sx, __shfl_sync(0xffffffffu, ax, l));
        }
    }
    if (lane != 0) return;
    sy = __ddiv_rn(sy, (double)d.n);
    sx = __ddiv_rn(sx, (double)d.n);
    const float step = (float)fmax(0.0, fmax(sy, sx));
    if (!(step > 0.0f)) {
        d.valid = 0;
        return;
    }
    const float step_sq = __fmul_rn(step, step);
    d.sw = (float)(1.0 / (double)step_sq);
    d.inv_w = __fdiv_rn(1.0f, d.sw);
}

// SLIC centre table rows (cy, cx, 0 ...) from the float64 centroids
__global__ void __launch_bounds__(256)
km_batch_centres_kernel(const double *__restrict__ cent, int64_t n_total, int Cf, float *__restrict__ centres)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_total) return;
    float *c = centres + k * (2 + Cf);
    c[0] = (float)cent[2 * k];
    c[1] = (float)cent[2 * k + 1];
    for (int f = 0; f < Cf; ++f) c[2 + f] = 0.0f;
}

}  // namespace obia

using namespace obia;

constexpr int64_t kKmBruteMax = 1024;   // up to here brute force over a shared-memory chunk is faster

extern "C" int64_t obia_b200_mask_kmeans_workspace_bytes(int64_t n, int64_t extent_y, int64_t extent_x)
{
    if (n <= 0 || extent_y <= 0 || extent_x <= 0) return -1;
    const int64_t sums = round_up(n * 3 * 8, 256);
    KmGrid g = km_grid_layout(nullptr, n, extent_y, extent_x);
    return sums + round_up((int64_t)g.ncy * g.ncx * 4, 256) + round_up(n * 4, 256);
}

extern "C" int obia_b200_mask_kmeans(const int32_t *points_yx, int64_t m, double *centroids_yx, int64_t n,
                                     int32_t iters, int64_t extent_y, int64_t extent_x, void *workspace,
                                     void *stream)
{
    if (!points_yx || !centroids_yx || !workspace || m <= 0 || n <= 0 || iters < 0 || n > 0x7fffffffLL ||
        extent_y <= 0 || extent_x <= 0)
        return set_err(OBIA_B200_ERR_ARG, "mask_kmeans: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *sums = (unsigned long long *)workspace;
    KmGrid g = km_grid_layout((char *)workspace + round_up(n * 3 * 8, 256), n, extent_y, extent_x);
    OBIA_CUDA_CHECK(cudaMemsetAsync(sums, 0, (size_t)n * 3 * 8, st));
    for (int it = 0; it < iters; ++it) {
        if (n <= kKmBruteMax) {
            kmeans_assign_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, st>>>(points_yx, m, centroids_yx, (int)n, sums);
            OBIA_LAUNCH_CHECK();
        } else {
            OBIA_CUDA_CHECK(cudaMemsetAsync(g.head, 0xff, (size_t)g.ncy * g.ncx * 4, st));
            km_bin_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(centroids_yx, (int)n, g);
            OBIA_LAUNCH_CHECK();
            kmeans_assign_grid_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, st>>>(points_yx, m, centroids_yx, g, sums);
            OBIA_LAUNCH_CHECK();
        }
        kmeans_update_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(centroids_yx, sums, (int)n);
        OBIA_LAUNCH_CHECK();
    }
    return OBIA_B200_OK;
}

extern "C" int obia_b200_nearest_centroid(const double *centroids_yx, int64_t n, int64_t extent_y,
                                          int64_t extent_x, int32_t *closest, void *workspace, void *stream)
{
    if (!centroids_yx || !closest || !workspace || n <= 0 || n > 0x7fffffffLL || extent_y <= 0 || extent_x <= 0)
        return set_err(OBIA_B200_ERR_ARG, "nearest_centroid: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    KmGrid g = km_grid_layout((char *)workspace + round_up(n * 3 * 8, 256), n, extent_y, extent_x);
    OBIA_CUDA_CHECK(cudaMemsetAsync(g.head, 0xff, (size_t)g.ncy * g.ncx * 4, st));
    km_bin_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(centroids_yx, (int)n, g);
    OBIA_LAUNCH_CHECK();
    km_closest_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(centroids_yx, (int)n, g, closest);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

// ---- batched maskSLIC initialisation ------------------------------------------------------------------------
extern "C" int64_t obia_b200_mask_kmeans_batch_workspace_bytes(int64_t n_total, int64_t km_cells_total)
{
    if (n_total <= 0 || km_cells_total <= 0) return -1;
    return round_up(n_total * 3 * 8, 256) + round_up(km_cells_total * 4, 256) + round_up(n_total * 4, 256) +
           round_up(n_total * 4, 256);
}

// k-means (`iters` sweeps from the seed pixels), nearest-other-centroid steps and the SLIC centre rows for
// every window of a slab.  descs (device, batch.cuh) supply n / c0 / km grid per window and receive sw, inv_w
// and valid = 0 for degenerate windows; centroids: (n_total, 2) float64 scratch, kept as the result.
extern "C" int obia_b200_mask_kmeans_batch(const int32_t *points_pos, int64_t m_total, const uint8_t *dense_mask_slab,
                                           int32_t hmax, int32_t wmax, const int32_t *seed_pos, const int32_t *cwin,
                                           void *descs, int64_t B, int64_t n_total, int64_t km_cells_total,
                                           int32_t slab_w, int32_t win_rows, int32_t iters, int32_t Cf,
                                           double *centroids, float *centres, void *workspace, void *stream)
{
    if ((!points_pos && !dense_mask_slab) || (dense_mask_slab && (hmax <= 0 || wmax <= 0 || B > 65535)) ||
        (!dense_mask_slab && m_total <= 0) || !seed_pos || !cwin || !descs || !centroids || !centres || !workspace || B <= 0 ||
        n_total <= 0 || km_cells_total <= 0 || slab_w <= 0 || win_rows <= 0 || iters < 0 || Cf <= 0)
        return set_err(OBIA_B200_ERR_ARG, "mask_kmeans_batch: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    char *p = (char *)workspace;
    unsigned long long *sums = (unsigned long long *)p;
    p += round_up(n_total * 3 * 8, 256);
    int32_t *head = (int32_t *)p;
    p += round_up(km_cells_total * 4, 256);
    int32_t *next = (int32_t *)p;
    p += round_up(n_total * 4, 256);
    int32_t *closest = (int32_t *)p;
    WinDesc *batch = (WinDesc *)descs;
    const unsigned gn = (unsigned)ceil_div(n_total, 256);
    km_batch_seed_kernel<<<gn, 256, 0, st>>>(seed_pos, cwin, batch, n_total, slab_w, centroids);
    OBIA_LAUNCH_CHECK();
    OBIA_CUDA_CHECK(cudaMemsetAsync(sums, 0, (size_t)n_total * 3 * 8, st));
    for (int it = 0; it < iters; ++it) {
        OBIA_CUDA_CHECK(cudaMemsetAsync(head, 0xff, (size_t)km_cells_total * 4, st));
        km_batch_bin_kernel<<<gn, 256, 0, st>>>(centroids, cwin, batch, n_total, head, next);
        OBIA_LAUNCH_CHECK();
        if (dense_mask_slab) {      // the points of every window are all of its mask pixels
            const int tiles_x = (int)ceil_div(wmax, kKmTile);
            dim3 grid((unsigned)(tiles_x * ceil_div(hmax, kKmTile)), (unsigned)B);
            km_batch_assign_tile_kernel<<<grid, 256, 0, st>>>(dense_mask_slab, slab_w, centroids, batch, tiles_x, head,
                                                              next, sums);
        } else {
            km_batch_assign_kernel<<<(unsigned)ceil_div(m_total, 256), 256, 0, st>>>(points_pos, m_total, centroids, batch,
                                                                                    slab_w, win_rows, head, next, sums);
        }
        OBIA_LAUNCH_CHECK();
        kmeans_update_kernel<<<gn, 256, 0, st>>>(centroids, sums, (int)n_total);
        OBIA_LAUNCH_CHECK();
    }
    OBIA_CUDA_CHECK(cudaMemsetAsync(head, 0xff, (size_t)km_cells_total * 4, st));
    km_batch_bin_kernel<<<gn, 256, 0, st>>>(centroids, cwin, batch, n_total, head, next);
    OBIA_LAUNCH_CHECK();
    km_batch_closest_kernel<<<gn, 256, 0, st>>>(centroids, cwin, batch, n_total, head, next, closest);
    OBIA_LAUNCH_CHECK();
    km_batch_steps_kernel<<<(unsigned)ceil_div(B * 32, 128), 128, 0, st>>>(centroids, closest, batch, B);
    OBIA_LAUNCH_CHECK();
    km_batch_centres_kernel<<<gn, 256, 0, st>>>(centroids, n_total, Cf, centres);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

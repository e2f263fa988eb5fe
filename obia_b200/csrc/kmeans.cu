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

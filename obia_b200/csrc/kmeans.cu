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

extern "C" int64_t obia_b200_mask_kmeans_workspace_bytes(int64_t n)
{
    if (n <= 0) return -1;
    return round_up(n * 3 * 8, 256);
}

extern "C" int obia_b200_mask_kmeans(const int32_t *points_yx, int64_t m, double *centroids_yx, int64_t n,
                                     int32_t iters, void *workspace, void *stream)
{
    if (!points_yx || !centroids_yx || !workspace || m <= 0 || n <= 0 || iters < 0 || n > 0x7fffffffLL)
        return set_err(OBIA_B200_ERR_ARG, "mask_kmeans: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *sums = (unsigned long long *)workspace;
    OBIA_CUDA_CHECK(cudaMemsetAsync(sums, 0, (size_t)n * 3 * 8, st));
    for (int it = 0; it < iters; ++it) {
        kmeans_assign_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, st>>>(points_yx, m, centroids_yx, (int)n, sums);
        OBIA_LAUNCH_CHECK();
        kmeans_update_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(centroids_yx, sums, (int)n);
        OBIA_LAUNCH_CHECK();
    }
    return OBIA_B200_OK;
}

// K4: per-segment, per-band zonal statistics (replaces the per-segment loop of
// obia/segmentation/segment_statistics.py:475-508 -> calculate_spectral_stats
// :113-176, i.e. np.mean / np.var / np.min / np.max / scipy.stats.skew /
// scipy.stats.kurtosis over the pixels of each segment).
//
// Two kernels:
//   zonal_bbox    one pass over the label raster (4 B/pixel): bounding box and
//                 pixel count per label, one set of atomics per run of equal
//                 labels inside a warp;
//   zonal_gather  one warp per (label, chunk of 8 bands): walks the label's
//                 bounding box with coalesced label/raster loads, accumulates
//                 pivot-shifted power sums in registers (float32 partials
//                 folded into float64 every 16 pixels), reduces them with
//                 shuffles and writes the finished statistics.  No atomics, no
//                 scratch accumulators, deterministic; SLIC segments are
//                 compact so neighbouring boxes overlap in L2, not in HBM.
#include "common.cuh"

namespace obia {

constexpr int kZB = 8;  // bands per warp pass

struct ZonalWs {
    int32_t *xmin, *xmax, *ymin, *ymax, *count;
    int64_t bytes;
};

static ZonalWs zonal_ws_layout(void *base, int64_t max_label)
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

__global__ void zonal_init_kernel(ZonalWs w, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    w.xmin[i] = 0x7fffffff;
    w.ymin[i] = 0x7fffffff;
    w.xmax[i] = -1;
    w.ymax[i] = -1;
    w.count[i] = 0;
}

__global__ void __launch_bounds__(256)
zonal_bbox_kernel(const int32_t *__restrict__ labels, ZonalWs w, int64_t N, int W, int64_t max_label)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int32_t l = -1;
    int x = 0, y = 0;
    if (i < N) {
        l = labels[i];
        if (l < 0 || (int64_t)l > max_label) l = -1;
        y = (int)(i / W);
        x = (int)(i - (int64_t)y * W);
    }
    const int32_t prev = __shfl_up_sync(0xffffffffu, l, 1);
    const bool is_head = (lane == 0) || (prev != l) || (x == 0);
    const unsigned heads = __ballot_sync(0xffffffffu, is_head);
    if (is_head && l >= 0) {
        const unsigned later = (lane == 31) ? 0u : (heads & ~((2u << lane) - 1u));
        const int end = later ? (__ffs(later) - 1) : 32;
        const int len = end - lane;
        atomicMin(w.xmin + l, x);
        atomicMax(w.xmax + l, x + len - 1);
        atomicMin(w.ymin + l, y);
        atomicMax(w.ymax + l, y);
        atomicAdd(w.count + l, len);
    }
}

struct ZBands {
    int32_t band[OBIA_B200_MAX_BANDS];
};

__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// stats layout per (label, band): count, mean, variance, min, max, skewness, kurtosis, sum
__global__ void __launch_bounds__(256)
zonal_gather_kernel(const int32_t *__restrict__ labels, const float *__restrict__ raw, ZonalWs w, int W,
                    int C, ZBands zb, int Cz, int64_t max_label, double resolution,
                    double *__restrict__ stats)
{
    const int lane = threadIdx.x & 31;
    const int64_t L = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (L > max_label) return;
    const int b0 = blockIdx.y * kZB;
    const int nb = min(kZB, Cz - b0);
    double *out = stats + (L * Cz + b0) * 8;
    const int cnt_total = w.count[L];
    const double NAND = __longlong_as_double(0x7ff8000000000000LL);
    if (cnt_total == 0) {
        for (int i = lane; i < nb * 8; i += 32) out[i] = ((i & 7) == 0 || (i & 7) == 7) ? 0.0 : NAND;
        return;
    }
    const int x0 = w.xmin[L], x1 = w.xmax[L], y0 = w.ymin[L], y1 = w.ymax[L];
    int bidx[kZB];
#pragma unroll
    for (int b = 0; b < kZB; ++b) bidx[b] = zb.band[min(b0 + b, Cz - 1)];

    // pivot = the segment's first pixel in its first row (row y0 holds one by construction)
    float pivot[kZB];
    {
        int xr = -1;
        for (int xs = x0; xs <= x1 && xr < 0; xs += 32) {
            const int x = xs + lane;
            const bool hit = (x <= x1) && (labels[(int64_t)y0 * W + x] == (int32_t)L);
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) xr = xs + __ffs(m) - 1;
        }
        const float *p = raw + ((int64_t)y0 * W + xr) * C;
#pragma unroll
        for (int b = 0; b < kZB; ++b) pivot[b] = p[bidx[b]];
    }

    const float INF = __int_as_float(0x7f800000);
    float s1[kZB], s2[kZB], s3[kZB], s4[kZB], mn[kZB], mx[kZB];
    double d1[kZB], d2[kZB], d3[kZB], d4[kZB];
#pragma unroll
    for (int b = 0; b < kZB; ++b) {
        s1[b] = s2[b] = s3[b] = s4[b] = 0.0f;
        d1[b] = d2[b] = d3[b] = d4[b] = 0.0;
        mn[b] = INF;
        mx[b] = -INF;
    }
    int pending = 0;
    for (int y = y0; y <= y1; ++y) {
        const int64_t row = (int64_t)y * W;
        for (int xs = x0; xs <= x1; xs += 32) {
            const int x = xs + lane;
            if (x <= x1 && labels[row + x] == (int32_t)L) {
                const float *p = raw + (row + x) * C;
#pragma unroll
                for (int b = 0; b < kZB; ++b) {
                    const float v = p[bidx[b]];
                    const float d = v - pivot[b];
                    const float dd = d * d;
                    s1[b] += d;
                    s2[b] += dd;
                    s3[b] = fmaf(dd, d, s3[b]);
                    s4[b] = fmaf(dd, dd, s4[b]);
                    mn[b] = fminf(mn[b], v);
                    mx[b] = fmaxf(mx[b], v);
                }
                if (++pending == 16) {
                    pending = 0;
#pragma unroll
                    for (int b = 0; b < kZB; ++b) {
                        d1[b] += (double)s1[b]; d2[b] += (double)s2[b];
                        d3[b] += (double)s3[b]; d4[b] += (double)s4[b];
                        s1[b] = s2[b] = s3[b] = s4[b] = 0.0f;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int b = 0; b < kZB; ++b) {
        d1[b] = warp_sum_d(d1[b] + (double)s1[b]);
        d2[b] = warp_sum_d(d2[b] + (double)s2[b]);
        d3[b] = warp_sum_d(d3[b] + (double)s3[b]);
        d4[b] = warp_sum_d(d4[b] + (double)s4[b]);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            mn[b] = fminf(mn[b], __shfl_xor_sync(0xffffffffu, mn[b], o));
            mx[b] = fmaxf(mx[b], __shfl_xor_sync(0xffffffffu, mx[b], o));
        }
    }
    // lane b finishes band b
#pragma unroll
    for (int b = 0; b < kZB; ++b) {
        if (lane == b && b < nb) {
            const double n = (double)cnt_total;
            const double md = d1[b] / n;
            const double e2 = d2[b] / n, e3 = d3[b] / n, e4 = d4[b] / n;
            const double mean = (double)pivot[b] + md;
            double m2 = e2 - md * md;
            if (m2 < 0.0) m2 = 0.0;
            const double m3 = e3 - 3.0 * md * e2 + 2.0 * md * md * md;
            const double m4 = e4 - 4.0 * md * e3 + 6.0 * md * md * e2 - 3.0 * md * md * md * md;
            // scipy.stats.skew/kurtosis: NaN when the data are (nearly) constant
            const double thr = resolution * mean;
            const bool degenerate = m2 <= thr * thr;
            double *o = out + b * 8;
            o[0] = n;
            o[1] = mean;
            o[2] = m2;
            o[3] = (double)mn[b];
            o[4] = (double)mx[b];
            o[5] = degenerate ? NAND : m3 / (m2 * sqrt(m2));
            o[6] = degenerate ? NAND : m4 / (m2 * m2) - 3.0;
            o[7] = mean * n;
        }
    }
}

}  // namespace obia

using namespace obia;

extern "C" int64_t obia_b200_zonal_workspace_bytes(int64_t max_label, int32_t Cz)
{
    if (max_label < 0 || Cz <= 0) return -1;
    return zonal_ws_layout(nullptr, max_label).bytes;
}

extern "C" int obia_b200_zonal_stats(const int32_t *labels, const float *raw, int64_t H, int64_t W,
                                     int32_t C, const int32_t *bands_host, int32_t Cz, int64_t max_label,
                                     double resolution, double *stats, void *workspace, void *stream)
{
    if (!labels || !raw || !bands_host || !stats || !workspace || H <= 0 || W <= 0 || C <= 0 || Cz <= 0 ||
        max_label < 0)
        return set_err(OBIA_B200_ERR_ARG, "zonal_stats: bad argument");
    if (Cz > OBIA_B200_MAX_BANDS)
        return set_err(OBIA_B200_ERR_UNSUPPORTED, "zonal_stats: at most %d statistics bands per call",
                       OBIA_B200_MAX_BANDS);
    if (H * W >= 0x7fffffffLL || max_label >= 0x7fffffffLL)
        return set_err(OBIA_B200_ERR_UNSUPPORTED, "zonal_stats: H*W exceeds int32");
    ZBands zb;
    memset(&zb, 0, sizeof(zb));
    for (int b = 0; b < Cz; ++b) {
        if (bands_host[b] < 0 || bands_host[b] >= C)
            return set_err(OBIA_B200_ERR_ARG, "zonal_stats: band %d out of range", bands_host[b]);
        zb.band[b] = bands_host[b];
    }
    cudaStream_t st = (cudaStream_t)stream;
    ZonalWs w = zonal_ws_layout(workspace, max_label);
    const int64_t n = max_label + 1;
    const int64_t N = H * W;
    zonal_init_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(w, n);
    OBIA_LAUNCH_CHECK();
    zonal_bbox_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(labels, w, N, (int)W, max_label);
    OBIA_LAUNCH_CHECK();
    dim3 grid((unsigned)ceil_div(n, 8), (unsigned)ceil_div(Cz, kZB));
    zonal_gather_kernel<<<grid, 256, 0, st>>>(labels, raw, w, (int)W, C, zb, Cz, max_label, resolution,
                                              stats);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

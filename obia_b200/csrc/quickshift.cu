// method="quickshift": replaces `skimage.segmentation.quickshift` as the reference calls it
// (obia/segmentation/segment_boundaries.py:48-49; skimage/segmentation/_quickshift_cython.pyx,
// restated in oracle/quickshift_oracle.py).  Three data-parallel passes, one thread per pixel:
//   density  sum over the clipped (2w+1)^2 window, row-major like the reference's loops, of
//            exp(-dist / (2 kernel_size^2)): float32 `dist` (separately rounded multiply / add),
//            double `exp`, the running sum rounded to float32 after every term; then the caller's
//            tie-breaking noise (numpy default_rng(rng).normal(scale=1e-5), drawn on the host) is added;
//   parent   the window pixel of strictly higher density at the smallest dist (first in row-major order
//            among equal distances); sqrt(dist) > max_dist or no such pixel -> the pixel is a root;
//   labels   every pixel follows its parents to the root; roots are numbered in raster order
//            (np.unique(..., return_inverse=True)) with a bitmap + prefix population count.
// Features are the planar [Cf][H][pitch] float32 array of obia_b200_slic_features (per-band normalise,
// select, optional CIELAB, multiply by `ratio`).  The window reads hit L1 / L2: neighbouring threads
// share all but one column of their windows.
#include "common.cuh"

namespace obia {

constexpr int kQsChunk = 1024;   // bitmap words per numbering block

struct QsWs {
    float *dens;
    int32_t *parent, *fin, *chunksum, *ctr;
    uint32_t *bits;
    int64_t nwords, nchunks, bytes;
};

static QsWs qs_ws_layout(void *base, int64_t N)
{
    QsWs w;
    char *p = (char *)base;
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        char *r = p + off;
        off += round_up(bytes, 256);
        return r;
    };
    w.dens = (float *)take(N * 4);
    w.parent = (int32_t *)take(N * 4);
    w.fin = (int32_t *)take(N * 4);
    w.nwords = ceil_div(N, 32);
    w.nchunks = ceil_div(w.nwords, kQsChunk);
    w.bits = (uint32_t *)take(w.nchunks * kQsChunk * 4);
    w.chunksum = (int32_t *)take((w.nchunks + 1) * 4);
    w.ctr = (int32_t *)take(64);
    w.bytes = off;
    return w;
}

template <int CF>
__device__ __forceinline__ float qs_dist(const float *__restrict__ feat, const float (&cur)[CF], int Cf, int64_t plane,
                                         int64_t q, int dr, int dc)
{
    float dist = 0.0f;
#pragma unroll
    for (int ch = 0; ch < CF; ++ch) {
        if (ch < Cf) {
            const float t = __fsub_rn(cur[ch], __ldg(feat + ch * plane + q));
            dist = __fadd_rn(dist, __fmul_rn(t, t));
        }
    }
    const float tr = (float)(-dr), tc = (float)(-dc);
    dist = __fadd_rn(dist, __fmul_rn(tr, tr));
    dist = __fadd_rn(dist, __fmul_rn(tc, tc));
    return dist;
}

template <int CF>
__global__ void __launch_bounds__(256)
qs_density_kernel(const float *__restrict__ feat, const double *__restrict__ noise, float *__restrict__ dens, int H,
                  int W, int64_t pitch, int Cf, int kw, float inv)
{
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), r = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (r >= H || c >= W) return;
    const int64_t plane = (int64_t)H * pitch;
    float cur[CF];
#pragma unroll
    for (int ch = 0; ch < CF; ++ch) cur[ch] = ch < Cf ? feat[ch * plane + (int64_t)r * pitch + c] : 0.0f;
    const int r0 = max(r - kw, 0), r1 = min(r + kw + 1, H), c0 = max(c - kw, 0), c1 = min(c + kw + 1, W);
    float d = 0.0f;
    for (int rr = r0; rr < r1; ++rr)
        for (int cc = c0; cc < c1; ++cc) {
            const float dist = qs_dist<CF>(feat, cur, Cf, plane, (int64_t)rr * pitch + cc, rr - r, cc - c);
            d = (float)((double)d + exp((double)__fmul_rn(dist, inv)));
        }
    dens[(int64_t)r * W + c] = (float)((double)d + noise[(int64_t)r * W + c]);
}

template <int CF>
__global__ void __launch_bounds__(256)
qs_parent_kernel(const float *__restrict__ feat, const float *__restrict__ dens, int32_t *__restrict__ parent, int H,
                 int W, int64_t pitch, int Cf, int kw, float max_dist)
{
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), r = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (r >= H || c >= W) return;
    const int64_t plane = (int64_t)H * pitch;
    float cur[CF];
#pragma unroll
    for (int ch = 0; ch < CF; ++ch) cur[ch] = ch < Cf ? feat[ch * plane + (int64_t)r * pitch + c] : 0.0f;
    const int r0 = max(r - kw, 0), r1 = min(r + kw + 1, H), c0 = max(c - kw, 0), c1 = min(c + kw + 1, W);
    const float mine = dens[(int64_t)r * W + c];
    float closest = __int_as_float(0x7f800000);
    int32_t best = r * W + c;
    for (int rr = r0; rr < r1; ++rr)
        for (int cc = c0; cc < c1; ++cc) {
            if (!(__ldg(dens + (int64_t)rr * W + cc) > mine)) continue;
            const float dist = qs_dist<CF>(feat, cur, Cf, plane, (int64_t)rr * pitch + cc, rr - r, cc - c);
            if (dist < closest) {
                closest = dist;
                best = rr * W + cc;
            }
        }
    // dist_parent = sqrt(closest) stored as float32; parents further than max_dist are cut
    if ((float)sqrt((double)closest) > max_dist) best = r * W + c;
    parent[(int64_t)r * W + c] = best;
}

// follow the parents to the root (densities strictly increase along a chain: no cycles); roots set their bit
__global__ void __launch_bounds__(256)
qs_root_kernel(const int32_t *__restrict__ parent, int32_t *__restrict__ root, uint32_t *bits, int64_t N)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int32_t x = (int32_t)i, p = __ldg(parent + x);
    while (p != x) {
        x = p;
        p = __ldg(parent + x);
    }
    root[i] = x;
    if (x == (int32_t)i) atomicOr(bits + (i >> 5), 1u << (i & 31));
}

__global__ void __launch_bounds__(256)
qs_bits_count_kernel(const uint32_t *__restrict__ bits, int32_t *chunksum)
{
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int c = 0;
    for (int j = threadIdx.x; j < kQsChunk; j += 256) c += __popc(bits[(int64_t)blockIdx.x * kQsChunk + j]);
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) chunksum[blockIdx.x] = s_cnt;
}

__global__ void qs_scan_kernel(int32_t *chunksum, int64_t n, int32_t *ctr)
{
    // one thread: n = ceil(N / 32768) entries (3052 for a 10^8-pixel raster)
    int run = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int v = chunksum[i];
        chunksum[i] = run;
        run += v;
    }
    ctr[0] = run;
}

__global__ void __launch_bounds__(256)
qs_bits_number_kernel(const uint32_t *__restrict__ bits, const int32_t *__restrict__ chunksum, int32_t *fin)
{
    __shared__ int s_warp[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t w0 = (int64_t)blockIdx.x * kQsChunk;
    int running = chunksum[blockIdx.x];
    for (int j0 = 0; j0 < kQsChunk; j0 += 256) {
        const uint32_t v = bits[w0 + j0 + threadIdx.x];
        const int c = __popc(v);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int before = running + incl - c, tot = 0;
        for (int w = 0; w < 8; ++w) {
            if (w < warp) before += s_warp[w];
            tot += s_warp[w];
        }
        const int64_t p0 = (w0 + j0 + threadIdx.x) * 32;
        int k = 0;
        for (uint32_t m = v; m; m &= m - 1) fin[p0 + __ffs(m) - 1] = before + k++;
        running += tot;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
qs_label_kernel(const int32_t *__restrict__ root, const int32_t *__restrict__ fin, int32_t *__restrict__ labels, int64_t N)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) labels[i] = fin[root[i]];
}

template <int CF>
static int qs_run(const float *features, const double *noise, int32_t *labels, const QsWs &w, int64_t H, int64_t W,
                  int64_t pitch, int Cf, int kw, float inv, float max_dist, cudaStream_t st)
{
    dim3 grid((unsigned)ceil_div(W, 32), (unsigned)ceil_div(H, 8));
    qs_density_kernel<CF><<<grid, 256, 0, st>>>(features, noise, w.dens, (int)H, (int)W, pitch, Cf, kw, inv);
    OBIA_LAUNCH_CHECK();
    qs_parent_kernel<CF><<<grid, 256, 0, st>>>(features, w.dens, w.parent, (int)H, (int)W, pitch, Cf, kw, max_dist);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

}  // namespace obia

using namespace obia;

extern "C" int64_t obia_b200_quickshift_workspace_bytes(int64_t H, int64_t W)
{
    if (H <= 0 || W <= 0) return -1;
    return qs_ws_layout(nullptr, H * W).bytes;
}

extern "C" int obia_b200_quickshift(const float *features, const double *noise, int32_t *labels, void *workspace,
                                    int64_t H, int64_t W, int64_t pitch, int32_t Cf, float kernel_size, float max_dist,
                                    int64_t *n_labels_host, void *stream)
{
    if (!features || !noise || !labels || !workspace || H <= 0 || W <= 0 || pitch < W || Cf <= 0)
        return set_err(OBIA_B200_ERR_ARG, "quickshift: bad argument");
    if (!(kernel_size >= 1.0f)) return set_err(OBIA_B200_ERR_ARG, "`kernel_size` should be >= 1.");
    if (Cf > OBIA_B200_MAX_BANDS) return set_err(OBIA_B200_ERR_UNSUPPORTED, "quickshift: more than %d channels", OBIA_B200_MAX_BANDS);
    const int64_t N = H * W;
    if (N >= 0x7fffffffLL) return set_err(OBIA_B200_ERR_UNSUPPORTED, "quickshift: H*W exceeds int32");
    cudaStream_t st = (cudaStream_t)stream;
    QsWs w = qs_ws_layout(workspace, N);
    const int kw = (int)ceil(3.0 * (double)kernel_size);
    const float inv = -0.5f / (kernel_size * kernel_size);
    OBIA_CUDA_CHECK(cudaMemsetAsync(w.bits, 0, (size_t)w.nchunks * kQsChunk * 4, st));
    int rc;
    if (Cf <= 4) rc = qs_run<4>(features, noise, labels, w, H, W, pitch, Cf, kw, inv, max_dist, st);
    else if (Cf <= 8) rc = qs_run<8>(features, noise, labels, w, H, W, pitch, Cf, kw, inv, max_dist, st);
    else if (Cf <= 16) rc = qs_run<16>(features, noise, labels, w, H, W, pitch, Cf, kw, inv, max_dist, st);
    else rc = qs_run<OBIA_B200_MAX_BANDS>(features, noise, labels, w, H, W, pitch, Cf, kw, inv, max_dist, st);
    if (rc) return rc;
    int32_t *root = reinterpret_cast<int32_t *>(w.dens);   // densities are no longer needed
    qs_root_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(w.parent, root, w.bits, N);
    OBIA_LAUNCH_CHECK();
    qs_bits_count_kernel<<<(unsigned)w.nchunks, 256, 0, st>>>(w.bits, w.chunksum);
    OBIA_LAUNCH_CHECK();
    qs_scan_kernel<<<1, 1, 0, st>>>(w.chunksum, w.nchunks, w.ctr);
    OBIA_LAUNCH_CHECK();
    qs_bits_number_kernel<<<(unsigned)w.nchunks, 256, 0, st>>>(w.bits, w.chunksum, w.fin);
    OBIA_LAUNCH_CHECK();
    qs_label_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(root, w.fin, labels, N);
    OBIA_LAUNCH_CHECK();
    if (n_labels_host) {
        int32_t n = 0;
        OBIA_CUDA_CHECK(cudaMemcpyAsync(&n, w.ctr, 4, cudaMemcpyDeviceToHost, st));
        OBIA_CUDA_CHECK(cudaStreamSynchronize(st));
        *n_labels_host = n;
    }
    return OBIA_B200_OK;
}

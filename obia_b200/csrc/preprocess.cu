// K1: per-band min/max, in-place normalise, fused feature preparation
// (band select + normalise + global rescale + RGB->CIELAB + 1/compactness),
// separable Gaussian.  HBM-bound streaming kernels: 128-bit coalesced loads of
// the pixel-interleaved raster, staged through shared memory where the access
// has to be re-indexed per band.
#include <math.h>

#include "batch.cuh"
#include "common.cuh"

namespace obia {

// ---------------------------------------------------------------- min/max --
__global__ void minmax_init_kernel(uint32_t *keys, int32_t *nonfinite, int C)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) {
        keys[c * 4 + 0] = 0xffffffffu;
        keys[c * 4 + 1] = 0u;
        keys[c * 4 + 2] = 0xffffffffu;
        keys[c * 4 + 3] = 0u;
        nonfinite[c] = 0;
    }
}

__device__ __forceinline__ int classify_nonfinite(float v)
{
    uint32_t b = __float_as_uint(v);
    if ((b & 0x7f800000u) != 0x7f800000u) return 0;
    return (b & 0x007fffffu) ? 1 : 2;  // NaN : inf
}

// Each thread reads float4 vectors v = tid, tid+NT, ... with 4*NT % C == 0, so
// lane slot j of a thread always carries the same band (4*tid + j) % C and the
// per-band running min/max live in registers.
template <bool MASK>
__global__ void __launch_bounds__(256)
minmax_kernel(const float *__restrict__ raw, int64_t n_elems, int C,
              const uint8_t *__restrict__ mask, uint32_t *keys, int32_t *nonfinite)
{
    extern __shared__ uint32_t s_keys[];  // [C*4] keys, then [C] flags
    uint32_t *s_flag = s_keys + (size_t)C * 4;
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        s_keys[i * 4 + 0] = 0xffffffffu;
        s_keys[i * 4 + 1] = 0u;
        s_keys[i * 4 + 2] = 0xffffffffu;
        s_keys[i * 4 + 3] = 0u;
        s_flag[i] = 0u;
    }
    __syncthreads();

    const int64_t NT = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nvec = n_elems >> 2;
    const float INF = __int_as_float(0x7f800000);
    float mn[4], mx[4], mmn[4], mmx[4];
    int fl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        mn[j] = INF; mx[j] = -INF; mmn[j] = INF; mmx[j] = -INF; fl[j] = 0;
    }
    const float4 *raw4 = reinterpret_cast<const float4 *>(raw);
    for (int64_t v = tid; v < nvec; v += NT) {
        float4 x = ldg_stream_f4(raw4 + v);
        float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            fl[j] |= classify_nonfinite(xs[j]);
            mn[j] = fminf(mn[j], xs[j]);
            mx[j] = fmaxf(mx[j], xs[j]);
            if (MASK) {
                int64_t pix = (v * 4 + j) / C;
                if (mask[pix]) {
                    mmn[j] = fminf(mmn[j], xs[j]);
                    mmx[j] = fmaxf(mmx[j], xs[j]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int band = (int)((tid * 4 + j) % C);
        if (mn[j] <= mx[j]) {  // saw at least one non-NaN value
            atomicMin(&s_keys[band * 4 + 0], float_to_key(mn[j]));
            atomicMax(&s_keys[band * 4 + 1], float_to_key(mx[j]));
        }
        if (MASK && mmn[j] <= mmx[j]) {
            atomicMin(&s_keys[band * 4 + 2], float_to_key(mmn[j]));
            atomicMax(&s_keys[band * 4 + 3], float_to_key(mmx[j]));
        }
        if (fl[j]) atomicOr(&s_flag[band], (uint32_t)fl[j]);
    }
    // scalar tail (n_elems % 4 elements)
    if (tid < (n_elems & 3)) {
        int64_t e = nvec * 4 + tid;
        float x = raw[e];
        int band = (int)(e % C);
        int f = classify_nonfinite(x);
        if (f) atomicOr(&s_flag[band], (uint32_t)f);
        if (f != 1) {
            atomicMin(&s_keys[band * 4 + 0], float_to_key(x));
            atomicMax(&s_keys[band * 4 + 1], float_to_key(x));
            if (MASK && mask[e / C]) {
                atomicMin(&s_keys[band * 4 + 2], float_to_key(x));
                atomicMax(&s_keys[band * 4 + 3], float_to_key(x));
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        atomicMin(&keys[i * 4 + 0], s_keys[i * 4 + 0]);
        atomicMax(&keys[i * 4 + 1], s_keys[i * 4 + 1]);
        if (MASK) {
            atomicMin(&keys[i * 4 + 2], s_keys[i * 4 + 2]);
            atomicMax(&keys[i * 4 + 3], s_keys[i * 4 + 3]);
        }
        if (s_flag[i]) atomicOr(&nonfinite[i], (int)s_flag[i]);
    }
}

__global__ void minmax_finish_kernel(uint32_t *keys, int C, int has_mask)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float *out = reinterpret_cast<float *>(keys);
    const float NANF = __int_as_float(0x7fc00000);
    uint32_t k0 = keys[c * 4 + 0], k1 = keys[c * 4 + 1];
    uint32_t k2 = keys[c * 4 + 2], k3 = keys[c * 4 + 3];
    float mn = (k0 == 0xffffffffu && k1 == 0u) ? NANF : key_to_float(k0);
    float mx = (k0 == 0xffffffffu && k1 == 0u) ? NANF : key_to_float(k1);
    float mmn, mmx;
    if (!has_mask) {
        mmn = mn; mmx = mx;
    } else if (k2 == 0xffffffffu && k3 == 0u) {
        mmn = NANF; mmx = NANF;
    } else {
        mmn = key_to_float(k2); mmx = key_to_float(k3);
    }
    out[c * 4 + 0] = mn; out[c * 4 + 1] = mx; out[c * 4 + 2] = mmn; out[c * 4 + 3] = mmx;
}

static int gcd_i(int a, int b) { return b ? gcd_i(b, a % b) : a; }

// ----------------------------------------------------------- normalise --
// src == dst: in place.  dst may be device-accessible (pinned, mapped) HOST memory: the stores are
// full 16-byte vectors, coalesced per warp, so they stream over PCIe without the copy engine.
__global__ void __launch_bounds__(256)
normalize_kernel(const float *src, float *dst, int64_t n_elems, int C, const float *__restrict__ minmax)
{
    const int64_t NT = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nvec = n_elems >> 2;
    float mn[4], d[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int band = (int)((tid * 4 + j) % C);
        mn[j] = minmax[band * 4 + 0];
        d[j] = __fsub_rn(minmax[band * 4 + 1], mn[j]);
    }
    const float4 *src4 = reinterpret_cast<const float4 *>(src);
    float4 *dst4 = reinterpret_cast<float4 *>(dst);
    for (int64_t v = tid; v < nvec; v += NT) {
        float4 x = src4[v];
        x.x = __fdiv_rn(__fsub_rn(x.x, mn[0]), d[0]);
        x.y = __fdiv_rn(__fsub_rn(x.y, mn[1]), d[1]);
        x.z = __fdiv_rn(__fsub_rn(x.z, mn[2]), d[2]);
        x.w = __fdiv_rn(__fsub_rn(x.w, mn[3]), d[3]);
        dst4[v] = x;
    }
    if (tid < (n_elems & 3)) {
        int64_t e = nvec * 4 + tid;
        int band = (int)(e % C);
        float m = minmax[band * 4 + 0];
        dst[e] = __fdiv_rn(__fsub_rn(src[e], m), __fsub_rn(minmax[band * 4 + 1], m));
    }
}

// ---------------------------------------------------------------- features --
struct BandTable {
    int32_t band[OBIA_B200_MAX_BANDS];
    float mn[OBIA_B200_MAX_BANDS];
    float d[OBIA_B200_MAX_BANDS];  // max - min (float32 subtraction, like numpy)
};

// skimage.color.rgb2lab in float32
__device__ __forceinline__ float srgb_lin(float v)
{
    return (v > 0.04045f) ? powf(__fdiv_rn(__fadd_rn(v, 0.055f), 1.055f), 2.4f)
                          : __fdiv_rn(v, 12.92f);
}
__device__ __forceinline__ float lab_f(float t)
{
    return (t > 0.008856f) ? cbrtf(t)
                           : __fadd_rn(__fmul_rn(7.787f, t), (float)(16.0 / 116.0));
}
__device__ __forceinline__ void rgb2lab_f32(float r, float g, float b, float &L, float &A, float &B)
{
    r = srgb_lin(r); g = srgb_lin(g); b = srgb_lin(b);
    float x = __fadd_rn(__fadd_rn(__fmul_rn(r, 0.412453f), __fmul_rn(g, 0.357580f)), __fmul_rn(b, 0.180423f));
    float y = __fadd_rn(__fadd_rn(__fmul_rn(r, 0.212671f), __fmul_rn(g, 0.715160f)), __fmul_rn(b, 0.072169f));
    float z = __fadd_rn(__fadd_rn(__fmul_rn(r, 0.019334f), __fmul_rn(g, 0.119193f)), __fmul_rn(b, 0.950227f));
    x = __fdiv_rn(x, 0.95047f);
    y = __fdiv_rn(y, 1.0f);
    z = __fdiv_rn(z, 1.08883f);
    float fx = lab_f(x), fy = lab_f(y), fz = lab_f(z);
    L = __fsub_rn(__fmul_rn(116.0f, fy), 16.0f);
    A = __fmul_rn(500.0f, __fsub_rn(fx, fy));
    B = __fmul_rn(200.0f, __fsub_rn(fy, fz));
}

constexpr int kFeatPix = 128;  // pixels per CTA tile (flat, row-crossing)

__device__ __forceinline__ int swz(int i) { return i + (i >> 5); }

// One CTA stages kFeatPix consecutive pixels (kFeatPix*C contiguous floats)
// into shared memory with 128-bit coalesced loads, then each thread converts
// one pixel and writes Cf coalesced planar stores.
__global__ void __launch_bounds__(kFeatPix)
features_kernel(const float *__restrict__ raw, int64_t n_pixels, int64_t W, int C, int Cs,
                BandTable tab, float imin, float idiff, int rescale, int to_lab, float ratio,
                float *__restrict__ feat, int64_t pitch, int64_t plane)
{
    extern __shared__ float s_raw[];
    const int64_t p0 = (int64_t)blockIdx.x * kFeatPix;
    const int64_t np = min((int64_t)kFeatPix, n_pixels - p0);
    const int64_t e0 = p0 * C;
    const int nelem = (int)(np * C);
    // e0 = 128*blk*C floats -> 16-byte aligned
    const float4 *src4 = reinterpret_cast<const float4 *>(raw + e0);
    const int nvec = nelem >> 2;
    for (int v = threadIdx.x; v < nvec; v += kFeatPix) {
        float4 x = ldg_stream_f4(src4 + v);
        s_raw[swz(4 * v + 0)] = x.x;
        s_raw[swz(4 * v + 1)] = x.y;
        s_raw[swz(4 * v + 2)] = x.z;
        s_raw[swz(4 * v + 3)] = x.w;
    }
    for (int e = nvec * 4 + threadIdx.x; e < nelem; e += kFeatPix) s_raw[swz(e)] = raw[e0 + e];
    __syncthreads();

    const int t = threadIdx.x;
    if (t >= np) return;
    const int64_t p = p0 + t;
    const int64_t y = p / W;
    const int64_t x = p - y * W;
    float *dst = feat + y * pitch + x;
    if (to_lab) {
        float v[3];
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            float r = s_raw[swz(t * C + tab.band[s])];
            r = __fdiv_rn(__fsub_rn(r, tab.mn[s]), tab.d[s]);
            r = __fsub_rn(r, imin);
            if (rescale) r = __fdiv_rn(r, idiff);
            v[s] = r;
        }
        float L, A, B;
        rgb2lab_f32(v[0], v[1], v[2], L, A, B);
        dst[0] = __fmul_rn(L, ratio);
        dst[plane] = __fmul_rn(A, ratio);
        dst[2 * plane] = __fmul_rn(B, ratio);
    } else {
        for (int s = 0; s < Cs; ++s) {
            float r = s_raw[swz(t * C + tab.band[s])];
            r = __fdiv_rn(__fsub_rn(r, tab.mn[s]), tab.d[s]);
            r = __fsub_rn(r, imin);
            if (rescale) r = __fdiv_rn(r, idiff);
            dst[(int64_t)s * plane] = __fmul_rn(r, ratio);
        }
    }
}

// ---------------------------------------------------------------- gaussian --
// scipy.ndimage.correlate1d for a symmetric kernel: centre tap first, then
// (left + right) pairs from the outermost tap inwards, all in float64; the
// result is stored as float32 between the two passes.  mode='reflect'
// (half-sample symmetric).
__device__ __forceinline__ int64_t reflect_idx(int64_t i, int64_t n)
{
    if (n == 1) return 0;
    const int64_t period = 2 * n;
    i %= period;
    if (i < 0) i += period;
    return (i < n) ? i : period - 1 - i;
}

constexpr int kMaxGaussRadius = 255;
__constant__ double c_gauss_w[2][kMaxGaussRadius + 1];  // [axis][0..radius], w[0] = centre

template <int AXIS>  // 0: along y, 1: along x
__global__ void __launch_bounds__(256)
gaussian_pass_kernel(const float *__restrict__ in, float *__restrict__ out, int64_t H, int64_t W,
                     int64_t pitch, int radius, float ratio, int apply_ratio)
{
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t y = blockIdx.y;
    const int64_t plane = (int64_t)blockIdx.z * H * pitch;
    if (x >= W) return;
    const float *src = in + plane;
    const int64_t n = AXIS == 0 ? H : W;
    const int64_t pos = AXIS == 0 ? y : x;
    auto at = [&](int64_t i) -> double {
        int64_t r = reflect_idx(i, n);
        return AXIS == 0 ? (double)src[r * pitch + x] : (double)src[y * pitch + r];
    };
    double acc = __dmul_rn(at(pos), c_gauss_w[AXIS][0]);
    for (int k = radius; k >= 1; --k) {
        double pair = __dadd_rn(at(pos - k), at(pos + k));
        acc = __dadd_rn(acc, __dmul_rn(pair, c_gauss_w[AXIS][k]));
    }
    float r = (float)acc;
    if (apply_ratio) r = __fmul_rn(r, ratio);
    out[plane + y * pitch + x] = r;
}

// Fused, shared-memory tiled version of the two passes (north_star subsystem 1): a CTA stages a
// (TY + 2 ry) x (TX + 2 rx) window of one feature plane with coalesced loads (half-sample reflect at
// the raster edges), runs the column pass into a second shared buffer -- rounded to float32 exactly
// where scipy stores the intermediate array -- then the row pass, and writes the TY x TX tile once.
// Arithmetic and order are those of gaussian_pass_kernel (float64, centre tap, pairs outermost first).
// The taps travel as kernel parameters: no constant-memory upload, no stream synchronisation.
constexpr int kGaussTX = 64, kGaussTY = 32, kGaussParamRadius = 63;
struct GaussTaps {
    double wy[kGaussParamRadius + 1], wx[kGaussParamRadius + 1];   // [0] = centre tap
};

__global__ void __launch_bounds__(256)
gaussian_tiled_kernel(const float *__restrict__ in, float *__restrict__ out, int H, int W, int64_t pitch, int Cf,
                      int ry, int rx, GaussTaps taps, float ratio)
{
    extern __shared__ float s_g[];
    const int IW = kGaussTX + 2 * rx, IH = kGaussTY + 2 * ry;
    float *s_in = s_g;                 // [IH][IW]
    float *s_mid = s_g + IH * IW;      // [TY][IW] column pass, float32 like scipy's intermediate
    const int x0 = blockIdx.x * kGaussTX, y0 = blockIdx.y * kGaussTY;
    for (int c = 0; c < Cf; ++c) {
        const float *src = in + (int64_t)c * H * pitch;
        for (int i = threadIdx.x; i < IH * IW; i += 256) {
            const int ly = i / IW, lx = i - ly * IW;
            const int64_t gy = reflect_idx(y0 + ly - ry, H), gx = reflect_idx(x0 + lx - rx, W);
            s_in[i] = src[gy * pitch + gx];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < kGaussTY * IW; i += 256) {
            const int ty = i / IW, lx = i - ty * IW;
            const float *col = s_in + (ty + ry) * IW + lx;
            double acc = __dmul_rn((double)col[0], taps.wy[0]);
            for (int k = ry; k >= 1; --k) {
                const double pair = __dadd_rn((double)col[-k * IW], (double)col[k * IW]);
                acc = __dadd_rn(acc, __dmul_rn(pair, taps.wy[k]));
            }
            s_mid[i] = (float)acc;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < kGaussTY * kGaussTX; i += 256) {
            const int ty = i / kGaussTX, tx = i - ty * kGaussTX;
            const int y = y0 + ty, x = x0 + tx;
            if (y >= H || x >= W) continue;
            const float *row = s_mid + ty * IW + tx + rx;
            double acc = __dmul_rn((double)row[0], taps.wx[0]);
            for (int k = rx; k >= 1; --k) {
                const double pair = __dadd_rn((double)row[-k], (double)row[k]);
                acc = __dadd_rn(acc, __dmul_rn(pair, taps.wx[k]));
            }
            out[(int64_t)c * H * pitch + (int64_t)y * pitch + x] = __fmul_rn((float)acc, ratio);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- batched windows (tiled driver) --
// Every window of a slab at once (batch.cuh): blockIdx.y = window.  `raw` is the source raster (H, Wl, C)
// interleaved; a window's pixel (y, x) is raw[((y0 + y) * Wl + x0 + x) * C + band].
__global__ void win_stats_init_kernel(uint32_t *keys, int32_t *flags, int32_t *counts, int64_t B, int C)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * C) {
        keys[i * 4 + 0] = 0xffffffffu;
        keys[i * 4 + 1] = 0u;
        keys[i * 4 + 2] = 0xffffffffu;
        keys[i * 4 + 3] = 0u;
        flags[i] = 0;
    }
    if (i < B) counts[i] = 0;
}

constexpr int kWinStatRows = 16;

// per window and band: min / max over the window, min / max over its mask pixels, NaN / inf flags (the
// quantities obia_b200_band_minmax gives for one raster), plus the number of mask pixels.  The first
// (256 / C) * C threads work, so a thread always sees the same band.
__global__ void __launch_bounds__(256)
win_stats_kernel(const float *__restrict__ raw, int64_t Wl, int C, const WinDesc *__restrict__ batch,
                 const uint8_t *__restrict__ mask_slab, int slab_w, uint32_t *keys, int32_t *flags, int32_t *counts)
{
    const WinDesc &d = batch[blockIdx.y];
    const int r0 = blockIdx.x * kWinStatRows;
    if (r0 >= d.h) return;
    const int r1 = min(d.h, r0 + kWinStatRows);
    const int nt = (256 / C) * C;
    const int t = threadIdx.x;
    const float INF = __int_as_float(0x7f800000);
    float mn = INF, mx = -INF, mmn = INF, mmx = -INF;
    int fl = 0, cnt = 0;
    if (t < nt) {
        const int c = t % C;
        const int ne = d.w * C;
        for (int r = r0; r < r1; ++r) {
            const float *row = raw + ((int64_t)(d.y0 + r) * Wl + d.x0) * C;
            const uint8_t *mrow = mask_slab ? mask_slab + (int64_t)(d.row0 + r) * slab_w : nullptr;
            for (int e = t; e < ne; e += nt) {
                const float v = row[e];
                fl |= classify_nonfinite(v);
                mn = fminf(mn, v);
                mx = fmaxf(mx, v);
                if (mrow) {
                    if (mrow[e / C]) {
                        mmn = fminf(mmn, v);
                        mmx = fmaxf(mmx, v);
                        if (c == 0) ++cnt;
                    }
                }
            }
        }
        uint32_t *k = keys + ((int64_t)blockIdx.y * C + c) * 4;
        if (mn <= mx) {
            atomicMin(k + 0, float_to_key(mn));
            atomicMax(k + 1, float_to_key(mx));
        }
        if (mask_slab && mmn <= mmx) {
            atomicMin(k + 2, float_to_key(mmn));
            atomicMax(k + 3, float_to_key(mmx));
        }
        if (fl) atomicOr(flags + (int64_t)blockIdx.y * C + c, fl);
        if (cnt) atomicAdd(counts + blockIdx.y, cnt);
    }
}

__global__ void win_stats_finish_kernel(uint32_t *keys, int64_t n, int has_mask)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    float *out = reinterpret_cast<float *>(keys);
    const float NANF = __int_as_float(0x7fc00000);
    const uint32_t k0 = keys[c * 4 + 0], k1 = keys[c * 4 + 1], k2 = keys[c * 4 + 2], k3 = keys[c * 4 + 3];
    const bool none = k0 == 0xffffffffu && k1 == 0u;
    const float mn = none ? NANF : key_to_float(k0), mx = none ? NANF : key_to_float(k1);
    float mmn, mmx;
    if (!has_mask) {
        mmn = mn; mmx = mx;
    } else if (k2 == 0xffffffffu && k3 == 0u) {
        mmn = NANF; mmx = NANF;
    } else {
        mmn = key_to_float(k2); mmx = key_to_float(k3);
    }
    out[c * 4 + 0] = mn; out[c * 4 + 1] = mx; out[c * 4 + 2] = mmn; out[c * 4 + 3] = mmx;
}

// window of a (H, Wm) mask raster -> slab rows (black tiles with a user mask)
__global__ void __launch_bounds__(256)
win_mask_copy_kernel(const uint8_t *__restrict__ mask, int64_t Wm, const WinDesc *__restrict__ batch,
                     uint8_t *__restrict__ mask_slab, int slab_w, int hw_max)
{
    const WinDesc &d = batch[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.h * d.w || i >= hw_max) return;
    const int y = i / d.w, x = i - y * d.w;
    mask_slab[(int64_t)(d.row0 + y) * slab_w + x] = mask[(int64_t)(d.y0 + y) * Wm + d.x0 + x] != 0;
}

// features_kernel per window: same float32 operation order; band ranges per window in bmin / bdiff [B][Cs]
__global__ void __launch_bounds__(256)
win_features_kernel(const float *__restrict__ raw, int64_t Wl, int C, int Cs, const int32_t *__restrict__ bands,
                    const float *__restrict__ bmin, const float *__restrict__ bdiff, const WinDesc *__restrict__ batch,
                    int to_lab, float ratio, float *__restrict__ feat, int64_t pitch, int64_t plane, int hw_max)
{
    const WinDesc &d = batch[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (!d.valid || i >= d.h * d.w || i >= hw_max) return;
    const int y = i / d.w, x = i - y * d.w;
    const float *src = raw + ((int64_t)(d.y0 + y) * Wl + d.x0 + x) * C;
    const float *mn = bmin + (int64_t)blockIdx.y * Cs, *df = bdiff + (int64_t)blockIdx.y * Cs;
    float *dst = feat + (int64_t)(d.row0 + y) * pitch + x;
    const float imin = d.imin, idiff = d.idiff;
    const int rescale = d.rescale;
    if (to_lab) {
        float v[3];
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            float r = src[bands[s]];
            r = __fdiv_rn(__fsub_rn(r, mn[s]), df[s]);
            r = __fsub_rn(r, imin);
            if (rescale) r = __fdiv_rn(r, idiff);
            v[s] = r;
        }
        float L, A, B;
        rgb2lab_f32(v[0], v[1], v[2], L, A, B);
        dst[0] = __fmul_rn(L, ratio);
        dst[plane] = __fmul_rn(A, ratio);
        dst[2 * plane] = __fmul_rn(B, ratio);
    } else {
        for (int s = 0; s < Cs; ++s) {
            float r = src[bands[s]];
            r = __fdiv_rn(__fsub_rn(r, mn[s]), df[s]);
            r = __fsub_rn(r, imin);
            if (rescale) r = __fdiv_rn(r, idiff);
            dst[(int64_t)s * plane] = __fmul_rn(r, ratio);
        }
    }
}

}  // namespace obia

using namespace obia;

extern "C" int obia_b200_band_minmax(const float *raw, int64_t n_pixels, int32_t C,
                                     const uint8_t *mask, float *out, int32_t *nonfinite,
                                     void *stream)
{
    if (!raw || !out || !nonfinite || n_pixels <= 0 || C <= 0 || C > 4096)
        return set_err(OBIA_B200_ERR_ARG, "band_minmax: bad argument");
    if (reinterpret_cast<uintptr_t>(raw) & 15)
        return set_err(OBIA_B200_ERR_ARG, "band_minmax: raw must be 16-byte aligned (128-bit loads)");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *keys = reinterpret_cast<uint32_t *>(out);
    minmax_init_kernel<<<(int)ceil_div(C, 128), 128, 0, st>>>(keys, nonfinite, C);
    OBIA_LAUNCH_CHECK();
    const int64_t n_elems = n_pixels * C;
    // grid such that 4 * gridDim * 256 is a multiple of C
    const int unit = C / gcd_i(C, 1024);
    int64_t want = ceil_div(n_elems / 4 + 1, 256 * 8);  // ~8 vectors per thread
    int64_t grid = std::min<int64_t>(want, (int64_t)kNumSMs * 8);
    grid = round_up(std::max<int64_t>(grid, 1), unit);
    const size_t smem = (size_t)C * 5 * sizeof(uint32_t);
    if (mask)
        minmax_kernel<true><<<(int)grid, 256, smem, st>>>(raw, n_elems, C, mask, keys, nonfinite);
    else
        minmax_kernel<false><<<(int)grid, 256, smem, st>>>(raw, n_elems, C, mask, keys, nonfinite);
    OBIA_LAUNCH_CHECK();
    minmax_finish_kernel<<<(int)ceil_div(C, 128), 128, 0, st>>>(keys, C, mask ? 1 : 0);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

extern "C" int obia_b200_normalize_inplace(float *raw, int64_t n_pixels, int32_t C,
                                           const float *minmax, void *stream)
{
    if (!raw || !minmax || n_pixels <= 0 || C <= 0)
        return set_err(OBIA_B200_ERR_ARG, "normalize_inplace: bad argument");
    if (reinterpret_cast<uintptr_t>(raw) & 15)
        return set_err(OBIA_B200_ERR_ARG, "normalize_inplace: raw must be 16-byte aligned (128-bit loads)");
    const int64_t n_elems = n_pixels * C;
    const int unit = C / gcd_i(C, 1024);
    int64_t grid = std::min<int64_t>(ceil_div(n_elems / 4 + 1, 256 * 4), (int64_t)kNumSMs * 16);
    grid = round_up(std::max<int64_t>(grid, 1), unit);
    normalize_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(raw, raw, n_elems, C, minmax);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

extern "C" int obia_b200_normalize_to(const float *raw, float *out, int64_t n_pixels, int32_t C,
                                      const float *minmax, int32_t max_ctas, void *stream)
{
    if (!raw || !out || !minmax || n_pixels <= 0 || C <= 0)
        return set_err(OBIA_B200_ERR_ARG, "normalize_to: bad argument");
    if ((reinterpret_cast<uintptr_t>(raw) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
        return set_err(OBIA_B200_ERR_ARG, "normalize_to: raw and out must be 16-byte aligned (128-bit accesses)");
    const int64_t n_elems = n_pixels * C;
    const int unit = C / gcd_i(C, 1024);
    int64_t grid = std::min<int64_t>(ceil_div(n_elems / 4 + 1, 256 * 4), (int64_t)kNumSMs * 16);
    if (max_ctas > 0) grid = std::min<int64_t>(grid, max_ctas);
    grid = round_up(std::max<int64_t>(grid, 1), unit);
    normalize_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(raw, out, n_elems, C, minmax);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

extern "C" int obia_b200_slic_features(const float *raw, int64_t H, int64_t W, int32_t C,
                                       const int32_t *bands_host, int32_t Cs,
                                       const float *band_min_host, const float *band_max_host,
                                       float imin, float imax, int32_t to_lab, float ratio,
                                       float *features, int64_t pitch, void *stream)
{
    if (!raw || !features || !bands_host || !band_min_host || !band_max_host || H <= 0 || W <= 0 ||
        C <= 0 || Cs <= 0)
        return set_err(OBIA_B200_ERR_ARG, "slic_features: bad argument");
    if (Cs > OBIA_B200_MAX_BANDS)
        return set_err(OBIA_B200_ERR_UNSUPPORTED, "slic_features: at most %d segmentation bands",
                       OBIA_B200_MAX_BANDS);
    if (to_lab && Cs != 3) return set_err(OBIA_B200_ERR_ARG, "slic_features: Lab needs 3 bands");
    if ((reinterpret_cast<uintptr_t>(raw) & 15) || (reinterpret_cast<uintptr_t>(features) & 15))
        return set_err(OBIA_B200_ERR_ARG, "slic_features: raw and features must be 16-byte aligned");
    if (pitch < W || (pitch & 3)) return set_err(OBIA_B200_ERR_ARG, "slic_features: bad pitch");
    BandTable tab;
    memset(&tab, 0, sizeof(tab));
    for (int s = 0; s < Cs; ++s) {
        int b = bands_host[s];
        if (b < 0 || b >= C) return set_err(OBIA_B200_ERR_ARG, "slic_features: band %d out of range", b);
        tab.band[s] = b;
        tab.mn[s] = band_min_host[b];
        tab.d[s] = band_max_host[b] - band_min_host[b];  // float32 subtraction
    }
    const size_t smem = ((size_t)kFeatPix * C + (size_t)kFeatPix * C / 32 + 8) * sizeof(float);
    if (smem > 200 * 1024) return set_err(OBIA_B200_ERR_UNSUPPORTED, "slic_features: too many bands (%d)", C);
    if (smem > 48 * 1024)
        OBIA_CUDA_CHECK(cudaFuncSetAttribute(features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t n_pixels = H * W;
    const float idiff = imax - imin;
    // x / 1.0f is x: skip the IEEE division in the usual case (obia's per-band normalisation
    // already maps every band to [0, 1], so skimage's global rescale is the identity)
    const int rescale = (imax != imin && idiff != 1.0f) ? 1 : 0;
    features_kernel<<<(int)ceil_div(n_pixels, kFeatPix), kFeatPix, smem, (cudaStream_t)stream>>>(
        raw, n_pixels, W, C, Cs, tab, imin, idiff, rescale, to_lab, ratio, features, pitch, H * pitch);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

extern "C" int obia_b200_gaussian_planar(const float *in, float *tmp, float *out, int64_t H, int64_t W,
                                         int64_t pitch, int32_t Cf, const double *weights_y_host,
                                         int32_t radius_y, const double *weights_x_host,
                                         int32_t radius_x, float ratio, void *stream)
{
    if (!in || !tmp || !out || H <= 0 || W <= 0 || Cf <= 0 || tmp == in || tmp == out)
        return set_err(OBIA_B200_ERR_ARG, "gaussian_planar: bad argument");
    if (radius_y > kMaxGaussRadius || radius_x > kMaxGaussRadius || radius_y < 0 || radius_x < 0)
        return set_err(OBIA_B200_ERR_UNSUPPORTED, "gaussian_planar: radius > %d", kMaxGaussRadius);
    cudaStream_t st = (cudaStream_t)stream;
    if (radius_y <= kGaussParamRadius && radius_x <= kGaussParamRadius && out != in) {
        // fused tiled kernel; asynchronous (taps are kernel parameters)
        GaussTaps taps;
        memset(&taps, 0, sizeof(taps));
        for (int k = 0; k <= radius_y; ++k) taps.wy[k] = weights_y_host[radius_y + k];
        for (int k = 0; k <= radius_x; ++k) taps.wx[k] = weights_x_host[radius_x + k];
        const size_t smem = ((size_t)(kGaussTY + 2 * radius_y) + kGaussTY) * (kGaussTX + 2 * radius_x) * sizeof(float);
        if (smem <= 200 * 1024) {
            if (smem > 48 * 1024)
                OBIA_CUDA_CHECK(cudaFuncSetAttribute(gaussian_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     (int)smem));
            dim3 grid((unsigned)ceil_div(W, kGaussTX), (unsigned)ceil_div(H, kGaussTY));
            gaussian_tiled_kernel<<<grid, 256, smem, st>>>(in, out, (int)H, (int)W, pitch, Cf, radius_y, radius_x, taps,
                                                          ratio);
            OBIA_LAUNCH_CHECK();
            return OBIA_B200_OK;
        }
    }
    // very wide kernels (sigma > 15) or in-place calls: two global passes, taps in constant memory
    if (H > 65535) return set_err(OBIA_B200_ERR_UNSUPPORTED, "gaussian_planar: H > 65535 not supported by the wide path");
    // half kernels, index 0 = centre tap (weights are symmetric)
    double hw[2][kMaxGaussRadius + 1];
    memset(hw, 0, sizeof(hw));
    for (int k = 0; k <= radius_y; ++k) hw[0][k] = weights_y_host[radius_y + k];
    for (int k = 0; k <= radius_x; ++k) hw[1][k] = weights_x_host[radius_x + k];
    OBIA_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_gauss_w, hw, sizeof(hw), 0, cudaMemcpyHostToDevice, st));
    OBIA_CUDA_CHECK(cudaStreamSynchronize(st));  // hw is a stack buffer
    dim3 grid((unsigned)ceil_div(W, 256), (unsigned)H, (unsigned)Cf);
    gaussian_pass_kernel<0><<<grid, 256, 0, st>>>(in, tmp, H, W, pitch, radius_y, 1.0f, 0);
    OBIA_LAUNCH_CHECK();
    gaussian_pass_kernel<1><<<grid, 256, 0, st>>>(tmp, out, H, W, pitch, radius_x, ratio, 1);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

// ---- batched windows (tiled driver) -----------------------------------------------------------------------
// out: (B, C, 4) float32 = min, max, masked min, masked max; nonfinite: (B, C); mask_counts: (B) mask pixels.
extern "C" int obia_b200_window_stats(const float *raw, int64_t Wl, int32_t C, const void *descs, int64_t B,
                                      int32_t hmax, const uint8_t *mask_slab, int32_t slab_w, float *out,
                                      int32_t *nonfinite, int32_t *mask_counts, void *stream)
{
    if (!raw || !descs || !out || !nonfinite || !mask_counts || Wl <= 0 || C <= 0 || C > 256 || B <= 0 || hmax <= 0 ||
        B > 65535)
        return set_err(OBIA_B200_ERR_ARG, "window_stats: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *keys = reinterpret_cast<uint32_t *>(out);
    win_stats_init_kernel<<<(unsigned)ceil_div(B * C, 256), 256, 0, st>>>(keys, nonfinite, mask_counts, B, C);
    OBIA_LAUNCH_CHECK();
    dim3 grid((unsigned)ceil_div(hmax, kWinStatRows), (unsigned)B);
    win_stats_kernel<<<grid, 256, 0, st>>>(raw, Wl, C, (const WinDesc *)descs, mask_slab, slab_w, keys, nonfinite,
                                           mask_counts);
    OBIA_LAUNCH_CHECK();
    win_stats_finish_kernel<<<(unsigned)ceil_div(B * C, 256), 256, 0, st>>>(keys, B * C, mask_slab ? 1 : 0);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

extern "C" int obia_b200_window_mask_copy(const uint8_t *mask, int64_t Wm, const void *descs, int64_t B, int32_t hmax,
                                          int32_t wmax, uint8_t *mask_slab, int32_t slab_w, void *stream)
{
    if (!mask || !descs || !mask_slab || Wm <= 0 || B <= 0 || B > 65535 || hmax <= 0 || wmax <= 0 || slab_w < wmax)
        return set_err(OBIA_B200_ERR_ARG, "window_mask_copy: bad argument");
    dim3 grid((unsigned)ceil_div((int64_t)hmax * wmax, 256), (unsigned)B);
    win_mask_copy_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mask, Wm, (const WinDesc *)descs, mask_slab, slab_w,
                                                                  hmax * wmax);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

// features: (Cf, slab_rows, pitch) planar; bands (device, [Cs]); band_min / band_diff (device, [B][Cs]);
// imin / idiff / rescale per window come from the descriptors.
extern "C" int obia_b200_window_features(const float *raw, int64_t Wl, int32_t C, const int32_t *bands, int32_t Cs,
                                         const float *band_min, const float *band_diff, const void *descs, int64_t B,
                                         int32_t hmax, int32_t wmax, int32_t to_lab, float ratio, float *features,
                                         int64_t slab_rows, int64_t pitch, void *stream)
{
    if (!raw || !bands || !band_min || !band_diff || !descs || !features || Wl <= 0 || C <= 0 || Cs <= 0 || B <= 0 ||
        B > 65535 || hmax <= 0 || wmax <= 0 || pitch < wmax || slab_rows <= 0)
        return set_err(OBIA_B200_ERR_ARG, "window_features: bad argument");
    if (to_lab && Cs != 3) return set_err(OBIA_B200_ERR_ARG, "window_features: Lab needs 3 bands");
    dim3 grid((unsigned)ceil_div((int64_t)hmax * wmax, 256), (unsigned)B);
    win_features_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(raw, Wl, C, Cs, bands, band_min, band_diff,
                                                                 (const WinDesc *)descs, to_lab, ratio, features, pitch,
                                                                 slab_rows * pitch, hmax * wmax);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

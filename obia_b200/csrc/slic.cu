// K2: SLIC iterations (replaces skimage `_slic_cython`, reached from
// obia/segmentation/segment_boundaries.py:51).
//
// The reference scatters: for every centre k (ascending) it visits the pixels
// of its +-2*step window and keeps the centre when `distance > d` (strict).
// Here the loop is inverted into a gather that gives the same result:
//   * centres are binned by their CURRENT position into a step-sized cell
//     grid (linked lists rebuilt every iteration),
//   * a CTA owns a 32 x (8*PX) pixel tile, collects every centre whose window
//     can reach the tile into shared memory, and every thread evaluates
//     exactly the centres whose truncated window contains its pixel, in the
//     reference's float32 operation order (no FMA contraction), taking the
//     lexicographic minimum of (distance, k)  ==  "first strict improvement in
//     ascending k",
//   * the centre update is fused: per warp, pixels are grouped by winning
//     centre, reduced with shuffles in a fixed order, converted to 64-bit
//     fixed point and added with one coalesced RED.64 per field, so sums are
//     independent of scheduling.
// Features are band-planar so each lane streams 128-bit loads (4 pixels).
#include "common.cuh"

namespace obia {

constexpr int kIdCap = 2048;  // centre ids a tile can collect
constexpr int kChunk = 64;    // centre records resident in shared memory at once
constexpr int kWarps = 8;

// workspace layout (all 16-byte aligned)
struct SlicWs {
    unsigned long long *acc;  // [n][3+Cf] count, sum y, sum x, fixed-point colour sums
    int32_t *head;            // [ncy*ncx] cell -> first centre
    int32_t *next;            // [n]
    int64_t ncy, ncx;
    int64_t bytes;
};

static SlicWs slic_ws_layout(void *base, int64_t H, int64_t W, int Cf, int64_t n, int step_y, int step_x)
{
    SlicWs w;
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
    w.bytes = off;
    return w;
}

// ------------------------------------------------------------------------
// centres: finalise the means of the previous iteration (if from_acc) and
// bin every live centre into its cell list.
__global__ void __launch_bounds__(256)
slic_centres_kernel(float *centres, unsigned long long *acc, int32_t *head, int32_t *next, int64_t n,
                    int Cf, int from_acc, double inv_fix, int step_y, int step_x, int64_t ncy,
                    int64_t ncx)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int rec = 2 + Cf;
    float *c = centres + k * rec;
    if (from_acc) {
        unsigned long long *a = acc + k * (3 + Cf);
        const long long cnt = (long long)a[0];
        if (cnt > 0) {
            const double dc = (double)cnt;
            c[0] = (float)((double)(long long)a[1] / dc);
            c[1] = (float)((double)(long long)a[2] / dc);
            for (int f = 0; f < Cf; ++f) c[2 + f] = (float)((double)(long long)a[3 + f] * inv_fix / dc);
        } else {
            const float NANF = __int_as_float(0x7fc00000);  // 0/0 in the reference
            for (int f = 0; f < rec; ++f) c[f] = NANF;
        }
        for (int f = 0; f < 3 + Cf; ++f) a[f] = 0ull;
    }
    const float cy = c[0], cx = c[1];
    if (!(cy == cy) || !(cx == cx)) {  // dead centre: never wins a comparison
        next[k] = -1;
        return;
    }
    int64_t gy = (int64_t)floorf(cy / (float)step_y);
    int64_t gx = (int64_t)floorf(cx / (float)step_x);
    gy = max((int64_t)0, min(ncy - 1, gy));
    gx = max((int64_t)0, min(ncx - 1, gx));
    next[k] = atomicExch(&head[gy * ncx + gx], (int32_t)k);
}

__device__ __forceinline__ int floordiv_i(int a, int b)
{
    int q = a / b;
    return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// C cast float -> integer as the Cython code does (`<Py_ssize_t>`): truncation
__device__ __forceinline__ int trunc_i(float v) { return (int)v; }

// ---- packed fp32x2 arithmetic (sm_100a FADD2 / FFMA2): two pixels per instruction ------
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b)
{
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 pack2(float lo, float hi)
{
    u64 d;
    asm("mov.b64 %0, {%1,%2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi)
{
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

// Per-candidate record in shared memory.
struct __align__(16) CandHead {
    float cy, cx;
    int k, pad;
};

// Distance of the PX pixels of this lane to one candidate centre, in the reference's
// operation order:  d = ((dy + dx) * w) + sum_c fma(t_c, t_c, .)   (t_c = pixel_c - centre_c).
// Everything except the colour FMA is separately rounded (scalar FMUL feeding FADD2 is never
// contracted).  CHECK = per-pixel window test (skipped when the whole warp strip is inside).
template <int CP, int PX, bool CHECK>
__device__ __forceinline__ void eval_candidate(const u64 (&px2)[(PX + 1) / 2][CP], const float (&px1)[CP],
                                               const u64 (&nx2)[(PX + 1) / 2], float nx1, float fy,
                                               const CandHead h, const int4 w, const float *__restrict__ nf,
                                               float spatial_weight, int ignore_color, int y, int xb,
                                               float (&best)[PX], int (&bestk)[PX])
{
    int lo = 0, span = PX;
    if (CHECK) {
        lo = w.z - xb;
        span = (y >= w.x && y < w.y) ? (w.w - w.z) : 0;  // row outside the window: nothing valid
    }
    const float ty = __fsub_rn(h.cy, fy);
    const float dy = __fmul_rn(ty, ty);
    if constexpr (PX >= 2) {
#pragma unroll
        for (int p = 0; p < PX / 2; ++p) {
            const u64 tx2 = add2(nx2[p], pack2(h.cx, h.cx));  // cx - x  ==  cx + (-x)
            float tx0, tx1;
            unpack2(tx2, tx0, tx1);
            const u64 s2 = add2(pack2(dy, dy), pack2(__fmul_rn(tx0, tx0), __fmul_rn(tx1, tx1)));
            float s0, s1;
            unpack2(s2, s0, s1);
            u64 d2 = pack2(__fmul_rn(s0, spatial_weight), __fmul_rn(s1, spatial_weight));
            if (!ignore_color) {
                u64 acc = 0ull;
#pragma unroll
                for (int c = 0; c < CP; ++c) {
                    const float m = nf[c];  // negated centre colour, broadcast to both halves
                    const u64 t2 = add2(px2[p][c], pack2(m, m));
                    acc = fma2(t2, t2, acc);
                }
                d2 = add2(d2, acc);
            }
            float d0, d1;
            unpack2(d2, d0, d1);
            const int j0 = 2 * p, j1 = 2 * p + 1;
            const bool v0 = !CHECK || (unsigned)(j0 - lo) < (unsigned)span;
            const bool v1 = !CHECK || (unsigned)(j1 - lo) < (unsigned)span;
            if (v0 && d0 < best[j0]) { best[j0] = d0; bestk[j0] = h.k; }
            if (v1 && d1 < best[j1]) { best[j1] = d1; bestk[j1] = h.k; }
        }
    } else {
        const float tx = __fadd_rn(h.cx, nx1);
        float d = __fmul_rn(__fadd_rn(dy, __fmul_rn(tx, tx)), spatial_weight);
        if (!ignore_color) {
            float acc = 0.0f;
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                const float t = __fadd_rn(px1[c], nf[c]);
                acc = __fmaf_rn(t, t, acc);
            }
            d = __fadd_rn(d, acc);
        }
        const bool v0 = !CHECK || (unsigned)(0 - lo) < (unsigned)span;
        if (v0 && d < best[0]) { best[0] = d; bestk[0] = h.k; }
    }
}

template <int CP, int PX>
__global__ void __launch_bounds__(kWarps * 32)
slic_assign_update_kernel(const float *__restrict__ feat, const uint8_t *__restrict__ mask,
                          const float *__restrict__ centres, const int32_t *__restrict__ head,
                          const int32_t *__restrict__ next, int32_t *__restrict__ labels,
                          unsigned long long *__restrict__ acc, int H, int W, int64_t pitch, int Cf,
                          float spatial_weight, int step_y, int step_x, int ncy, int ncx,
                          int start_label, int ignore_color, double fix_scale, int32_t *status)
{
    constexpr int LX = 32 / PX;  // lanes along x
    constexpr int TH = kWarps * PX;
    constexpr int NP = (PX + 1) / 2;
    __shared__ int s_ids[kIdCap];
    __shared__ int s_sorted[kIdCap];
    __shared__ int s_nids;
    __shared__ int4 s_win[kChunk];
    __shared__ CandHead s_head[kChunk];
    __shared__ __align__(16) float s_nf[kChunk][CP];  // negated centre colours

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * 32, ty0 = blockIdx.y * TH;
    const int tx1 = min(tx0 + 32, W) - 1, ty1 = min(ty0 + TH, H) - 1;  // inclusive

    // ---- collect candidate centre ids -----------------------------------
    if (tid == 0) s_nids = 0;
    __syncthreads();
    {
        // a centre at cy reaches rows y with  y - 2s <= cy < y + 1 + 2s; two pixels
        // of slack absorb the float rounding of the cell index.  Centres always lie
        // inside the raster, so cells need no clamping beyond the grid itself.
        const int gy_lo = max(0, floordiv_i(ty0 - 2 * step_y - 2, step_y));
        const int gy_hi = min(ncy - 1, floordiv_i(ty1 + 2 * step_y + 2, step_y));
        const int gx_lo = max(0, floordiv_i(tx0 - 2 * step_x - 2, step_x));
        const int gx_hi = min(ncx - 1, floordiv_i(tx1 + 2 * step_x + 2, step_x));
        const int ny = gy_hi - gy_lo + 1, nx = gx_hi - gx_lo + 1;
        for (int i = tid; i < ny * nx; i += kWarps * 32) {
            const int gy = gy_lo + i / nx, gx = gx_lo + i % nx;
            int k = head[(int64_t)gy * ncx + gx];
            while (k >= 0) {
                const int slot = atomicAdd(&s_nids, 1);
                if (slot < kIdCap) s_ids[slot] = k;
                k = next[k];
            }
        }
    }
    __syncthreads();
    int nids = s_nids;
    if (nids > kIdCap) {
        if (tid == 0) atomicExch(&status[0], 1);
        nids = kIdCap;
    }
    // ascending centre index: "first strict improvement" then equals the reference's
    // lowest-k-wins tie rule (ids are unique, so the rank is a permutation)
    for (int i = tid; i < nids; i += kWarps * 32) {
        const int k = s_ids[i];
        int r = 0;
        for (int j = 0; j < nids; ++j) r += (s_ids[j] < k);
        s_sorted[r] = k;
    }

    // ---- this lane's pixels ----------------------------------------------
    const int y = ty0 + warp * PX + lane / LX;
    const int xb = tx0 + (lane % LX) * PX;
    const bool row_ok = y < H;
    u64 px2[NP][CP];
    float px1[CP];
    bool valid[PX];
#pragma unroll
    for (int j = 0; j < PX; ++j) {
        valid[j] = row_ok && (xb + j) < W;
        if (valid[j] && mask) valid[j] = mask[(int64_t)y * W + xb + j] != 0;
    }
    // features are always loaded: the centre update sums colours even in the
    // spatial-only (ignore_color) pass of masked SLIC
#pragma unroll
    for (int c = 0; c < CP; ++c) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < Cf && row_ok && xb < W) {
            const float *src = feat + (int64_t)c * H * pitch + (int64_t)y * pitch + xb;
            if constexpr (PX == 4) {
                v = *reinterpret_cast<const float4 *>(src);
            } else if constexpr (PX == 2) {
                const float2 t = *reinterpret_cast<const float2 *>(src);
                v.x = t.x; v.y = t.y;
            } else {
                v.x = *src;
            }
        }
        px1[c] = v.x;
        px2[0][c] = pack2(v.x, v.y);
        if constexpr (PX == 4) px2[NP - 1][c] = pack2(v.z, v.w);
    }
    u64 nx2[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) nx2[p] = pack2(-(float)(xb + 2 * p), -(float)(xb + 2 * p + 1));
    const float nx1 = -(float)xb;

    const float INF = __int_as_float(0x7f800000);
    float best[PX];
    int bestk[PX];
#pragma unroll
    for (int j = 0; j < PX; ++j) {
        best[j] = INF;
        bestk[j] = -1;
    }
    // warp strip (inclusive), clipped to the image
    const int wy0 = ty0 + warp * PX, wy1 = min(wy0 + PX, H) - 1;
    const float fy = (float)y;

    // ---- evaluate candidates chunk by chunk (ascending k) ---------------------
    for (int c0 = 0; c0 < nids; c0 += kChunk) {
        const int nc = min(kChunk, nids - c0);
        __syncthreads();  // previous chunk fully consumed (and s_sorted complete)
        for (int i = tid; i < nc * (2 + CP); i += kWarps * 32) {
            const int s = i / (2 + CP), f = i % (2 + CP);
            const int k = s_sorted[c0 + s];
            const float *rec = centres + (int64_t)k * (2 + Cf);
            if (f == 0) {
                const float cy = rec[0], cx = rec[1];
                CandHead h;
                h.cy = cy; h.cx = cx; h.k = k; h.pad = 0;
                s_head[s] = h;
                // windows exactly as the reference computes them (float32, then C cast)
                const float ylo = __fsub_rn(cy, (float)(2 * step_y));
                const float yhi = __fadd_rn(__fadd_rn(cy, (float)(2 * step_y)), 1.0f);
                const float xlo = __fsub_rn(cx, (float)(2 * step_x));
                const float xhi = __fadd_rn(__fadd_rn(cx, (float)(2 * step_x)), 1.0f);
                int4 w;
                w.x = trunc_i((0.0f > ylo) ? 0.0f : ylo);
                w.y = trunc_i(((float)H < yhi) ? (float)H : yhi);
                w.z = trunc_i((0.0f > xlo) ? 0.0f : xlo);
                w.w = trunc_i(((float)W < xhi) ? (float)W : xhi);
                s_win[s] = w;
            } else if (f >= 2) {
                const int c = f - 2;
                s_nf[s][c] = (c < Cf) ? -rec[2 + c] : 0.0f;
            }
        }
        __syncthreads();

        // every lane tests two candidates against the warp strip; the warp then walks
        // the surviving ones in ascending order
#pragma unroll
        for (int half = 0; half < kChunk / 32; ++half) {
            const int sc = half * 32 + lane;
            bool hit = false, full = false;
            if (sc < nc) {
                const int4 w = s_win[sc];
                hit = !(w.x > wy1 || w.y <= wy0 || w.z > tx1 || w.w <= tx0);
                full = w.x <= wy0 && w.y > wy1 && w.z <= tx0 && w.w > tx1;
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            const unsigned mfull = __ballot_sync(0xffffffffu, full);
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                const int s = half * 32 + b;
                const CandHead h = s_head[s];
                const int4 w = s_win[s];
                if (mfull & (1u << b))
                    eval_candidate<CP, PX, false>(px2, px1, nx2, nx1, fy, h, w, s_nf[s], spatial_weight,
                                                  ignore_color, y, xb, best, bestk);
                else
                    eval_candidate<CP, PX, true>(px2, px1, nx2, nx1, fy, h, w, s_nf[s], spatial_weight,
                                                 ignore_color, y, xb, best, bestk);
            }
        }
    }

    // ---- labels ------------------------------------------------------------
    int kk[PX];
    bool all_found = true;
#pragma unroll
    for (int j = 0; j < PX; ++j) {
        kk[j] = -1;
        if (valid[j]) {
            if (bestk[j] >= 0) {
                kk[j] = bestk[j];
            } else {
                kk[j] = labels[(int64_t)y * W + xb + j] - start_label;  // no window reached the pixel: keep
                all_found = false;
            }
        } else {
            all_found = false;
        }
    }
    if (PX == 4 && all_found && (W & 3) == 0) {
        *reinterpret_cast<int4 *>(labels + (int64_t)y * W + xb) =
            make_int4(kk[0] + start_label, kk[1 % PX] + start_label, kk[2 % PX] + start_label,
                      kk[3 % PX] + start_label);
    } else {
#pragma unroll
        for (int j = 0; j < PX; ++j)
            if (valid[j] && bestk[j] >= 0) labels[(int64_t)y * W + xb + j] = bestk[j] + start_label;
    }

    // ---- fused centre update: group by winner inside the warp ---------------
    unsigned pending = 0;
#pragma unroll
    for (int j = 0; j < PX; ++j)
        if (kk[j] >= 0) pending |= 1u << j;
    while (true) {
        int mine = -1;
#pragma unroll
        for (int j = PX - 1; j >= 0; --j)
            if (pending & (1u << j)) mine = kk[j];
        const unsigned vote = __ballot_sync(0xffffffffu, mine >= 0);
        if (!vote) break;
        const int L = __shfl_sync(0xffffffffu, mine, __ffs(vote) - 1);
        int cnt = 0, sy = 0, sx = 0;
        float fs[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) fs[c] = 0.0f;
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            if ((pending & (1u << j)) && kk[j] == L) {
                pending &= ~(1u << j);
                cnt += 1;
                sy += y - ty0;  // tile-local, re-based below
                sx += xb + j - tx0;
#pragma unroll
                for (int c = 0; c < CP; ++c) {
                    float lo, hi;
                    unpack2(px2[(j / 2) % NP][c], lo, hi);
                    fs[c] = __fadd_rn(fs[c], (PX == 1) ? px1[c] : ((j & 1) ? hi : lo));
                }
            }
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        sy = __reduce_add_sync(0xffffffffu, sy);
        sx = __reduce_add_sync(0xffffffffu, sx);
#pragma unroll
        for (int c = 0; c < CP; ++c) {
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) fs[c] = __fadd_rn(fs[c], __shfl_xor_sync(0xffffffffu, fs[c], o));
        }
        // lane f adds field f (coalesced 64-bit reductions)
        unsigned long long *a = acc + (int64_t)L * (3 + Cf);
        if (lane == 0) atomicAdd(&a[0], (unsigned long long)cnt);
        if (lane == 1) atomicAdd(&a[1], (unsigned long long)((long long)sy + (long long)cnt * ty0));
        if (lane == 2) atomicAdd(&a[2], (unsigned long long)((long long)sx + (long long)cnt * tx0));
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            if (c < Cf && lane == ((3 + c) & 31)) {
                const long long q = __double2ll_rn((double)fs[c] * fix_scale);
                atomicAdd(&a[3 + c], (unsigned long long)q);
            }
        }
    }
}

template <int CP, int PX>
static int launch_assign(const float *feat, const uint8_t *mask, const float *centres, const SlicWs &w,
                         int32_t *labels, int64_t H, int64_t W, int64_t pitch, int Cf, float sw,
                         int step_y, int step_x, int start_label, int ignore_color, double fix_scale,
                         int32_t *status, cudaStream_t st)
{
    dim3 grid((unsigned)ceil_div(W, 32), (unsigned)ceil_div(H, kWarps * PX));
    prof_begin(st);
    slic_assign_update_kernel<CP, PX><<<grid, kWarps * 32, 0, st>>>(
        feat, mask, centres, w.head, w.next, labels, w.acc, (int)H, (int)W, pitch, Cf, sw, step_y, step_x,
        (int)w.ncy, (int)w.ncx, start_label, ignore_color, fix_scale, status);
    prof_end(st);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

__global__ void fill_i32_kernel(int32_t *p, int64_t n, int32_t v)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

}  // namespace obia

using namespace obia;

extern "C" int64_t obia_b200_slic_workspace_bytes(int64_t H, int64_t W, int32_t Cf, int64_t n,
                                                  int32_t step_y, int32_t step_x)
{
    if (H <= 0 || W <= 0 || Cf <= 0 || n <= 0 || step_y <= 0 || step_x <= 0) return -1;
    return slic_ws_layout(nullptr, H, W, Cf, n, step_y, step_x).bytes;
}

extern "C" int obia_b200_slic_iterate(const float *features, const uint8_t *mask, float *centres,
                                      int32_t *labels, void *workspace, int64_t H, int64_t W,
                                      int64_t pitch, int32_t Cf, int64_t n, float step, int32_t step_y,
                                      int32_t step_x, int32_t max_num_iter, int32_t start_label,
                                      int32_t ignore_color, double fix_scale, int32_t *status,
                                      void *stream)
{
    if (!features || !centres || !labels || !workspace || !status || H <= 0 || W <= 0 || Cf <= 0 ||
        n <= 0 || step_y <= 0 || step_x <= 0 || max_num_iter < 0 || !(step > 0.0f) || !(fix_scale > 0.0))
        return set_err(OBIA_B200_ERR_ARG, "slic_iterate: bad argument");
    if (H > 2000000000 / 1 || W > 2000000000 || n > 2000000000)
        return set_err(OBIA_B200_ERR_UNSUPPORTED, "slic_iterate: dimension exceeds int32");
    if (pitch < W || (pitch & 3)) return set_err(OBIA_B200_ERR_ARG, "slic_iterate: pitch must be a multiple of 4 and >= W");
    if (Cf > 64) return set_err(OBIA_B200_ERR_UNSUPPORTED, "slic_iterate: more than 64 feature channels");
    if (start_label != 0 && start_label != 1) return set_err(OBIA_B200_ERR_ARG, "start_label should be 0 or 1.");
    cudaStream_t st = (cudaStream_t)stream;
    SlicWs w = slic_ws_layout(workspace, H, W, Cf, n, step_y, step_x);

    // `1.0 / (step * step)`: float product, double division, float store
    const float step_sq = step * step;
    const float sw = (float)(1.0 / (double)step_sq);

    OBIA_CUDA_CHECK(cudaMemsetAsync(status, 0, 4 * sizeof(int32_t), st));
    OBIA_CUDA_CHECK(cudaMemsetAsync(w.acc, 0, (size_t)n * (3 + Cf) * 8, st));
    fill_i32_kernel<<<kNumSMs * 4, 256, 0, st>>>(labels, H * W, start_label - 1);
    OBIA_LAUNCH_CHECK();

    for (int it = 0; it < max_num_iter; ++it) {
        OBIA_CUDA_CHECK(cudaMemsetAsync(w.head, 0xff, (size_t)(w.ncy * w.ncx) * 4, st));
        slic_centres_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(
            centres, w.acc, w.head, w.next, n, Cf, it > 0 ? 1 : 0, 1.0 / fix_scale, step_y, step_x, w.ncy,
            w.ncx);
        OBIA_LAUNCH_CHECK();
        int rc;
        if (Cf <= 4)
            rc = launch_assign<4, 4>(features, mask, centres, w, labels, H, W, pitch, Cf, sw, step_y, step_x,
                                     start_label, ignore_color, fix_scale, status, st);
        else if (Cf <= 8)
            rc = launch_assign<8, 4>(features, mask, centres, w, labels, H, W, pitch, Cf, sw, step_y, step_x,
                                     start_label, ignore_color, fix_scale, status, st);
        else if (Cf <= 16)
            rc = launch_assign<16, 2>(features, mask, centres, w, labels, H, W, pitch, Cf, sw, step_y, step_x,
                                      start_label, ignore_color, fix_scale, status, st);
        else if (Cf <= 32)
            rc = launch_assign<32, 1>(features, mask, centres, w, labels, H, W, pitch, Cf, sw, step_y, step_x,
                                      start_label, ignore_color, fix_scale, status, st);
        else
            rc = launch_assign<64, 1>(features, mask, centres, w, labels, H, W, pitch, Cf, sw, step_y, step_x,
                                      start_label, ignore_color, fix_scale, status, st);
        if (rc) return rc;
    }
    if (max_num_iter > 0) {
        // means of the last assignment (the reference updates after every sweep)
        OBIA_CUDA_CHECK(cudaMemsetAsync(w.head, 0xff, (size_t)(w.ncy * w.ncx) * 4, st));
        slic_centres_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(
            centres, w.acc, w.head, w.next, n, Cf, 1, 1.0 / fix_scale, step_y, step_x, w.ncy, w.ncx);
        OBIA_LAUNCH_CHECK();
    }
    return OBIA_B200_OK;
}

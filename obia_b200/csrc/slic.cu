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
#include "batch.cuh"
#include "slic_common.cuh"

namespace obia {

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

// batched form (tiled driver): centre k belongs to window cwin[k]; its cell grid, fixed-point scale and
// first centre come from the window descriptor, `next` holds window-local indices.
__global__ void __launch_bounds__(256)
slic_centres_batch_kernel(float *centres, unsigned long long *acc, int32_t *head, int32_t *next, int64_t n,
                          int Cf, int from_acc, const int32_t *__restrict__ cwin, const WinDesc *__restrict__ batch)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const WinDesc *d = batch + cwin[k];
    if (!d->valid) return;
    const int rec = 2 + Cf;
    float *c = centres + k * rec;
    if (from_acc) {
        unsigned long long *a = acc + k * (3 + Cf);
        const long long cnt = (long long)a[0];
        if (cnt > 0) {
            const double dc = (double)cnt;
            const double inv_fix = 1.0 / d->fix_scale;
            c[0] = (float)((double)(long long)a[1] / dc);
            c[1] = (float)((double)(long long)a[2] / dc);
            for (int f = 0; f < Cf; ++f) c[2 + f] = (float)((double)(long long)a[3 + f] * inv_fix / dc);
        } else {
            const float NANF = __int_as_float(0x7fc00000);
            for (int f = 0; f < rec; ++f) c[f] = NANF;
        }
        for (int f = 0; f < 3 + Cf; ++f) a[f] = 0ull;
    }
    const float cy = c[0], cx = c[1];
    if (!(cy == cy) || !(cx == cx)) {
        next[k] = -1;
        return;
    }
    int64_t gy = (int64_t)floorf(cy / (float)d->step_y);
    int64_t gx = (int64_t)floorf(cx / (float)d->step_x);
    gy = max((int64_t)0, min((int64_t)d->ncy - 1, gy));
    gx = max((int64_t)0, min((int64_t)d->ncx - 1, gx));
    next[k] = atomicExch(&head[d->cell0 + gy * d->ncx + gx], (int32_t)(k - d->c0));
}

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


// Distance of the PX pixels of this lane to one candidate centre, in the reference's
// operation order:  d = ((dy + dx) * w) + sum_c fma(t_c, t_c, .)   (t_c = pixel_c - centre_c).
// Everything except the colour FMA is separately rounded (scalar FMUL feeding FADD2 is never
// contracted).  CHECK = per-pixel window test (skipped when the whole warp strip is inside).
// Candidate slots are sorted by centre index and visited in ascending order, so a strict
// "less than" update is exactly the reference's rule (lowest k wins exact ties).
template <int CP, int PX, bool CHECK, bool SZ>
__device__ __forceinline__ void eval_candidate(const u64 (&px2)[(PX + 1) / 2][CP], const float (&px1)[CP],
                                               const u64 (&nx2)[(PX + 1) / 2], float nx1, float fy,
                                               float cy, float cx, const int4 w,
                                               const float *__restrict__ nf, float mdc, float spatial_weight,
                                               float sp_y, float sp_x, int ignore_color, int y, int xb, int slot,
                                               float (&best)[PX], int (&bests)[PX])
{
    int lo = 0, span = PX;
    if (CHECK) {
        lo = w.z - xb;
        span = (y >= w.x && y < w.y) ? (w.w - w.z) : 0;  // row outside the window: nothing valid
    }
    // skimage: dy = (sy * (cy - y)) ** 2 with the `spacing` of the call (a product with 1.0f is exact)
    const float ty = __fmul_rn(sp_y, __fsub_rn(cy, fy));
    const float dy = __fmul_rn(ty, ty);
    if constexpr (PX >= 2) {
#pragma unroll
        for (int p = 0; p < PX / 2; ++p) {
            const u64 tx2 = add2(nx2[p], pack2(cx, cx));  // cx - x  ==  cx + (-x)
            float tx0, tx1;
            unpack2(tx2, tx0, tx1);
            tx0 = __fmul_rn(sp_x, tx0);
            tx1 = __fmul_rn(sp_x, tx1);
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
                if (SZ) {   // SLICO: colour distance relative to the centre's running maximum
                    float a0, a1;
                    unpack2(acc, a0, a1);
                    acc = pack2(__fdiv_rn(a0, mdc), __fdiv_rn(a1, mdc));
                }
                d2 = add2(d2, acc);
            }
            float d[2];
            unpack2(d2, d[0], d[1]);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int j = 2 * p + q;
                const bool v = !CHECK || (unsigned)(j - lo) < (unsigned)span;
                if (v && d[q] < best[j]) {
                    best[j] = d[q];
                    bests[j] = slot;
                }
            }
        }
    } else {
        const float tx = __fmul_rn(sp_x, __fadd_rn(cx, nx1));
        float d = __fmul_rn(__fadd_rn(dy, __fmul_rn(tx, tx)), spatial_weight);
        if (!ignore_color) {
            float acc = 0.0f;
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                const float t = __fadd_rn(px1[c], nf[c]);
                acc = __fmaf_rn(t, t, acc);
            }
            if (SZ) acc = __fdiv_rn(acc, mdc);
            d = __fadd_rn(d, acc);
        }
        const bool v = !CHECK || (unsigned)(0 - lo) < (unsigned)span;
        if (v && d < best[0]) {
            best[0] = d;
            bests[0] = slot;
        }
    }
}

// shared-memory sizing per channel count (everything static, < 48 KB)
template <int CP> struct Traits {
    static constexpr int kIds = 1024;   // centre ids a tile can collect
    static constexpr int kChk = (CP <= 32) ? 64 : 32;       // centre records resident at once
    static constexpr int kAcc = (CP <= 16) ? 128 : 64;      // slots with a tile accumulator row
    static constexpr int kRec = (CP <= 4) ? 768 : (CP == 8) ? 640 : (CP == 16) ? 288 : 512;
    // >= 32 channels run one CTA per SM (registers), so the record pool can live in dynamic shared
    // memory beyond the 48 KB static limit and hold a whole strip phase
    static constexpr bool kDynRec = CP >= 32;
};

// Warp strip: 16 pixels wide x (32 / (16/PX)) rows; a CTA tile is 2 x 4 strips (32 px wide) times
// NS vertical phases: the candidate list, its sort, the centre records and the tile accumulators
// are set up once and reused by NS strips per warp, which amortises the latency-bound set-up.
template <int CP, int PX, int NS, bool SZ>
__global__ void __launch_bounds__(kWarps * 32, (CP <= 16) ? 3 : 1)
slic_assign_update_kernel(const float *__restrict__ feat, const uint8_t *__restrict__ mask,
                          const float *__restrict__ centres, const float *__restrict__ maxdc,
                          const int32_t *__restrict__ head,
                          const int32_t *__restrict__ next, int32_t *__restrict__ labels,
                          unsigned long long *__restrict__ acc, int H, int W, int64_t pitch, int Cf,
                          float spatial_weight, int step_y, int step_x, int ncy, int ncx,
                          int start_label, int ignore_color, double fix_scale, float fix_scale32,
                          long long fix_ratio, int32_t *status, int y_off, int Hg, float sp_y, float sp_x)
{
    // (H, W) is the strip resident on this GPU; it is rows [y_off, y_off + H) of a raster with Hg
    // rows.  Memory is addressed with strip-local rows, the arithmetic (centre windows, spatial
    // term, centre sums) uses global rows, so a sharded run is bit-identical to a single-GPU one.
    constexpr int LPR = 16 / PX;   // lanes per strip row
    constexpr int RW = 32 / LPR;   // rows per warp strip
    constexpr int TH = 4 * RW;     // tile rows (tile is 32 wide)
    constexpr int NP = (PX + 1) / 2;
    constexpr int NF = 3 + CP;     // accumulator fields per slot (count, sum y, sum x, colours)
    constexpr int kIds = Traits<CP>::kIds, kChk = Traits<CP>::kChk, kAcc = Traits<CP>::kAcc,
                  kRec = Traits<CP>::kRec;
    constexpr bool kRounds = CP >= 16;   // record pool smaller than a strip phase: several fold rounds
    constexpr int NT = kWarps * 32;
    __shared__ int s_ids[kIds];      // as collected
    __shared__ int s_sorted[kIds];   // ascending centre index
    __shared__ int s_nids, s_nrec;
    __shared__ int4 s_win[kChk];
    __shared__ float2 s_cyx[kChk];
    __shared__ __align__(16) float s_nf[kChk][CP];  // negated centre colours
    __shared__ float s_mdc[SZ ? kChk : 1];          // SLICO colour-distance maxima
    __shared__ int s_acc[kAcc][NF];
    extern __shared__ __align__(16) int s_dyn[];
    __shared__ __align__(16) int s_rec_static[Traits<CP>::kDynRec ? 1 : kRec][NF + 1];
    int(*s_rec)[NF + 1] = Traits<CP>::kDynRec ? reinterpret_cast<int(*)[NF + 1]>(s_dyn) : s_rec_static;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * 32, ty0 = blockIdx.y * (TH * NS);
    const int tx1 = min(tx0 + 32, W) - 1, ty1 = min(ty0 + TH * NS, H) - 1;  // inclusive

    // ---- collect candidate centre ids -----------------------------------
    if (tid == 0) {
        s_nids = 0;
        s_nrec = 0;
    }
    for (int i = tid; i < kAcc * NF; i += NT) (&s_acc[0][0])[i] = 0;
    __syncthreads();
    {
        // a centre at cy reaches rows y with  y - 2s <= cy < y + 1 + 2s; two pixels
        // of slack absorb the float rounding of the cell index.  Centres always lie
        // inside the raster, so cells need no clamping beyond the grid itself.
        const int gy_lo = max(0, floordiv_i(ty0 + y_off - 2 * step_y - 2, step_y));
        const int gy_hi = min(ncy - 1, floordiv_i(ty1 + y_off + 2 * step_y + 2, step_y));
        const int gx_lo = max(0, floordiv_i(tx0 - 2 * step_x - 2, step_x));
        const int gx_hi = min(ncx - 1, floordiv_i(tx1 + 2 * step_x + 2, step_x));
        const int ny = gy_hi - gy_lo + 1, nx = gx_hi - gx_lo + 1;
        for (int i = tid; i < ny * nx; i += NT) {
            const int gy = gy_lo + i / nx, gx = gx_lo + i % nx;
            int k = head[(int64_t)gy * ncx + gx];
            while (k >= 0) {
                const int slot = atomicAdd(&s_nids, 1);
                if (slot < kIds) s_ids[slot] = k;
                k = next[k];
            }
        }
    }
    __syncthreads();
    int nids = s_nids;
    if (nids > kIds) {
        if (tid == 0) atomicExch(&status[0], 1);
        nids = kIds;
    }
    // rank sort (ids are unique): slot order == centre-index order
    for (int i = tid; i < nids; i += NT) {
        const int k = s_ids[i];
        int r = 0;
        for (int j = 0; j < nids; ++j) r += (s_ids[j] < k);
        s_sorted[r] = k;
    }

    const bool single_chunk = nids <= kChk;
    for (int sp = 0; sp < NS; ++sp) {
    // ---- this lane's pixels ----------------------------------------------
    const int sx0 = tx0 + (warp & 1) * 16, sy0 = ty0 + sp * TH + (warp >> 1) * RW;   // strip origin
    const int y = sy0 + lane / LPR;
    const int xb = sx0 + (lane % LPR) * PX;
    const bool row_ok = y < H;
    u64 px2[NP][CP];
    float px1[CP];
    unsigned vmask = 0;   // bit j: pixel j is inside the raster and the mask
#pragma unroll
    for (int j = 0; j < PX; ++j) {
        bool v = row_ok && (xb + j) < W;
        if (v && mask) v = mask[(int64_t)y * W + xb + j] != 0;
        vmask |= (v ? 1u : 0u) << j;
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
    int bests[PX];
#pragma unroll
    for (int j = 0; j < PX; ++j) {
        best[j] = INF;
        bests[j] = -1;
    }
    // warp strip (inclusive), clipped to the image
    const int wx0 = sx0, wx1 = min(sx0 + 16, W) - 1;
    const int wy0 = sy0 + y_off, wy1 = min(sy0 + RW, H) - 1 + y_off;
    const int yg = y + y_off;   // global row
    const float fy = (float)yg;
    float wbound = INF;  // upper bound of every strip pixel's final distance (pruning bound)

    // ---- evaluate candidates chunk by chunk ------------------------------------
    for (int c0 = 0; c0 < nids; c0 += kChk) {
        const int nc = min(kChk, nids - c0);
        // the centre records are loaded once when they all fit (the usual case)
        const bool load_chunk = !(single_chunk && sp > 0);
        if (load_chunk) __syncthreads();  // previous chunk fully consumed; s_sorted complete
        if (load_chunk) {
            // positions and windows: one thread per candidate
            for (int sI = tid; sI < nc; sI += NT) {
                const int k = s_sorted[c0 + sI];
                const float *rec = centres + (int64_t)k * (2 + Cf);
                const float cy = rec[0], cx = rec[1];
                s_cyx[sI] = make_float2(cy, cx);
                // windows exactly as the reference computes them (float32, then C cast)
                const float ylo = __fsub_rn(cy, (float)(2 * step_y));
                const float yhi = __fadd_rn(__fadd_rn(cy, (float)(2 * step_y)), 1.0f);
                const float xlo = __fsub_rn(cx, (float)(2 * step_x));
                const float xhi = __fadd_rn(__fadd_rn(cx, (float)(2 * step_x)), 1.0f);
                int4 w;
                w.x = trunc_i((0.0f > ylo) ? 0.0f : ylo);
                w.y = trunc_i(((float)Hg < yhi) ? (float)Hg : yhi);
                w.z = trunc_i((0.0f > xlo) ? 0.0f : xlo);
                w.w = trunc_i(((float)W < xhi) ? (float)W : xhi);
                s_win[sI] = w;
                if (SZ) s_mdc[sI] = maxdc[k];
            }
            // colours: batches of independent loads (the address depends on a shared-memory
            // look-up, so the compiler does not overlap the round trips by itself)
            constexpr int UN = (CP >= 32) ? 8 : 2;
            for (int i0 = tid; i0 < nc * CP; i0 += NT * UN) {
                float v[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const int i = i0 + u * NT;
                    v[u] = 0.0f;
                    if (i < nc * CP) {
                        const int sI = i / CP, c = i % CP;
                        if (c < Cf) v[u] = centres[(int64_t)s_sorted[c0 + sI] * (2 + Cf) + 2 + c];
                    }
                }
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const int i = i0 + u * NT;
                    if (i < nc * CP) (&s_nf[0][0])[i] = -v[u];
                }
            }
        }
        if (load_chunk) __syncthreads();

        // Each lane owns kChk/32 candidates: does the window touch the strip, and what is a
        // lower bound of the spatial term over the strip (deflated so rounding can never lift
        // it above a true per-pixel value).
        bool hit[kChk / 32];
        float lb[kChk / 32];
        unsigned seedkey = 0xffffffffu;
#pragma unroll
        for (int half = 0; half < kChk / 32; ++half) {
            const int sc = half * 32 + lane;
            hit[half] = false;
            lb[half] = INF;
            if (sc < nc) {
                const int4 w = s_win[sc];
                if (!(w.x > wy1 || w.y <= wy0 || w.z > wx1 || w.w <= wx0)) {
                    const float2 c = s_cyx[sc];
                    const float ddy = fmaxf(0.0f, fmaxf((float)wy0 - c.x, c.x - (float)wy1));
                    const float ddx = fmaxf(0.0f, fmaxf((float)wx0 - c.y, c.y - (float)wx1));
                    hit[half] = true;
                    lb[half] = (sp_y * sp_y * (ddy * ddy) + sp_x * sp_x * (ddx * ddx)) * spatial_weight * 0.9999f;
                    seedkey = min(seedkey, (__float_as_uint(lb[half]) & ~63u) | (unsigned)sc);
                }
            }
        }
        // Seed the pruning bound with the nearest candidate: its distances bound every pixel's
        // final minimum from above.  (It is evaluated again at its own position below, so the
        // ascending-order tie rule is untouched.)
        seedkey = __reduce_min_sync(0xffffffffu, seedkey);
        if (seedkey != 0xffffffffu && __uint_as_float(seedkey & ~63u) < wbound) {
            const int s = (int)(seedkey & 63u);
            float tb[PX];
            int ts[PX];
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                tb[j] = ((vmask >> j) & 1u) ? INF : 0.0f;   // pixels outside raster/mask never bound
                ts[j] = -1;
            }
            const float2 c = s_cyx[s];
            eval_candidate<CP, PX, true, SZ>(px2, px1, nx2, nx1, fy, c.x, c.y, s_win[s], s_nf[s],
                                             SZ ? s_mdc[s] : 1.0f, spatial_weight, sp_y, sp_x, ignore_color, yg, xb, 0, tb, ts);
            float m = tb[0];
#pragma unroll
            for (int j = 1; j < PX; ++j) m = fmaxf(m, tb[j]);
            wbound = fminf(wbound, __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(m))));
        }
        // Candidates whose bound exceeds wbound cannot beat any pixel's current or final best
        // (d >= spatial term >= bound > best): skip them.  Survivors go in ascending slot order.
#pragma unroll
        for (int half = 0; half < kChk / 32; ++half) {
            unsigned m = __ballot_sync(0xffffffffu, hit[half] && lb[half] <= wbound);
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                const int s = half * 32 + b;
                const float2 c = s_cyx[s];
                const int4 w = s_win[s];
                const bool full = w.x <= wy0 && w.y > wy1 && w.z <= wx0 && w.w > wx1;
                const float mdc = SZ ? s_mdc[s] : 1.0f;
                if (full)
                    eval_candidate<CP, PX, false, SZ>(px2, px1, nx2, nx1, fy, c.x, c.y, w, s_nf[s], mdc,
                                                      spatial_weight, sp_y, sp_x, ignore_color, yg, xb, c0 + s, best, bests);
                else
                    eval_candidate<CP, PX, true, SZ>(px2, px1, nx2, nx1, fy, c.x, c.y, w, s_nf[s], mdc,
                                                     spatial_weight, sp_y, sp_x, ignore_color, yg, xb, c0 + s, best, bests);
            }
        }
    }

    // ---- labels ------------------------------------------------------------
    int kk[PX];
    bool all_found = true;
#pragma unroll
    for (int j = 0; j < PX; ++j) {
        const bool v = (vmask >> j) & 1u;
        kk[j] = (v && bests[j] >= 0) ? s_sorted[bests[j]] : -1;
        all_found = all_found && v && bests[j] >= 0;
    }
    if (PX == 4 && all_found && (W & 3) == 0) {
        *reinterpret_cast<int4 *>(labels + (int64_t)y * W + xb) =
            make_int4(kk[0] + start_label, kk[1 % PX] + start_label, kk[2 % PX] + start_label,
                      kk[3 % PX] + start_label);
    } else {
#pragma unroll
        for (int j = 0; j < PX; ++j)
            if (kk[j] >= 0) labels[(int64_t)y * W + xb + j] = kk[j] + start_label;
    }

    // ---- fused centre update ---------------------------------------------------
    // Per lane: the pixels that share the lane's leading winner are summed in registers (fixed
    // order) into one record, every other pixel becomes its own record.  Records are 32-bit
    // fixed point in shared memory; the tile then folds them into its per-slot accumulators
    // field-parallel (integer adds: independent of scheduling) and issues one RED.64 per touched
    // (centre, field).
    auto px_val = [&](int j, int c) -> float {
        if constexpr (PX == 1) {
            return px1[c];
        } else {
            float lo, hi;
            unpack2(px2[(j / 2) % NP][c], lo, hi);
            return (j & 1) ? hi : lo;
        }
    };
    int lead = -1;   // leading winner slot of this lane
#pragma unroll
    for (int j = PX - 1; j >= 0; --j)
        if (((vmask >> j) & 1u) && bests[j] >= 0) lead = bests[j];
    // Contributions ("items") of this lane: item PX = the pixels that share the leading winner, summed
    // in a fixed order; item j = pixel j on its own (another winner, or it kept its previous centre).
    // place(): into the record pool if there is room, or straight to HBM when the item's centre has no
    // tile accumulator; returns false when the pool is full and the item has to wait for the next round.
    auto place = [&](int item, int kcur) -> bool {
        int slot, cnt = 0, sxl = 0;
        float fs[CP];
        if (item == PX) {
            slot = lead;
#pragma unroll
            for (int c = 0; c < CP; ++c) fs[c] = 0.0f;
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                if (((vmask >> j) & 1u) && bests[j] == lead) {
                    cnt += 1;
                    sxl += xb + j - tx0;
#pragma unroll
                    for (int c = 0; c < CP; ++c) fs[c] = __fadd_rn(fs[c], px_val(j, c));
                }
            }
        } else {
            slot = bests[item];
            cnt = 1;
            sxl = xb + item - tx0;
#pragma unroll
            for (int c = 0; c < CP; ++c) fs[c] = px_val(item, c);
        }
        if (slot >= 0 && slot < kAcc) {
            const int ridx = atomicAdd(&s_nrec, 1);
            if (ridx < kRec) {
                int *r = s_rec[ridx];
                r[0] = slot;
                r[1] = cnt;
                r[2] = cnt * (y - ty0);
                r[3] = sxl;
#pragma unroll
                for (int c = 0; c < CP; ++c) r[4 + c] = __float2int_rn(fs[c] * fix_scale32);
                return true;
            }
            if (kRounds) return false;            // pool full: next round
        }
        // no tile accumulator for this centre (or, with few channels, the rare full pool): add to
        // HBM directly, with the SAME 32-bit quantisation as the tile path, so the sum does not
        // depend on which contributions took this route
        unsigned long long *a = acc + (int64_t)kcur * (3 + Cf);
        atomicAdd(&a[0], (unsigned long long)cnt);
        atomicAdd(&a[1], (unsigned long long)((long long)cnt * yg));
        atomicAdd(&a[2], (unsigned long long)((long long)sxl + (long long)cnt * tx0));
#pragma unroll
        for (int c = 0; c < CP; ++c)
            if (c < Cf)
                atomicAdd(&a[3 + c], (unsigned long long)((long long)__float2int_rn(fs[c] * fix_scale32) * fix_ratio));
        return true;
    };
    auto fold = [&]() {
        const int nrec = min(s_nrec, kRec);
        for (int e = tid; e < nrec * NF; e += NT) {
            const int r = e / NF, f = e - r * NF;
            const int v = s_rec[r][1 + f];
            if (v != 0) atomicAdd(&s_acc[s_rec[r][0]][f], v);
        }
    };
    if constexpr (!kRounds) {
        // few channels: the pool holds a whole strip phase
        if (lead >= 0) place(PX, s_sorted[lead]);
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            if (!((vmask >> j) & 1u) || bests[j] == lead) continue;
            int kcur = kk[j];
            if (kcur < 0) kcur = labels[(int64_t)y * W + xb + j] - start_label;   // kept its previous centre
            if (kcur < 0) continue;
            place(j, kcur);
        }
        __syncthreads();
        fold();
        __syncthreads();
        if (tid == 0) s_nrec = 0;
        __syncthreads();
    } else {
        // many channels: the pool (shared memory) is small against a strip phase and a direct RED.64
        // per field would dominate the kernel -> fill the pool, fold it, repeat while items are left
        unsigned pending = (lead >= 0) ? (1u << PX) : 0u;
        int kc[PX];
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            kc[j] = -1;
            if (!((vmask >> j) & 1u) || bests[j] == lead) continue;
            kc[j] = kk[j];
            if (kc[j] < 0) kc[j] = labels[(int64_t)y * W + xb + j] - start_label;
            if (kc[j] >= 0) pending |= 1u << j;
        }
        while (true) {
            if (((pending >> PX) & 1u) && place(PX, s_sorted[lead])) pending &= ~(1u << PX);
#pragma unroll
            for (int j = 0; j < PX; ++j)
                if (((pending >> j) & 1u) && place(j, kc[j])) pending &= ~(1u << j);
            const int more = __syncthreads_or(pending != 0);
            fold();
            __syncthreads();
            if (tid == 0) s_nrec = 0;
            __syncthreads();
            if (!more) break;
        }
    }
    }   // strip phases
    const int nslots = min(nids, kAcc);
    for (int i = tid; i < nslots * (3 + Cf); i += NT) {
        const int slot = i / (3 + Cf), f = i % (3 + Cf);
        const int cnt = s_acc[slot][0];
        if (cnt == 0) continue;
        const long long v = s_acc[slot][f];
        long long g;
        if (f == 0) g = v;
        else if (f == 1) g = v + (long long)cnt * (ty0 + y_off);
        else if (f == 2) g = v + (long long)cnt * tx0;
        else g = v * fix_ratio;
        if (g != 0) atomicAdd(&acc[(int64_t)s_sorted[slot] * (3 + Cf) + f], (unsigned long long)g);
    }
}

template <int CP, int PX, int NS>
static int launch_assign(const float *feat, const uint8_t *mask, const float *centres, int slic_zero,
                         const SlicWs &w,
                         int32_t *labels, int64_t H, int64_t W, int64_t pitch, int Cf, float sw,
                         int step_y, int step_x, int start_label, int ignore_color, double fix_scale,
                         int32_t *status, int y_off, int64_t Hg, cudaStream_t st)
{
    // 32-bit fixed point for the per-tile shared-memory sums.  fix_scale obeys
    //   max|feature| * fix_scale * reach <= 2^62,  reach = min(H*W, (4*step_y+1)*(4*step_x+1)),
    // so with bits_px = ceil(log2(reach+1)):  max|feature| * (fix_scale * 2^(bits_px-42)) <= 2^20,
    // and a tile of <= 1024 * NS pixels stays below 2^30 once scaled down by NS.
    const int64_t reach = std::min<int64_t>(Hg * W, (int64_t)(4 * step_y + 1) * (4 * step_x + 1));
    int bits_px = 1;
    while ((1LL << bits_px) < reach + 1) ++bits_px;
    int lg_ns = 0;
    while ((1 << lg_ns) < NS) ++lg_ns;
    const float fix_scale32 = (float)ldexp(fix_scale, bits_px - 42 - lg_ns);
    const long long fix_ratio = 1LL << (42 - bits_px + lg_ns);
    dim3 grid((unsigned)ceil_div(W, 32), (unsigned)ceil_div(H, NS * 4 * (32 / (16 / PX))));
    constexpr size_t dyn = Traits<CP>::kDynRec ? (size_t)Traits<CP>::kRec * (3 + CP + 1) * sizeof(int) : 0;
    if (dyn > 0) {
        OBIA_CUDA_CHECK(cudaFuncSetAttribute(slic_assign_update_kernel<CP, PX, NS, true>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        OBIA_CUDA_CHECK(cudaFuncSetAttribute(slic_assign_update_kernel<CP, PX, NS, false>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    }
    prof_begin(st);
    if (slic_zero)
        slic_assign_update_kernel<CP, PX, NS, true><<<grid, kWarps * 32, dyn, st>>>(
            feat, mask, centres, w.maxdc, w.head, w.next, labels, w.acc, (int)H, (int)W, pitch, Cf, sw, step_y, step_x,
            (int)w.ncy, (int)w.ncx, start_label, ignore_color, fix_scale, fix_scale32, fix_ratio, status, y_off, (int)Hg,
            w.sp_y, w.sp_x);
    else
        slic_assign_update_kernel<CP, PX, NS, false><<<grid, kWarps * 32, dyn, st>>>(
            feat, mask, centres, w.maxdc, w.head, w.next, labels, w.acc, (int)H, (int)W, pitch, Cf, sw, step_y, step_x,
            (int)w.ncy, (int)w.ncx, start_label, ignore_color, fix_scale, fix_scale32, fix_ratio, status, y_off, (int)Hg,
            w.sp_y, w.sp_x);
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

// SLICO (slic_zero): after the centres of a sweep are known, every pixel's colour distance to its
// own (new) centre raises that centre's running maximum (_slic.pyx "update the color distance
// maxima").  Distances are non-negative floats, so an unsigned atomicMax on the bit pattern is the
// float maximum; max is order-free, hence deterministic.
__global__ void __launch_bounds__(256)
slic_max_color_kernel(const float *__restrict__ feat, const uint8_t *__restrict__ mask,
                      const float *__restrict__ centres, const int32_t *__restrict__ labels,
                      float *maxdc, int H, int W, int64_t pitch, int Cf, int64_t n, int start_label)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)H * W) return;
    if (mask && !mask[i]) return;
    const int k = labels[i] - start_label;
    if (k < 0 || k >= n) return;
    const int y = (int)(i / W), x = (int)(i - (int64_t)y * W);
    const float *c = centres + (int64_t)k * (2 + Cf) + 2;
    const float *f = feat + (int64_t)y * pitch + x;
    float acc = 0.0f;
    for (int ch = 0; ch < Cf; ++ch) {
        const float t = __fsub_rn(f[(int64_t)ch * H * pitch], c[ch]);
        acc = __fmaf_rn(t, t, acc);   // same contraction as the assignment kernel
    }
    if (acc == acc && acc > maxdc[k]) atomicMax(reinterpret_cast<unsigned *>(maxdc + k), __float_as_uint(acc));
}

// Sharded runs that exchange only the boundary bands of the centre sums: a live centre outside the
// bands whose window reaches beyond this rank's rows would need sums from (or be a candidate on) the
// neighbour -> status[1] = 1, the caller falls back to the full all-reduce.
__global__ void __launch_bounds__(256)
slic_band_check_kernel(const float *__restrict__ centres, int64_t n, int Cf, int row_lo, int row_hi, int H_total,
                       int64_t up_lo, int64_t up_hi, int64_t down_lo, int64_t down_hi, int step_y, int32_t *status)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float cy = centres[k * (2 + Cf)];
    if (!(cy == cy)) return;
    const float reach = (float)(2 * step_y + 3);
    const bool in_up = k >= up_lo && k < up_hi, in_down = k >= down_lo && k < down_hi;
    if (row_lo > 0 && !in_up && cy - reach < (float)row_lo) atomicExch(status + 1, 1);
    if (row_hi < H_total && !in_down && cy + reach > (float)row_hi) atomicExch(status + 1, 1);
}

__global__ void fill_f32_kernel(float *p, int64_t n, float v)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace obia

namespace obia {
// slic_fast.cu
int launch_assign_fast(const float *feat, const uint8_t *mask, const float *centres, const SlicWs &w, int32_t *labels,
                       int64_t H, int64_t W, int64_t pitch, int Cf, float sw, int step_y, int step_x, int start_label,
                       int ignore_color, double fix_scale, int32_t *status, int y_off, int64_t Hg, cudaStream_t st);
int launch_assign_fast_batch(const float *feat, const uint8_t *mask, const float *centres, const int32_t *head,
                             const int32_t *next, unsigned long long *acc, int32_t *labels, const WinDesc *batch,
                             int64_t B, int hmax, int wmax, int64_t slab_rows, int LW, int64_t pitch, int Cf,
                             int start_label, int ignore_color, int32_t *status, int ns_mask, cudaStream_t st);
void fast_batch_fix_params(WinDesc *descs_host, int64_t B, int Cf);
}

using namespace obia;

extern "C" int64_t obia_b200_slic_workspace_bytes(int64_t H, int64_t W, int32_t Cf, int64_t n,
                                                  int32_t step_y, int32_t step_x)
{
    if (H <= 0 || W <= 0 || Cf <= 0 || n <= 0 || step_y <= 0 || step_x <= 0) return -1;
    return slic_ws_layout(nullptr, H, W, Cf, n, step_y, step_x).bytes;
}

static int slic_check_args(const void *features, const void *centres, const void *labels, const void *workspace,
                           const void *status, int64_t H, int64_t W, int64_t pitch, int32_t Cf, int64_t n,
                           float step, int32_t step_y, int32_t step_x, int32_t start_label, double fix_scale)
{
    if (!features || !centres || !labels || !workspace || !status || H <= 0 || W <= 0 || Cf <= 0 || n <= 0 ||
        step_y <= 0 || step_x <= 0 || !(step > 0.0f) || !(fix_scale > 0.0))
        return set_err(OBIA_B200_ERR_ARG, "slic: bad argument");
    if (H > 2000000000 || W > 2000000000 || n > 2000000000)
        return set_err(OBIA_B200_ERR_UNSUPPORTED, "slic: dimension exceeds int32");
    if (pitch < W || (pitch & 3)) return set_err(OBIA_B200_ERR_ARG, "slic: pitch must be a multiple of 4 and >= W");
    if ((reinterpret_cast<uintptr_t>(features) & 15) || (reinterpret_cast<uintptr_t>(labels) & 15) ||
        (reinterpret_cast<uintptr_t>(workspace) & 15))
        return set_err(OBIA_B200_ERR_ARG, "slic: features, labels and workspace must be 16-byte aligned");
    if (Cf > 64) return set_err(OBIA_B200_ERR_UNSUPPORTED, "slic: more than 64 feature channels");
    if (start_label != 0 && start_label != 1) return set_err(OBIA_B200_ERR_ARG, "start_label should be 0 or 1.");
    return OBIA_B200_OK;
}

extern "C" int obia_b200_slic_begin(int32_t *labels, void *workspace, int64_t H, int64_t W, int64_t H_total,
                                    int32_t Cf, int64_t n, int32_t step_y, int32_t step_x, int32_t start_label,
                                    int32_t *status, void *stream)
{
    if (!labels || !workspace || !status || H <= 0 || W <= 0 || H_total < H || Cf <= 0 || n <= 0)
        return set_err(OBIA_B200_ERR_ARG, "slic_begin: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    SlicWs w = slic_ws_layout(workspace, H_total, W, Cf, n, step_y, step_x);
    OBIA_CUDA_CHECK(cudaMemsetAsync(status, 0, 4 * sizeof(int32_t), st));
    OBIA_CUDA_CHECK(cudaMemsetAsync(w.acc, 0, (size_t)n * (3 + Cf) * 8, st));
    fill_i32_kernel<<<kNumSMs * 4, 256, 0, st>>>(labels, H * W, start_label - 1);
    OBIA_LAUNCH_CHECK();
    fill_f32_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(w.maxdc, n, 1.0f);   // np.ones(n_segments)
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

static int slic_sweep_impl(const float *features, const uint8_t *mask, const float *centres,
                           int32_t *labels, void *workspace, int64_t H, int64_t W, int64_t pitch,
                           int32_t Cf, int64_t n, float step, int32_t step_y, int32_t step_x,
                           int32_t start_label, int32_t ignore_color, int32_t slic_zero,
                           double fix_scale, int64_t y_offset, int64_t H_total, int32_t *status,
                           void *stream, int fast, float sp_y = 1.0f, float sp_x = 1.0f)
{
    int rc = slic_check_args(features, centres, labels, workspace, status, H, W, pitch, Cf, n, step, step_y, step_x,
                             start_label, fix_scale);
    if (rc) return rc;
    if (y_offset < 0 || y_offset + H > H_total) return set_err(OBIA_B200_ERR_ARG, "slic_sweep: strip outside the raster");
    cudaStream_t st = (cudaStream_t)stream;
    SlicWs w = slic_ws_layout(workspace, H_total, W, Cf, n, step_y, step_x);
    if (!(sp_y > 0.0f) || !(sp_x > 0.0f) || sp_y == INFINITY || sp_x == INFINITY)
        return set_err(OBIA_B200_ERR_ARG, "slic_sweep: spacing must be positive and finite");
    if ((sp_y != 1.0f || sp_x != 1.0f) && fast && !slic_zero)
        return set_err(OBIA_B200_ERR_UNSUPPORTED, "slic_sweep: anisotropic spacing runs on the exact kernel only");
    w.sp_y = sp_y;
    w.sp_x = sp_x;
    // `1.0 / (step * step)`: float product, double division, float store
    const float step_sq = step * step;
    const float sw = (float)(1.0 / (double)step_sq);
    OBIA_CUDA_CHECK(cudaMemsetAsync(w.head, 0xff, (size_t)(w.ncy * w.ncx) * 4, st));
    slic_centres_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(
        const_cast<float *>(centres), w.acc, w.head, w.next, n, Cf, 0, 1.0 / fix_scale, step_y, step_x, w.ncy, w.ncx);
    OBIA_LAUNCH_CHECK();
    const int yo = (int)y_offset;
    // tolerance mode (slic_fast.cu); SLICO divides the colour term per centre and stays on the exact kernel
    if (fast && !slic_zero)
        return launch_assign_fast(features, mask, centres, w, labels, H, W, pitch, Cf, sw, step_y, step_x, start_label,
                                  ignore_color, fix_scale, status, yo, H_total, st);
    if (Cf <= 4)
        rc = launch_assign<4, 4, OBIA_NS>(features, mask, centres, slic_zero, w, labels, H, W, pitch, Cf, sw, step_y, step_x,
                                          start_label, ignore_color, fix_scale, status, yo, H_total, st);
    else if (Cf <= 8)
        rc = launch_assign<8, 4, OBIA_NS>(features, mask, centres, slic_zero, w, labels, H, W, pitch, Cf, sw, step_y, step_x,
                                          start_label, ignore_color, fix_scale, status, yo, H_total, st);
    else if (Cf <= 16)
        rc = launch_assign<16, 2, OBIA_NS>(features, mask, centres, slic_zero, w, labels, H, W, pitch, Cf, sw, step_y, step_x,
                                           start_label, ignore_color, fix_scale, status, yo, H_total, st);
    else if (Cf <= 32)
        rc = launch_assign<32, 2, OBIA_NS>(features, mask, centres, slic_zero, w, labels, H, W, pitch, Cf, sw, step_y, step_x,
                                           start_label, ignore_color, fix_scale, status, yo, H_total, st);
    else
        rc = launch_assign<64, 2, OBIA_NS>(features, mask, centres, slic_zero, w, labels, H, W, pitch, Cf, sw, step_y, step_x,
                                           start_label, ignore_color, fix_scale, status, yo, H_total, st);
    return rc;
}

extern "C" int obia_b200_slic_sweep(const float *features, const uint8_t *mask, const float *centres,
                                    int32_t *labels, void *workspace, int64_t H, int64_t W, int64_t pitch,
                                    int32_t Cf, int64_t n, float step, int32_t step_y, int32_t step_x,
                                    int32_t start_label, int32_t ignore_color, int32_t slic_zero,
                                    double fix_scale, int64_t y_offset, int64_t H_total, int32_t *status,
                                    void *stream)
{
    return slic_sweep_impl(features, mask, centres, labels, workspace, H, W, pitch, Cf, n, step, step_y, step_x,
                           start_label, ignore_color, slic_zero, fix_scale, y_offset, H_total, status, stream, 0);
}

extern "C" int obia_b200_slic_sweep_fast(const float *features, const uint8_t *mask, const float *centres,
                                         int32_t *labels, void *workspace, int64_t H, int64_t W, int64_t pitch,
                                         int32_t Cf, int64_t n, float step, int32_t step_y, int32_t step_x,
                                         int32_t start_label, int32_t ignore_color, int32_t slic_zero,
                                         double fix_scale, int64_t y_offset, int64_t H_total, int32_t *status,
                                         void *stream)
{
    return slic_sweep_impl(features, mask, centres, labels, workspace, H, W, pitch, Cf, n, step, step_y, step_x,
                           start_label, ignore_color, slic_zero, fix_scale, y_offset, H_total, status, stream, 1);
}

extern "C" int obia_b200_slic_finish_sweep(float *centres, void *workspace, int64_t H_total, int64_t W, int32_t Cf,
                                           int64_t n, int32_t step_y, int32_t step_x, double fix_scale,
                                           void *stream)
{
    if (!centres || !workspace || H_total <= 0 || W <= 0 || Cf <= 0 || n <= 0 || !(fix_scale > 0.0))
        return set_err(OBIA_B200_ERR_ARG, "slic_finish_sweep: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    SlicWs w = slic_ws_layout(workspace, H_total, W, Cf, n, step_y, step_x);
    // means of the sweep's assignment; also zeroes the sums for the next sweep
    OBIA_CUDA_CHECK(cudaMemsetAsync(w.head, 0xff, (size_t)(w.ncy * w.ncx) * 4, st));
    slic_centres_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(centres, w.acc, w.head, w.next, n, Cf, 1,
                                                                   1.0 / fix_scale, step_y, step_x, w.ncy, w.ncx);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

extern "C" int obia_b200_slic_band_check(const float *centres, int64_t n, int32_t Cf, int64_t row_lo, int64_t row_hi,
                                         int64_t H_total, int64_t up_lo, int64_t up_hi, int64_t down_lo,
                                         int64_t down_hi, int32_t step_y, int32_t *status, void *stream)
{
    if (!centres || !status || n <= 0 || Cf <= 0 || row_lo < 0 || row_hi <= row_lo || row_hi > H_total || step_y <= 0)
        return set_err(OBIA_B200_ERR_ARG, "slic_band_check: bad argument");
    slic_band_check_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
        centres, n, Cf, (int)row_lo, (int)row_hi, (int)H_total, up_lo, up_hi, down_lo, down_hi, step_y, status);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

extern "C" int obia_b200_slic_update_max_color(const float *features, const uint8_t *mask, const float *centres,
                                               const int32_t *labels, void *workspace, int64_t H, int64_t W,
                                               int64_t H_total, int64_t pitch, int32_t Cf, int64_t n,
                                               int32_t step_y, int32_t step_x, int32_t start_label, void *stream)
{
    if (!features || !centres || !labels || !workspace || H <= 0 || W <= 0 || H_total < H || Cf <= 0 || n <= 0 ||
        pitch < W)
        return set_err(OBIA_B200_ERR_ARG, "slic_update_max_color: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    SlicWs w = slic_ws_layout(workspace, H_total, W, Cf, n, step_y, step_x);
    slic_max_color_kernel<<<(unsigned)ceil_div(H * W, 256), 256, 0, st>>>(features, mask, centres, labels, w.maxdc,
                                                                         (int)H, (int)W, pitch, Cf, n, start_label);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

static int slic_iterate_impl(const float *features, const uint8_t *mask, float *centres,
                             int32_t *labels, void *workspace, int64_t H, int64_t W,
                             int64_t pitch, int32_t Cf, int64_t n, float step, int32_t step_y,
                             int32_t step_x, int32_t max_num_iter, int32_t start_label,
                             int32_t ignore_color, int32_t slic_zero, double fix_scale,
                             int32_t *status, void *stream, int fast, float sp_y = 1.0f, float sp_x = 1.0f)
{
    int rc = slic_check_args(features, centres, labels, workspace, status, H, W, pitch, Cf, n, step, step_y, step_x,
                             start_label, fix_scale);
    if (rc) return rc;
    if (max_num_iter < 0) return set_err(OBIA_B200_ERR_ARG, "slic_iterate: bad argument");
    rc = obia_b200_slic_begin(labels, workspace, H, W, H, Cf, n, step_y, step_x, start_label, status, stream);
    for (int it = 0; it < max_num_iter && !rc; ++it) {
        rc = slic_sweep_impl(features, mask, centres, labels, workspace, H, W, pitch, Cf, n, step, step_y,
                             step_x, start_label, ignore_color, slic_zero, fix_scale, 0, H, status, stream, fast, sp_y, sp_x);
        // the reference updates the centres after every sweep, the last one included
        if (!rc) rc = obia_b200_slic_finish_sweep(centres, workspace, H, W, Cf, n, step_y, step_x, fix_scale, stream);
        if (!rc && slic_zero)
            rc = obia_b200_slic_update_max_color(features, mask, centres, labels, workspace, H, W, H, pitch, Cf, n,
                                                 step_y, step_x, start_label, stream);
    }
    return rc;
}

extern "C" int obia_b200_slic_iterate(const float *features, const uint8_t *mask, float *centres,
                                      int32_t *labels, void *workspace, int64_t H, int64_t W,
                                      int64_t pitch, int32_t Cf, int64_t n, float step, int32_t step_y,
                                      int32_t step_x, int32_t max_num_iter, int32_t start_label,
                                      int32_t ignore_color, int32_t slic_zero, double fix_scale,
                                      int32_t *status, void *stream)
{
    return slic_iterate_impl(features, mask, centres, labels, workspace, H, W, pitch, Cf, n, step, step_y, step_x,
                             max_num_iter, start_label, ignore_color, slic_zero, fix_scale, status, stream, 0);
}

extern "C" int obia_b200_slic_iterate_spacing(const float *features, const uint8_t *mask, float *centres,
                                              int32_t *labels, void *workspace, int64_t H, int64_t W,
                                              int64_t pitch, int32_t Cf, int64_t n, float step, int32_t step_y,
                                              int32_t step_x, int32_t max_num_iter, int32_t start_label,
                                              int32_t ignore_color, int32_t slic_zero, double fix_scale,
                                              float spacing_y, float spacing_x, int32_t *status, void *stream)
{
    return slic_iterate_impl(features, mask, centres, labels, workspace, H, W, pitch, Cf, n, step, step_y, step_x,
                             max_num_iter, start_label, ignore_color, slic_zero, fix_scale, status, stream, 0,
                             spacing_y, spacing_x);
}

extern "C" int obia_b200_slic_iterate_fast(const float *features, const uint8_t *mask, float *centres,
                                           int32_t *labels, void *workspace, int64_t H, int64_t W,
                                           int64_t pitch, int32_t Cf, int64_t n, float step, int32_t step_y,
                                           int32_t step_x, int32_t max_num_iter, int32_t start_label,
                                           int32_t ignore_color, int32_t slic_zero, double fix_scale,
                                           int32_t *status, void *stream)
{
    return slic_iterate_impl(features, mask, centres, labels, workspace, H, W, pitch, Cf, n, step, step_y, step_x,
                             max_num_iter, start_label, ignore_color, slic_zero, fix_scale, status, stream, 1);
}

// ---- batched iterations (tiled driver; batch.cuh) ---------------------------------------------------
extern "C" int64_t obia_b200_slic_batch_workspace_bytes(int64_t n_total, int64_t cells_total, int32_t Cf)
{
    if (n_total <= 0 || cells_total <= 0 || Cf <= 0) return -1;
    return round_up(n_total * (3 + Cf) * 8, 256) + round_up(cells_total * 4, 256) + round_up(n_total * 4, 256);
}

// fills the kernel-variant dependent fixed-point fields of HOST descriptors (before they are uploaded)
extern "C" int obia_b200_slic_batch_prepare(void *descs_host, int64_t B, int32_t Cf)
{
    if (!descs_host || B <= 0 || Cf <= 0 || Cf > 64) return set_err(OBIA_B200_ERR_ARG, "slic_batch_prepare: bad argument");
    fast_batch_fix_params((WinDesc *)descs_host, B, Cf);
    return OBIA_B200_OK;
}

// `max_num_iter` sweeps (tolerance-mode kernel) over every window of the slab: the batched equivalent of
// obia_b200_slic_iterate_fast called once per window.  labels: (slab_rows, slab_w) int32, filled with
// start_label - 1 here; features: (Cf, slab_rows, pitch); centres: (n_total, 2 + Cf); descs / cwin on the device;
// variants = OR over the usable windows of the `pad` field obia_b200_slic_batch_prepare wrote (tile variants present).
extern "C" int obia_b200_slic_iterate_batch(const float *features, const uint8_t *mask, float *centres, int32_t *labels,
                                            void *workspace, const void *descs, const int32_t *cwin, int64_t B,
                                            int64_t n_total, int64_t cells_total, int32_t hmax, int32_t wmax,
                                            int64_t slab_rows, int32_t slab_w, int64_t pitch, int32_t Cf,
                                            int32_t max_num_iter, int32_t start_label, int32_t ignore_color,
                                            int32_t variants, int32_t *status, void *stream)
{
    if (!features || !centres || !labels || !workspace || !descs || !cwin || !status || B <= 0 || n_total <= 0 ||
        cells_total <= 0 || hmax <= 0 || wmax <= 0 || slab_rows <= 0 || slab_w < wmax || pitch < wmax || (pitch & 3) ||
        Cf <= 0 || Cf > 64 || max_num_iter < 0)
        return set_err(OBIA_B200_ERR_ARG, "slic_iterate_batch: bad argument");
    if (slab_rows * (int64_t)slab_w >= 0x7fffffffLL)
        return set_err(OBIA_B200_ERR_UNSUPPORTED, "slic_iterate_batch: slab exceeds int32 pixels");
    if ((reinterpret_cast<uintptr_t>(features) & 15) || (reinterpret_cast<uintptr_t>(labels) & 15) ||
        (reinterpret_cast<uintptr_t>(workspace) & 15))
        return set_err(OBIA_B200_ERR_ARG, "slic_iterate_batch: features, labels and workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    char *p = (char *)workspace;
    unsigned long long *acc = (unsigned long long *)p;
    p += round_up(n_total * (3 + Cf) * 8, 256);
    int32_t *head = (int32_t *)p;
    p += round_up(cells_total * 4, 256);
    int32_t *next = (int32_t *)p;
    const WinDesc *batch = (const WinDesc *)descs;
    OBIA_CUDA_CHECK(cudaMemsetAsync(status, 0, (size_t)(4 + B) * sizeof(int32_t), st));   // [4 + i]: window i overflowed
    OBIA_CUDA_CHECK(cudaMemsetAsync(acc, 0, (size_t)n_total * (3 + Cf) * 8, st));
    fill_i32_kernel<<<kNumSMs * 4, 256, 0, st>>>(labels, slab_rows * slab_w, start_label - 1);
    OBIA_LAUNCH_CHECK();
    int rc = OBIA_B200_OK;
    for (int it = 0; it < max_num_iter && !rc; ++it) {
        if (it == 0) {   // (later sweeps find the centres binned by the update at the end of the previous one)
            OBIA_CUDA_CHECK(cudaMemsetAsync(head, 0xff, (size_t)cells_total * 4, st));
            slic_centres_batch_kernel<<<(unsigned)ceil_div(n_total, 256), 256, 0, st>>>(centres, acc, head, next, n_total,
                                                                                       Cf, 0, cwin, batch);
            OBIA_LAUNCH_CHECK();
        }
        rc = launch_assign_fast_batch(features, mask, centres, head, next, acc, labels, batch, B, hmax, wmax, slab_rows,
                                      slab_w, pitch, Cf, start_label, ignore_color, status, variants, st);
        if (rc) break;
        // the reference updates the centres after every sweep, the last one included
        OBIA_CUDA_CHECK(cudaMemsetAsync(head, 0xff, (size_t)cells_total * 4, st));
        slic_centres_batch_kernel<<<(unsigned)ceil_div(n_total, 256), 256, 0, st>>>(centres, acc, head, next, n_total, Cf, 1,
                                                                                   cwin, batch);
        OBIA_LAUNCH_CHECK();
    }
    return rc;
}

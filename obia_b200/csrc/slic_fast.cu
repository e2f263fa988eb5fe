// K2, tolerance mode: the SLIC assignment + fused centre update with ONE fused multiply-add per
// channel and candidate (replaces skimage `_slic_cython`, reached from
// obia/segmentation/segment_boundaries.py:51; semantics restated in oracle/slic_core.c:82-192).
//
// The exact kernel (slic.cu) evaluates  d = ((cy-y)^2 + (cx-x)^2) * w + sum_c (f_c - m_c)^2  in the
// reference's float32 operation order: a subtract and an FMA per channel.  Here the square is
// expanded,
//     d(p, k) = A_p + B_k - 2 (w y cy + w x cx + sum_c f_c m_c),
// A_p does not depend on the candidate, so argmin_k d = argmin_k (B_k + g_k . q_p) with the
// (2 + Cf)-vector q_p = (y, x, f) of the pixel and g_k = -2 (w cy, w cx, m) of the centre: one FMA per
// channel, half the colour instructions.  To keep float32 cancellation harmless every CTA tile
// works in LOCAL coordinates: rows / columns relative to the tile centre and colours relative to
// the colour o of one of the tile's candidate centres (|f - o| is the local variation, not the
// absolute level), which leaves the distances unchanged.  Window membership (the reference's
// truncated +-2*step windows), ascending-index tie rule, exact spatial pruning, the deterministic
// fixed-point centre update and the strip addressing (y_off / Hg) are those of the exact kernel;
// sums are taken over f - o and the tile adds count * o back in 64-bit fixed point.
// The score is kept in units of 1/w (the colour side is scaled by step^2 instead of the spatial side
// by w): with integer pixel and centre coordinates -- the first sweep, when all centre colours are
// still zero -- the spatial part is exact integer arithmetic, so the exact ties of the regular grid
// (pixels half-way between two centres) go to the lowest centre index exactly as in the reference.
// Results differ from the exact kernel only where two candidates are within float32 rounding of
// each other (tests: >= 99.9 % agreement per sweep, ARI reported).
#include "batch.cuh"
#include "slic_common.cuh"

#ifndef OBIA_K2_LOAD
#define OBIA_K2_LOAD 1      // 1: one running pointer for the feature planes, mask test behind a uniform branch
#endif
#ifndef OBIA_K2_SKIP
#define OBIA_K2_SKIP 0      // 1: groups of 32 candidate records beyond the chunk size are skipped
#endif

namespace obia {

template <int CP, int NW> struct FastTraits {
    static constexpr int CR = (3 + CP + 3) / 4 * 4;          // floats per candidate record
    static constexpr int kIds = 1024;
    static constexpr int kChk = (CP <= 8) ? 128 : (CP <= 32) ? 64 : 32;   // candidate records resident at once (<= 128)
    static constexpr int kAcc = (CP <= 16) ? 128 : 64;       // slots with a tile accumulator row
    static constexpr int PX = (CP <= 8) ? 4 : 2;             // pixels per lane
    static constexpr int kRecW = 32 * PX;                    // records per warp and phase: one per pixel at most
    static constexpr int kRec = NW * kRecW;
    static constexpr int NFR = 2 + CP;                       // words per record / accumulator row
    static constexpr int RS = NFR | 1;                       // record stride in words (odd: conflict-free stores)
    static constexpr size_t kDyn = ((size_t)kRec * RS + (size_t)kAcc * NFR) * sizeof(int);
};

// Score of the PX pixels of this lane against one candidate record (see the header):
//   rec[0] = B_k / w, rec[1] = -2 cy', rec[2] = -2 cx', rec[3 + c] = -2 m'_c / w.
template <int CP, int PX, int CR, bool CHECK>
__device__ __forceinline__ void eval_fast(const float (&pf)[PX][CP], float yr, float xbr, const float *__restrict__ cand,
                                          const int4 w, int y, int xb, int slot, float (&best)[PX],
                                          int (&bests)[PX])
{
    float rec[CR];
#pragma unroll
    for (int q = 0; q < CR / 4; ++q) {
        const float4 v = reinterpret_cast<const float4 *>(cand)[q];
        rec[4 * q] = v.x; rec[4 * q + 1] = v.y; rec[4 * q + 2] = v.z; rec[4 * q + 3] = v.w;
    }
    int lo = 0, span = PX;
    if (CHECK) {
        lo = w.z - xb;
        span = (y >= w.x && y < w.y) ? (w.w - w.z) : 0;
    }
    const float t0 = fmaf(rec[2], xbr, fmaf(rec[1], yr, rec[0]));
#pragma unroll
    for (int j = 0; j < PX; ++j) {
        float s = fmaf(rec[2], (float)j, t0);
#pragma unroll
        for (int c = 0; c < CP; ++c) s = fmaf(rec[3 + c], pf[j][c], s);
        const bool v = !CHECK || (unsigned)(j - lo) < (unsigned)span;
        if (v && s < best[j]) {
            best[j] = s;
            bests[j] = slot;
        }
    }
}

// Tile accumulator / record layout (NFR = 2 + CP 32-bit words, field-major records):
//   word 0 = count | (sum of tile-local rows << 13)     (tile <= 4096 pixels, <= 128 rows: 13 + 19 bits, unsigned)
//   word 1 = sum of tile-local columns                 (records carry their slot in bits 16..)
//   word 2 + c = 32-bit fixed-point sum of (f_c - o_c)
// BATCH: one launch over all windows of a slab (batch.cuh): blockIdx.z selects the window, whose
// descriptor replaces the scalar parameters; its rows, centres, cells and sums sit at the descriptor's
// offsets inside the batch-wide arrays.  `LW` = row stride of labels / mask (W, or the slab width).
template <int CP, int PX, int NS, int NW, bool BATCH>
__global__ void __launch_bounds__(NW * 32, (CP <= 16) ? (24 / NW) : 1)
slic_assign_fast_kernel(const float *__restrict__ feat, const uint8_t *__restrict__ mask,
                        const float *__restrict__ centres, const int32_t *__restrict__ head,
                        const int32_t *__restrict__ next, int32_t *__restrict__ labels,
                        unsigned long long *__restrict__ acc, int H, int W, int64_t pitch, int Cf,
                        float spatial_weight, float inv_weight, int step_y, int step_x, int ncy, int ncx,
                        int start_label, int ignore_color, double fix_scale, float fix_scale32,
                        long long fix_ratio, int32_t *status, int y_off, int Hg, int dbg,
                        const WinDesc *__restrict__ batch, int LW, int64_t plane)
{
    using T = FastTraits<CP, NW>;
    if constexpr (BATCH) {
        const WinDesc *d = batch + blockIdx.z;
        H = d->h;
        W = d->w;
        if (!d->valid || (int)blockIdx.x * 32 >= W) return;
        if (CP <= 8 && d->pad != NS) return;      // (the window is served by the variant with d->pad strip phases)
        {
            constexpr int tile_rows = (NW / 2) * (32 / (16 / PX)) * NS;
            if ((int)blockIdx.y * tile_rows >= H) return;
        }
        const int64_t r0 = d->row0, c0 = d->c0;
        feat += r0 * pitch;
        labels += r0 * LW;
        if (mask) mask += r0 * LW;
        centres += c0 * (2 + Cf);
        next += c0;
        head += d->cell0;
        acc += c0 * (3 + Cf);
        spatial_weight = d->sw;
        inv_weight = d->inv_w;
        step_y = d->step_y;
        step_x = d->step_x;
        ncy = d->ncy;
        ncx = d->ncx;
        fix_scale = d->fix_scale;
        fix_scale32 = d->fix_scale32;
        fix_ratio = d->fix_ratio;
        y_off = 0;
        Hg = H;
    }
    static_assert(PX == T::PX, "pixels per lane");
    constexpr int LPR = 16 / PX;         // lanes per strip row
    constexpr int RW = 32 / LPR;         // rows per warp strip
    constexpr int TH = (NW / 2) * RW;    // rows per phase (two strips side by side)
    constexpr int NFR = T::NFR;
    constexpr int CR = T::CR, kIds = T::kIds, kChk = T::kChk, kAcc = T::kAcc, kRec = T::kRec;
    constexpr int NT = NW * 32;
    static_assert(TH * NS <= 128 && 32 * TH * NS <= 4096 && kChk <= 128, "packed record fields");
    __shared__ int s_sorted[kIds];
    __shared__ int s_nids;
    __shared__ int4 s_win[kChk];
    __shared__ float2 s_cyx[kChk];
    __shared__ __align__(16) float s_cand[kChk][CR];
    __shared__ unsigned short s_wl[NW][kChk];    // per-warp list of surviving candidates (slot | full << 15)
    __shared__ float s_off[CP];
    __shared__ long long s_off64[CP];
    extern __shared__ __align__(16) int s_dyn[];
    constexpr int RS = T::RS;
    static_assert((size_t)T::kRec * RS >= (size_t)kIds, "candidate ids share the record region");
    int *s_ids = s_dyn;                                                          // as collected; dead after the sort
    int(*s_rec)[RS] = reinterpret_cast<int(*)[RS]>(s_dyn);                      // [kRec][RS]
    int(*s_acc)[NFR] = reinterpret_cast<int(*)[NFR]>(s_dyn + RS * kRec);        // [kAcc][NFR]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * 32, ty0 = blockIdx.y * (TH * NS);
    const int tx1 = min(tx0 + 32, W) - 1, ty1 = min(ty0 + TH * NS, H) - 1;   // inclusive
    // tile-local origin (global row / column of the tile centre)
    const int X0 = tx0 + 16, Y0 = ty0 + y_off + (TH * NS) / 2;

    // ---- collect candidate centre ids (as in the exact kernel) --------------------------------
    if (tid == 0) s_nids = 0;
    for (int i = tid; i < kAcc * NFR; i += NT) (&s_acc[0][0])[i] = 0;
    __syncthreads();
    {
        // cell range whose centres can reach the tile: a centre at cy reaches rows y with
        // y - 2s <= cy < y + 1 + 2s; two pixels of slack absorb the float rounding of the cell index
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
        if (tid == 0) atomicExch(&status[BATCH ? blockIdx.z : 0], 1);   // (batch: one word per window)
        nids = kIds;
    }
    for (int i = tid; i < nids; i += NT) {
        const int k = s_ids[i];
        int r = 0;
        for (int j = 0; j < nids; ++j) r += (s_ids[j] < k);
        s_sorted[r] = k;
    }
    __syncthreads();
    // (the colour origin of the tile -- the colour of its middle candidate: live centres only are binned,
    //  so it is finite; 0 in the spatial-only pass and when the tile has no candidate -- is published by
    //  the thread that loads that candidate's record in the first chunk build)
    if (nids == 0 && tid < CP) {
        s_off[tid] = 0.0f;
        s_off64[tid] = 0;
    }

    const bool single_chunk = nids <= kChk;
    const float INF = __int_as_float(0x7f800000);
    for (int sp = 0; sp < NS; ++sp) {
        // ---- this lane's pixels ------------------------------------------------------------
        const int sx0 = tx0 + (warp & 1) * 16, sy0 = ty0 + sp * TH + (warp >> 1) * RW;
        const int y = sy0 + lane / LPR;
        const int xb = sx0 + (lane % LPR) * PX;
        const bool ld_ok = y < H && xb < W;
        float pf[PX][CP];
        unsigned vmask = 0;
#if !OBIA_K2_LOAD
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            bool v = ld_ok && (xb + j) < W;
            if (v && mask) v = mask[(int64_t)y * LW + xb + j] != 0;
            vmask |= (v ? 1u : 0u) << j;
        }
#else
        if (ld_ok) {
            // columns of the lane that exist; then the mask (uniform branch: most runs have none)
            vmask = (xb + PX <= W) ? ((1u << PX) - 1u) : ((1u << (W - xb)) - 1u);
            if (mask) {
                const uint8_t *mrow = mask + (int64_t)y * LW + xb;
#pragma unroll
                for (int j = 0; j < PX; ++j)
                    if (((vmask >> j) & 1u) && mrow[j] == 0) vmask &= ~(1u << j);
            }
        }
#endif
        {
            // one pointer, advanced plane by plane (no 64-bit multiply per channel)
            const float *src = feat + (int64_t)y * pitch + xb;
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ld_ok && c < Cf) {
#if OBIA_K2_LOAD
                    const float *sc_ = src;
#else
                    const float *sc_ = src + c * plane;
#endif
                    if constexpr (PX == 4) {
                        v = ldg_stream_f4(reinterpret_cast<const float4 *>(sc_));
                    } else {
                        const float2 t = *reinterpret_cast<const float2 *>(sc_);
                        v.x = t.x; v.y = t.y;
                    }
                }
#if OBIA_K2_LOAD
                src += plane;
#endif
                // raw features for now: the loads are in flight while the candidate records are built,
                // the tile's colour origin is subtracted right after
                pf[0][c] = v.x;
                pf[1][c] = v.y;
                if constexpr (PX == 4) {
                    pf[2][c] = v.z;
                    pf[3][c] = v.w;
                }
            }
        }
        bool centred = false;
        const int yg = y + y_off;
        const float yr = (float)(yg - Y0), xbr = (float)(xb - X0);

        float best[PX];
        int bests[PX];
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            best[j] = INF;
            bests[j] = -1;
        }
        const int wx0 = sx0, wx1 = min(sx0 + 16, W) - 1;
        const int wy0 = sy0 + y_off, wy1 = min(sy0 + RW, H) - 1 + y_off;
        float wbound = INF;

        for (int c0 = 0; c0 < nids; c0 += kChk) {
            const int nc = min(kChk, nids - c0);
            const bool load_chunk = !(single_chunk && sp > 0);
            if (load_chunk) {
                __syncthreads();
                // stage A: every candidate's record into registers (one global round trip)
                constexpr int PER = (kChk + NT - 1) / NT;
                float rcy[PER], rcx[PER], rm[PER][CP];
#pragma unroll
                for (int u = 0; u < PER; ++u) {
                    const int sI = tid + u * NT;
                    rcy[u] = rcx[u] = 0.0f;
#pragma unroll
                    for (int c = 0; c < CP; ++c) rm[u][c] = 0.0f;
                    if (sI < nc) {
                        const int k = s_sorted[c0 + sI];
                        const float *rec = centres + (int64_t)k * (2 + Cf);
                        rcy[u] = rec[0];
                        rcx[u] = rec[1];
#pragma unroll
                        for (int c = 0; c < CP; ++c) rm[u][c] = (c < Cf && !ignore_color) ? rec[2 + c] : 0.0f;
                        if (c0 == 0 && sI == min(nids >> 1, nc - 1)) {   // the origin candidate sits in the first chunk
#pragma unroll
                            for (int c = 0; c < CP; ++c) {
                                s_off[c] = rm[u][c];
                                s_off64[c] = __double2ll_rn((double)rm[u][c] * fix_scale);
                            }
                        }
                    }
                }
                if (c0 == 0) __syncthreads();     // s_off published
                // stage B: windows and score records relative to the tile origin
#pragma unroll
                for (int u = 0; u < PER; ++u) {
                    const int sI = tid + u * NT;
                    if (sI >= nc) continue;
                    const float cy = rcy[u], cx = rcx[u];
                    s_cyx[sI] = make_float2(cy, cx);
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
                    const float cyr = cy - (float)Y0, cxr = cx - (float)X0;
                    float Bc = 0.0f;
                    float *cr = s_cand[sI];
                    cr[1] = -2.0f * cyr;
                    cr[2] = -2.0f * cxr;
#pragma unroll
                    for (int c = 0; c < CP; ++c) {
                        const float mc = (c < Cf && !ignore_color) ? rm[u][c] - s_off[c] : 0.0f;
                        Bc = fmaf(mc, mc, Bc);
                        cr[3 + c] = -2.0f * inv_weight * mc;
                    }
                    // integer-valued (exact) while the centres sit on the pixel grid and Bc == 0
                    cr[0] = fmaf(cyr, cyr, cxr * cxr) + inv_weight * Bc;
#pragma unroll
                    for (int c = 3 + CP; c < CR; ++c) cr[c] = 0.0f;
                }
                __syncthreads();
            }
            if (!centred) {
                centred = true;
#pragma unroll
                for (int c = 0; c < CP; ++c) {
                    const float o = s_off[c];
#pragma unroll
                    for (int j = 0; j < PX; ++j) pf[j][c] -= o;
                }
            }

            bool hit[kChk / 32];
            float lb[kChk / 32];
            unsigned fullbits[kChk / 32];
            unsigned seedkey = 0xffffffffu;
#pragma unroll
            for (int half = 0; half < kChk / 32; ++half) {
                const int sc = half * 32 + lane;
                hit[half] = false;
                lb[half] = INF;
                fullbits[half] = 0u;
#if OBIA_K2_SKIP
                if (half * 32 >= nc) continue;      // (warp-uniform: no record in this group of 32)
#endif
                bool full = false;
                if (sc < nc) {
                    const int4 w = s_win[sc];
                    if (!(w.x > wy1 || w.y <= wy0 || w.z > wx1 || w.w <= wx0)) {
                        const float2 c = s_cyx[sc];
                        const float ddy = fmaxf(0.0f, fmaxf((float)wy0 - c.x, c.x - (float)wy1));
                        const float ddx = fmaxf(0.0f, fmaxf((float)wx0 - c.y, c.y - (float)wx1));
                        hit[half] = true;
                        lb[half] = (ddy * ddy + ddx * ddx) * spatial_weight * 0.9999f;
                        seedkey = min(seedkey, (__float_as_uint(lb[half]) & ~127u) | (unsigned)sc);
                        full = w.x <= wy0 && w.y > wy1 && w.z <= wx0 && w.w > wx1;
                    }
                }
                fullbits[half] = __ballot_sync(0xffffffffu, full);
            }
            // Seed the pruning bound with the nearest candidate: its distances bound every pixel's final
            // minimum from above.  Skipped when its spatial bound alone says nothing can be pruned
            // (colour-dominated runs: the bound would exceed every candidate's spatial term anyway).
            seedkey = __reduce_min_sync(0xffffffffu, seedkey);
            if (dbg & 4) seedkey = 0xffffffffu;
            if (seedkey != 0xffffffffu && __uint_as_float(seedkey & ~127u) < wbound) {
                const int s = (int)(seedkey & 127u);
                float tb[PX];
                int ts[PX];
#pragma unroll
                for (int j = 0; j < PX; ++j) {
                    tb[j] = ((vmask >> j) & 1u) ? INF : 0.0f;
                    ts[j] = -1;
                }
                eval_fast<CP, PX, CR, true>(pf, yr, xbr, s_cand[s], s_win[s], yg, xb, 0, tb, ts);
                // true distance = w * score + A, A = candidate-independent part
                float m = 0.0f;
#pragma unroll
                for (int j = 0; j < PX; ++j) {
                    const float xr = xbr + (float)j;
                    float a = spatial_weight * (yr * yr + xr * xr);
                    if (!ignore_color) {
#pragma unroll
                        for (int c = 0; c < CP; ++c) a = fmaf(pf[j][c], pf[j][c], a);
                    }
                    if ((vmask >> j) & 1u) m = fmaxf(m, fmaf(tb[j], spatial_weight, a));
                }
                m = fmaxf(m, 0.0f) * 1.0001f + 1e-30f;   // rounding slack of the expanded form
                wbound = fminf(wbound, __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(m))));
            }
            // survivors, ascending slot order, compacted into the warp's list
            int wcnt = 0;
#pragma unroll
            for (int half = 0; half < kChk / 32; ++half) {
#if OBIA_K2_SKIP
                if (half * 32 >= nc) continue;
#endif
                const unsigned m = __ballot_sync(0xffffffffu, hit[half] && lb[half] <= wbound);
                if ((m >> lane) & 1u)
                    s_wl[warp][wcnt + __popc(m & ((1u << lane) - 1u))] =
                        (unsigned short)((half * 32 + lane) | (((fullbits[half] >> lane) & 1u) << 15));
                wcnt += __popc(m);
            }
            if (dbg & 2) wcnt = min(wcnt, 1);
            __syncwarp();
            const unsigned short *wl = s_wl[warp];
            for (int i = 0; i < wcnt; ++i) {
                const unsigned e = wl[i];
                const int s = (int)(e & 0x7fffu);
                const float *cand = &s_cand[0][0] + s * CR;
                if (e & 0x8000u)
                    eval_fast<CP, PX, CR, false>(pf, yr, xbr, cand, make_int4(0, 0, 0, 0), yg, xb, c0 + s, best, bests);
                else
                    eval_fast<CP, PX, CR, true>(pf, yr, xbr, cand, s_win[s], yg, xb, c0 + s, best, bests);
            }
            __syncwarp();
        }

        // ---- labels --------------------------------------------------------------------------
        int kk[PX];
        bool all_found = true;
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            const bool v = (vmask >> j) & 1u;
            kk[j] = (v && bests[j] >= 0) ? s_sorted[bests[j]] : -1;
            all_found = all_found && v && bests[j] >= 0;
        }
        if (PX == 4 && all_found && (LW & 3) == 0) {
            *reinterpret_cast<int4 *>(labels + (int64_t)y * LW + xb) =
                make_int4(kk[0] + start_label, kk[1 % PX] + start_label, kk[2 % PX] + start_label,
                          kk[3 % PX] + start_label);
        } else {
#pragma unroll
            for (int j = 0; j < PX; ++j)
                if (kk[j] >= 0) labels[(int64_t)y * LW + xb + j] = kk[j] + start_label;
        }

        // ---- fused centre update -----------------------------------------------------------------
        // Per lane: the pixels that share the lane's leading winner are summed (fixed order) into one
        // record, every other pixel is a record of its own; records sit in shared memory at an odd
        // stride (conflict-free stores), the tile folds them into its per-slot accumulators with lanes
        // across fields and issues one RED.64 per touched (centre, field) at the end.  Integer sums of
        // f - o: independent of scheduling.  Pixels that kept their previous centre (no candidate) and
        // slots beyond the accumulator rows go to HBM directly with the same quantisation.
        auto direct = [&](int kcur, int cnt, int sxl, const float (&fs)[CP]) {
            unsigned long long *a = acc + (int64_t)kcur * (3 + Cf);
            atomicAdd(&a[0], (unsigned long long)cnt);
            atomicAdd(&a[1], (unsigned long long)((long long)cnt * yg));
            atomicAdd(&a[2], (unsigned long long)((long long)sxl + (long long)cnt * tx0));
#pragma unroll
            for (int c = 0; c < CP; ++c)
                if (c < Cf)
                    atomicAdd(&a[3 + c], (unsigned long long)((long long)__float2int_rn(fs[c] * fix_scale32) * fix_ratio +
                                                              (long long)cnt * s_off64[c]));
        };
        int lead = -1;
#pragma unroll
        for (int j = PX - 1; j >= 0; --j)
            if (((vmask >> j) & 1u) && bests[j] >= 0) lead = bests[j];
        // number of records of this lane: the lead group + every other assigned pixel
        int nrec_l = 0;
        if (lead >= 0 && lead < kAcc) nrec_l = 1;
#pragma unroll
        for (int j = 0; j < PX; ++j)
            if (((vmask >> j) & 1u) && bests[j] >= 0 && bests[j] != lead && bests[j] < kAcc) ++nrec_l;
        int incl = nrec_l;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int nrec_w = __shfl_sync(0xffffffffu, incl, 31);   // records of this warp in this phase
        int(*w_rec)[RS] = s_rec + warp * T::kRecW;               // the warp's own record region: no CTA barrier
        int rbase = incl - nrec_l;
        if (lead >= 0) {
            int cnt = 0, sxl = 0;
            float fs[CP];
#pragma unroll
            for (int c = 0; c < CP; ++c) fs[c] = 0.0f;
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                if (((vmask >> j) & 1u) && bests[j] == lead) {
                    cnt += 1;
                    sxl += xb + j - tx0;
#pragma unroll
                    for (int c = 0; c < CP; ++c) fs[c] = __fadd_rn(fs[c], pf[j][c]);
                }
            }
            if (lead < kAcc) {
                int *r = w_rec[rbase];
                r[0] = cnt | ((cnt * (y - ty0)) << 13);
                r[1] = sxl | (lead << 16);
#pragma unroll
                for (int c = 0; c < CP; ++c) r[2 + c] = __float2int_rn(fs[c] * fix_scale32);
                ++rbase;
            } else {
                direct(s_sorted[lead], cnt, sxl, fs);
            }
        }
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            if (!((vmask >> j) & 1u) || (bests[j] == lead && lead >= 0)) continue;
            if (bests[j] >= 0 && bests[j] < kAcc) {
                int *r = w_rec[rbase];
                r[0] = 1 | ((y - ty0) << 13);
                r[1] = (xb + j - tx0) | (bests[j] << 16);
#pragma unroll
                for (int c = 0; c < CP; ++c) r[2 + c] = __float2int_rn(pf[j][c] * fix_scale32);
                ++rbase;
            } else {
                int kcur = kk[j];
                if (kcur < 0) kcur = labels[(int64_t)y * LW + xb + j] - start_label;   // kept its previous centre
                if (kcur < 0) continue;
                float fs[CP];
#pragma unroll
                for (int c = 0; c < CP; ++c) fs[c] = pf[j][c];
                direct(kcur, 1, xb + j - tx0, fs);
            }
        }
        __syncwarp();
        {
            // lanes across FIELDS (RPW records per warp instruction): the atomics of one instruction go to
            // different words unless two of its records share a slot.  Each warp folds its own records.
            constexpr int RPW = (NFR <= 10) ? 3 : (NFR <= 16) ? 2 : 1;   // records per warp pass
            constexpr int FPL = (NFR + 31) / 32;                           // fields per lane when NFR > 32
            const int rr = (RPW > 1) ? lane / NFR : 0, f0 = (RPW > 1) ? lane - rr * NFR : lane;
            if (rr < RPW) {
                for (int r = rr; r < nrec_w; r += RPW) {
                    const int *rec = w_rec[r];
                    const int w1 = rec[1];
                    int *a = s_acc[w1 >> 16];
#pragma unroll
                    for (int q = 0; q < FPL; ++q) {
                        const int f = f0 + 32 * q;
                        if (f < NFR) {
                            const int v = (f == 1) ? (w1 & 0xffff) : rec[f];
                            atomicAdd(&a[f], v);
                        }
                    }
                }
            }
        }
        __syncwarp();
    }   // strip phases
    __syncthreads();
    const int nslots = min(nids, kAcc);
    for (int i = tid; i < nslots * (3 + CP); i += NT) {
        const int slot = i / (3 + CP), f = i - slot * (3 + CP);
        if (f >= 3 + Cf) continue;
        const unsigned w0 = (unsigned)s_acc[slot][0];
        const int cnt = (int)(w0 & 0x1fffu);
        if (cnt == 0) continue;
        long long g;
        if (f == 0) g = cnt;
        else if (f == 1) g = (long long)(w0 >> 13) + (long long)cnt * (ty0 + y_off);
        else if (f == 2) g = (long long)s_acc[slot][1] + (long long)cnt * tx0;
        else g = (long long)s_acc[slot][f - 1] * fix_ratio + (long long)cnt * s_off64[f - 3];
        if (g != 0) atomicAdd(&acc[(int64_t)s_sorted[slot] * (3 + Cf) + f], (unsigned long long)g);
    }
}

static int g_fast_warps = 8;   // tuning knob (obia_b200_slic_fast_variant): warps per CTA, 8 or 4
static int g_fast_dbg = 0;     // ablation switches (bits 8.. of the same call): 2 one candidate per chunk, 4 no seed

// Strip phases per CTA tile (32 columns x 32 * NS rows for up to 8 channels).  A tile stages the centres within
// 2 * step of it; with small steps a tall tile collects more than the 128 records that are resident at once and
// would rebuild them per phase, so the tile height follows the grid step.  (The fixed-point scale of the per-tile
// sums depends on the tile size: the choice is a function of (Cf, step) only, so every path -- whole raster, row
// strips, window batches -- quantises identically.)
int fast_pick_ns(int Cf, int step_y)
{
    if (Cf > 8) return 2;
    return step_y >= 14 ? 4 : (step_y >= 7 ? 2 : 1);
}

// 32-bit fixed point for the per-tile sums: same derivation as the exact kernel (slic.cu), with
// one more bit of head-room because |f - o| can reach twice the feature range
static void fast_fix_params(int64_t Hg, int64_t W, int step_y, int step_x, double fix_scale, int strips32,
                            float &fix_scale32, long long &fix_ratio)
{
    const int64_t reach = std::min<int64_t>(Hg * W, (int64_t)(4 * step_y + 1) * (4 * step_x + 1));
    int bits_px = 1;
    while ((1LL << bits_px) < reach + 1) ++bits_px;
    int lg_ns = 0;
    while ((1 << lg_ns) < strips32) ++lg_ns;
    lg_ns += 1;
    fix_scale32 = (float)ldexp(fix_scale, bits_px - 42 - lg_ns);
    fix_ratio = 1LL << (42 - bits_px + lg_ns);
}

template <int CP, int PX, int NS, int NW>
static int launch_fast_t(const float *feat, const uint8_t *mask, const float *centres, const SlicWs &w, int32_t *labels,
                         int64_t H, int64_t W, int64_t pitch, int Cf, float sw, int step_y, int step_x, int start_label,
                         int ignore_color, double fix_scale, int32_t *status, int y_off, int64_t Hg, cudaStream_t st)
{
    using T = FastTraits<CP, NW>;
    constexpr int RW = 32 / (16 / PX), TH = (NW / 2) * RW;
    float fix_scale32;
    long long fix_ratio;
    fast_fix_params(Hg, W, step_y, step_x, fix_scale, NS * TH / 32, fix_scale32, fix_ratio);
    dim3 grid((unsigned)ceil_div(W, 32), (unsigned)ceil_div(H, NS * TH));
    constexpr size_t dyn = T::kDyn;
    auto kern = slic_assign_fast_kernel<CP, PX, NS, NW, false>;
    OBIA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    prof_begin(st);
    kern<<<grid, NW * 32, dyn, st>>>(feat, mask, centres, w.head, w.next, labels, w.acc, (int)H, (int)W, pitch, Cf, sw,
                                     1.0f / sw, step_y, step_x, (int)w.ncy, (int)w.ncx, start_label, ignore_color, fix_scale,
                                     fix_scale32, fix_ratio, status, y_off, (int)Hg, g_fast_dbg, nullptr, (int)W,
                                     H * pitch);
    prof_end(st);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

int launch_assign_fast(const float *feat, const uint8_t *mask, const float *centres, const SlicWs &w, int32_t *labels,
                       int64_t H, int64_t W, int64_t pitch, int Cf, float sw, int step_y, int step_x, int start_label,
                       int ignore_color, double fix_scale, int32_t *status, int y_off, int64_t Hg, cudaStream_t st)
{
#define OBIA_FAST_ARGS feat, mask, centres, w, labels, H, W, pitch, Cf, sw, step_y, step_x, start_label, ignore_color, \
                       fix_scale, status, y_off, Hg, st
    const int ns = fast_pick_ns(Cf, step_y);
    if (Cf <= 4) {
        if (ns == 1) return launch_fast_t<4, 4, 1, 8>(OBIA_FAST_ARGS);
        if (ns == 2) return launch_fast_t<4, 4, 2, 8>(OBIA_FAST_ARGS);
        if (g_fast_warps == 4) return launch_fast_t<4, 4, 4, 4>(OBIA_FAST_ARGS);
        return launch_fast_t<4, 4, 4, 8>(OBIA_FAST_ARGS);
    }
    if (Cf <= 8) {
        if (ns == 1) return launch_fast_t<8, 4, 1, 8>(OBIA_FAST_ARGS);
        if (ns == 2) return launch_fast_t<8, 4, 2, 8>(OBIA_FAST_ARGS);
        if (g_fast_warps == 4) return launch_fast_t<8, 4, 4, 4>(OBIA_FAST_ARGS);
        return launch_fast_t<8, 4, 4, 8>(OBIA_FAST_ARGS);
    }
    if (Cf <= 16) return launch_fast_t<16, 2, 2, 8>(OBIA_FAST_ARGS);
    if (Cf <= 32) return launch_fast_t<32, 2, 2, 8>(OBIA_FAST_ARGS);
    return launch_fast_t<64, 2, 2, 8>(OBIA_FAST_ARGS);
#undef OBIA_FAST_ARGS
}

// ---- batched launch (tiled driver): every window of a slab in one grid -------------------------------
// rows of a CTA tile and its number of 32-row strips for the variant that serves (Cf, ns)
static void fast_variant_geometry(int Cf, int ns, int &tile_rows, int &strips32)
{
    if (Cf <= 8) {            // <CP, 4, ns, 8>: 8 rows per warp strip, 32 rows per phase
        tile_rows = 32 * ns;
        strips32 = ns;
    } else {                  // <CP, 2, 2, 8>: 4 rows per warp strip, 16 rows per phase, 2 phases
        tile_rows = 32;
        strips32 = 1;
    }
}

void fast_batch_fix_params(WinDesc *descs_host, int64_t B, int Cf)
{
    for (int64_t i = 0; i < B; ++i) {
        WinDesc &d = descs_host[i];
        if (!d.valid) continue;
        int tile_rows, strips32;
        d.pad = fast_pick_ns(Cf, d.step_y);
        fast_variant_geometry(Cf, d.pad, tile_rows, strips32);
        long long ratio;
        fast_fix_params(d.h, d.w, d.step_y, d.step_x, d.fix_scale, strips32, d.fix_scale32, ratio);
        d.fix_ratio = ratio;
    }
}

template <int CP, int PX, int NS, int NW>
static int launch_fast_batch_t(const float *feat, const uint8_t *mask, const float *centres, const int32_t *head,
                               const int32_t *next, unsigned long long *acc, int32_t *labels, const WinDesc *batch,
                               int64_t B, int hmax, int wmax, int64_t slab_rows, int LW, int64_t pitch, int Cf,
                               int start_label, int ignore_color, int32_t *status, cudaStream_t st)
{
    using T = FastTraits<CP, NW>;
    constexpr int RW = 32 / (16 / PX), TH = (NW / 2) * RW;
    constexpr size_t dyn = T::kDyn;
    auto kern = slic_assign_fast_kernel<CP, PX, NS, NW, true>;
    OBIA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    for (int64_t z0 = 0; z0 < B; z0 += 65535) {
        const int64_t nz = std::min<int64_t>(65535, B - z0);
        dim3 grid((unsigned)ceil_div(wmax, 32), (unsigned)ceil_div(hmax, NS * TH), (unsigned)nz);
        kern<<<grid, NW * 32, dyn, st>>>(feat, mask, centres, head, next, labels, acc, 0, 0, pitch, Cf, 0.0f, 0.0f, 1, 1, 1, 1,
                                         start_label, ignore_color, 1.0, 1.0f, 1, status + 4 + z0, 0, 0, 0, batch + z0, LW,
                                         slab_rows * pitch);
        OBIA_LAUNCH_CHECK();
    }
    return OBIA_B200_OK;
}

int launch_assign_fast_batch(const float *feat, const uint8_t *mask, const float *centres, const int32_t *head,
                             const int32_t *next, unsigned long long *acc, int32_t *labels, const WinDesc *batch,
                             int64_t B, int hmax, int wmax, int64_t slab_rows, int LW, int64_t pitch, int Cf,
                             int start_label, int ignore_color, int32_t *status, int ns_mask, cudaStream_t st)
{
#define OBIA_FAST_ARGS feat, mask, centres, head, next, acc, labels, batch, B, hmax, wmax, slab_rows, LW, pitch, Cf, \
                       start_label, ignore_color, status, st
    if (Cf <= 8) {      // one launch per tile variant present in the batch (windows of the others leave at once)
        int rc = OBIA_B200_OK;
        if (Cf <= 4) {
            if (!rc && (ns_mask & 1)) rc = launch_fast_batch_t<4, 4, 1, 8>(OBIA_FAST_ARGS);
            if (!rc && (ns_mask & 2)) rc = launch_fast_batch_t<4, 4, 2, 8>(OBIA_FAST_ARGS);
            if (!rc && (ns_mask & 4)) rc = launch_fast_batch_t<4, 4, 4, 8>(OBIA_FAST_ARGS);
        } else {
            if (!rc && (ns_mask & 1)) rc = launch_fast_batch_t<8, 4, 1, 8>(OBIA_FAST_ARGS);
            if (!rc && (ns_mask & 2)) rc = launch_fast_batch_t<8, 4, 2, 8>(OBIA_FAST_ARGS);
            if (!rc && (ns_mask & 4)) rc = launch_fast_batch_t<8, 4, 4, 8>(OBIA_FAST_ARGS);
        }
        return rc;
    }
    if (Cf <= 16) return launch_fast_batch_t<16, 2, 2, 8>(OBIA_FAST_ARGS);
    if (Cf <= 32) return launch_fast_batch_t<32, 2, 2, 8>(OBIA_FAST_ARGS);
    return launch_fast_batch_t<64, 2, 2, 8>(OBIA_FAST_ARGS);
#undef OBIA_FAST_ARGS
}

}  // namespace obia

extern "C" int obia_b200_slic_fast_variant(int32_t warps_per_cta)
{
    const int warps = warps_per_cta & 0xff;
    if (warps != 4 && warps != 8) return obia::set_err(OBIA_B200_ERR_ARG, "slic_fast_variant: 4 or 8");
    obia::g_fast_warps = warps;
    obia::g_fast_dbg = (warps_per_cta >> 8) & 0xff;
    return OBIA_B200_OK;
}

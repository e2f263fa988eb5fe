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
// Results differ from the exact kernel only where two candidates are within float32 rounding of
// each other (tests: >= 99.9 % agreement per sweep, ARI reported).
#include "slic_common.cuh"

namespace obia {

template <int CP, int NW> struct FastTraits {
    static constexpr int CR = (3 + CP + 3) / 4 * 4;          // floats per candidate record
    static constexpr int kIds = 1024;
    static constexpr int kChk = (CP <= 32) ? 64 : 32;
    static constexpr int kAcc = (CP <= 16) ? 128 : 64;
    static constexpr int kRecFull = (CP <= 4) ? 768 : (CP == 8) ? 608 : (CP == 16) ? 272 : 512;
    static constexpr int kRec = kRecFull * NW / 8;
    static constexpr bool kDynRec = CP >= 32;
};

// Score of the PX pixels of this lane against one candidate record (see the header):
//   rec[0] = B_k, rec[1] = -2 w cy', rec[2] = -2 w cx', rec[3 + c] = -2 m'_c.
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

template <int CP, int PX, int NS, int NW>
__global__ void __launch_bounds__(NW * 32, (CP <= 16) ? (24 / NW) : 1)
slic_assign_fast_kernel(const float *__restrict__ feat, const uint8_t *__restrict__ mask,
                        const float *__restrict__ centres, const int32_t *__restrict__ head,
                        const int32_t *__restrict__ next, int32_t *__restrict__ labels,
                        unsigned long long *__restrict__ acc, int H, int W, int64_t pitch, int Cf,
                        float spatial_weight, int step_y, int step_x, int ncy, int ncx, int start_label,
                        int ignore_color, double fix_scale, float fix_scale32, long long fix_ratio,
                        int32_t *status, int y_off, int Hg)
{
    using T = FastTraits<CP, NW>;
    constexpr int LPR = 16 / PX;         // lanes per strip row
    constexpr int RW = 32 / LPR;         // rows per warp strip
    constexpr int TH = (NW / 2) * RW;    // rows per phase (two strips side by side)
    constexpr int NF = 3 + CP;
    constexpr int CR = T::CR, kIds = T::kIds, kChk = T::kChk, kAcc = T::kAcc, kRec = T::kRec;
    constexpr bool kRounds = CP >= 16;
    constexpr int NT = NW * 32;
    __shared__ int s_ids[kIds];
    __shared__ int s_sorted[kIds];
    __shared__ int s_nids, s_nrec;
    __shared__ int4 s_win[kChk];
    __shared__ float2 s_cyx[kChk];
    __shared__ __align__(16) float s_cand[kChk][CR];
    __shared__ float s_off[CP];
    __shared__ long long s_off64[CP];
    __shared__ int s_acc[kAcc][NF];
    extern __shared__ __align__(16) int s_dyn[];
    __shared__ __align__(16) int s_rec_static[T::kDynRec ? 1 : kRec][NF + 1];
    int(*s_rec)[NF + 1] = T::kDynRec ? reinterpret_cast<int(*)[NF + 1]>(s_dyn) : s_rec_static;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * 32, ty0 = blockIdx.y * (TH * NS);
    const int tx1 = min(tx0 + 32, W) - 1, ty1 = min(ty0 + TH * NS, H) - 1;   // inclusive
    // tile-local origin (global row / column of the tile centre)
    const int X0 = tx0 + 16, Y0 = ty0 + y_off + (TH * NS) / 2;

    // ---- collect candidate centre ids (as in the exact kernel) --------------------------------
    if (tid == 0) {
        s_nids = 0;
        s_nrec = 0;
    }
    for (int i = tid; i < kAcc * NF; i += NT) (&s_acc[0][0])[i] = 0;
    __syncthreads();
    {
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
    for (int i = tid; i < nids; i += NT) {
        const int k = s_ids[i];
        int r = 0;
        for (int j = 0; j < nids; ++j) r += (s_ids[j] < k);
        s_sorted[r] = k;
    }
    __syncthreads();
    // colour origin of the tile: the colour of its middle candidate (live centres only are binned,
    // so it is finite); 0 in the spatial-only pass and when the tile has no candidate
    if (tid < CP) {
        float o = 0.0f;
        if (nids > 0 && tid < Cf && !ignore_color) o = centres[(int64_t)s_sorted[nids >> 1] * (2 + Cf) + 2 + tid];
        s_off[tid] = o;
        s_off64[tid] = __double2ll_rn((double)o * fix_scale);
    }
    __syncthreads();

    const bool single_chunk = nids <= kChk;
    const float INF = __int_as_float(0x7f800000);
    for (int sp = 0; sp < NS; ++sp) {
        // ---- this lane's pixels ------------------------------------------------------------
        const int sx0 = tx0 + (warp & 1) * 16, sy0 = ty0 + sp * TH + (warp >> 1) * RW;
        const int y = sy0 + lane / LPR;
        const int xb = sx0 + (lane % LPR) * PX;
        const bool row_ok = y < H;
        float pf[PX][CP];
        unsigned vmask = 0;
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            bool v = row_ok && (xb + j) < W;
            if (v && mask) v = mask[(int64_t)y * W + xb + j] != 0;
            vmask |= (v ? 1u : 0u) << j;
        }
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < Cf && row_ok && xb < W) {
                const float *src = feat + (int64_t)c * H * pitch + (int64_t)y * pitch + xb;
                if constexpr (PX == 4) {
                    v = *reinterpret_cast<const float4 *>(src);
                } else {
                    const float2 t = *reinterpret_cast<const float2 *>(src);
                    v.x = t.x; v.y = t.y;
                }
            }
            const float o = s_off[c];
            pf[0][c] = v.x - o;
            pf[1][c] = v.y - o;
            if constexpr (PX == 4) {
                pf[2][c] = v.z - o;
                pf[3][c] = v.w - o;
            }
        }
        const int yg = y + y_off;
        const float yr = (float)(yg - Y0), xbr = (float)(xb - X0);
        // candidate-independent part of the distance (needed for the pruning bound only)
        float A[PX];
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            const float xr = xbr + (float)j;
            float a = spatial_weight * (yr * yr + xr * xr);
            if (!ignore_color) {
#pragma unroll
                for (int c = 0; c < CP; ++c) a = fmaf(pf[j][c], pf[j][c], a);
            }
            A[j] = ((vmask >> j) & 1u) ? a : 0.0f;
        }

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
                for (int sI = tid; sI < nc; sI += NT) {
                    const int k = s_sorted[c0 + sI];
                    const float *rec = centres + (int64_t)k * (2 + Cf);
                    const float cy = rec[0], cx = rec[1];
                    float m[CP];
#pragma unroll
                    for (int c = 0; c < CP; ++c) m[c] = (c < Cf && !ignore_color) ? rec[2 + c] : 0.0f;
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
                    float B = spatial_weight * (cyr * cyr + cxr * cxr);
                    float *cr = s_cand[sI];
                    cr[1] = -2.0f * spatial_weight * cyr;
                    cr[2] = -2.0f * spatial_weight * cxr;
#pragma unroll
                    for (int c = 0; c < CP; ++c) {
                        const float mc = (c < Cf && !ignore_color) ? m[c] - s_off[c] : 0.0f;
                        B = fmaf(mc, mc, B);
                        cr[3 + c] = -2.0f * mc;
                    }
                    cr[0] = B;
#pragma unroll
                    for (int c = 3 + CP; c < CR; ++c) cr[c] = 0.0f;
                }
                __syncthreads();
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
                bool full = false;
                if (sc < nc) {
                    const int4 w = s_win[sc];
                    if (!(w.x > wy1 || w.y <= wy0 || w.z > wx1 || w.w <= wx0)) {
                        const float2 c = s_cyx[sc];
                        const float ddy = fmaxf(0.0f, fmaxf((float)wy0 - c.x, c.x - (float)wy1));
                        const float ddx = fmaxf(0.0f, fmaxf((float)wx0 - c.y, c.y - (float)wx1));
                        hit[half] = true;
                        lb[half] = (ddy * ddy + ddx * ddx) * spatial_weight * 0.9999f;
                        seedkey = min(seedkey, (__float_as_uint(lb[half]) & ~63u) | (unsigned)sc);
                        full = w.x <= wy0 && w.y > wy1 && w.z <= wx0 && w.w > wx1;
                    }
                }
                fullbits[half] = __ballot_sync(0xffffffffu, full);
            }
            seedkey = __reduce_min_sync(0xffffffffu, seedkey);
            if (seedkey != 0xffffffffu && __uint_as_float(seedkey & ~63u) < wbound) {
                const int s = (int)(seedkey & 63u);
                float tb[PX];
                int ts[PX];
#pragma unroll
                for (int j = 0; j < PX; ++j) {
                    tb[j] = ((vmask >> j) & 1u) ? INF : 0.0f;
                    ts[j] = -1;
                }
                eval_fast<CP, PX, CR, true>(pf, yr, xbr, s_cand[s], s_win[s], yg, xb, 0, tb, ts);
                float m = tb[0] + A[0];
#pragma unroll
                for (int j = 1; j < PX; ++j) m = fmaxf(m, tb[j] + A[j]);
                m = fmaxf(m, 0.0f) * 1.0001f + 1e-30f;   // rounding slack of the expanded form
                wbound = fminf(wbound, __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(m))));
            }
#pragma unroll
            for (int half = 0; half < kChk / 32; ++half) {
                unsigned m = __ballot_sync(0xffffffffu, hit[half] && lb[half] <= wbound);
                const unsigned fb = fullbits[half];
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    const int s = half * 32 + b;
                    if ((fb >> b) & 1u)
                        eval_fast<CP, PX, CR, false>(pf, yr, xbr, s_cand[s], make_int4(0, 0, 0, 0), yg, xb, c0 + s,
                                                     best, bests);
                    else
                        eval_fast<CP, PX, CR, true>(pf, yr, xbr, s_cand[s], s_win[s], yg, xb, c0 + s, best, bests);
                }
            }
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
        if (PX == 4 && all_found && (W & 3) == 0) {
            *reinterpret_cast<int4 *>(labels + (int64_t)y * W + xb) =
                make_int4(kk[0] + start_label, kk[1 % PX] + start_label, kk[2 % PX] + start_label,
                          kk[3 % PX] + start_label);
        } else {
#pragma unroll
            for (int j = 0; j < PX; ++j)
                if (kk[j] >= 0) labels[(int64_t)y * W + xb + j] = kk[j] + start_label;
        }

        // ---- fused centre update (records -> tile accumulators -> RED.64), sums of f - o ----------
        int lead = -1;
#pragma unroll
        for (int j = PX - 1; j >= 0; --j)
            if (((vmask >> j) & 1u) && bests[j] >= 0) lead = bests[j];
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
                        for (int c = 0; c < CP; ++c) fs[c] = __fadd_rn(fs[c], pf[j][c]);
                    }
                }
            } else {
                slot = bests[item];
                cnt = 1;
                sxl = xb + item - tx0;
#pragma unroll
                for (int c = 0; c < CP; ++c) fs[c] = pf[item % PX][c];
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
                if (kRounds) return false;
            }
            unsigned long long *a = acc + (int64_t)kcur * (3 + Cf);
            atomicAdd(&a[0], (unsigned long long)cnt);
            atomicAdd(&a[1], (unsigned long long)((long long)cnt * yg));
            atomicAdd(&a[2], (unsigned long long)((long long)sxl + (long long)cnt * tx0));
#pragma unroll
            for (int c = 0; c < CP; ++c)
                if (c < Cf)
                    atomicAdd(&a[3 + c], (unsigned long long)((long long)__float2int_rn(fs[c] * fix_scale32) * fix_ratio +
                                                              (long long)cnt * s_off64[c]));
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
            if (lead >= 0) place(PX, s_sorted[lead]);
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                if (!((vmask >> j) & 1u) || bests[j] == lead) continue;
                int kcur = kk[j];
                if (kcur < 0) kcur = labels[(int64_t)y * W + xb + j] - start_label;
                if (kcur < 0) continue;
                place(j, kcur);
            }
            __syncthreads();
            fold();
            __syncthreads();
            if (tid == 0) s_nrec = 0;
            __syncthreads();
        } else {
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
        else g = v * fix_ratio + (long long)cnt * s_off64[f - 3];
        if (g != 0) atomicAdd(&acc[(int64_t)s_sorted[slot] * (3 + Cf) + f], (unsigned long long)g);
    }
}

static int g_fast_warps = 8;   // tuning knob (obia_b200_slic_fast_variant): warps per CTA, 8 or 4

template <int CP, int PX, int NS, int NW>
static int launch_fast_t(const float *feat, const uint8_t *mask, const float *centres, const SlicWs &w, int32_t *labels,
                         int64_t H, int64_t W, int64_t pitch, int Cf, float sw, int step_y, int step_x, int start_label,
                         int ignore_color, double fix_scale, int32_t *status, int y_off, int64_t Hg, cudaStream_t st)
{
    using T = FastTraits<CP, NW>;
    constexpr int RW = 32 / (16 / PX), TH = (NW / 2) * RW;
    // 32-bit fixed point for the per-tile sums: same derivation as the exact kernel (slic.cu), with
    // one more bit of head-room because |f - o| can reach twice the feature range
    const int64_t reach = std::min<int64_t>(Hg * W, (int64_t)(4 * step_y + 1) * (4 * step_x + 1));
    int bits_px = 1;
    while ((1LL << bits_px) < reach + 1) ++bits_px;
    int lg_ns = 0;
    while ((1 << lg_ns) < NS * TH / 32) ++lg_ns;
    lg_ns += 1;
    const float fix_scale32 = (float)ldexp(fix_scale, bits_px - 42 - lg_ns);
    const long long fix_ratio = 1LL << (42 - bits_px + lg_ns);
    dim3 grid((unsigned)ceil_div(W, 32), (unsigned)ceil_div(H, NS * TH));
    constexpr size_t dyn = T::kDynRec ? (size_t)T::kRec * (3 + CP + 1) * sizeof(int) : 0;
    auto kern = slic_assign_fast_kernel<CP, PX, NS, NW>;
    if (dyn > 0)
        OBIA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    prof_begin(st);
    kern<<<grid, NW * 32, dyn, st>>>(feat, mask, centres, w.head, w.next, labels, w.acc, (int)H, (int)W, pitch, Cf, sw,
                                     step_y, step_x, (int)w.ncy, (int)w.ncx, start_label, ignore_color, fix_scale,
                                     fix_scale32, fix_ratio, status, y_off, (int)Hg);
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
    if (Cf <= 4) {
        if (g_fast_warps == 4) return launch_fast_t<4, 4, 4, 4>(OBIA_FAST_ARGS);
        return launch_fast_t<4, 4, 2, 8>(OBIA_FAST_ARGS);
    }
    if (Cf <= 8) {
        if (g_fast_warps == 4) return launch_fast_t<8, 4, 4, 4>(OBIA_FAST_ARGS);
        return launch_fast_t<8, 4, 2, 8>(OBIA_FAST_ARGS);
    }
    if (Cf <= 16) return launch_fast_t<16, 2, 2, 8>(OBIA_FAST_ARGS);
    if (Cf <= 32) return launch_fast_t<32, 2, 2, 8>(OBIA_FAST_ARGS);
    return launch_fast_t<64, 2, 2, 8>(OBIA_FAST_ARGS);
#undef OBIA_FAST_ARGS
}

}  // namespace obia

extern "C" int obia_b200_slic_fast_variant(int32_t warps_per_cta)
{
    if (warps_per_cta != 4 && warps_per_cta != 8) return obia::set_err(OBIA_B200_ERR_ARG, "slic_fast_variant: 4 or 8");
    obia::g_fast_warps = warps_per_cta;
    return OBIA_B200_OK;
}

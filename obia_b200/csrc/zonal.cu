// K4: per-segment, per-band zonal statistics (replaces the per-segment loop of
// obia/segmentation/segment_statistics.py:475-508 -> calculate_spectral_stats
// :113-176, i.e. np.mean / np.var / np.min / np.max / scipy.stats.skew /
// scipy.stats.kurtosis over the pixels of each segment).
//
// Two kernels:
//   zonal_bbox    one pass over the label raster (4 B/pixel): bounding box and
//                 pixel count per label; runs of equal labels are folded per CTA
//                 tile in a shared-memory hash table, one set of global atomics
//                 per distinct label of a tile (bbox.cuh);
//   zonal_gather  one warp per (label, chunk of 4 or 8 bands): walks the label's
//                 bounding box with coalesced label/raster loads, accumulates
//                 pivot-shifted power sums in registers (float32 partials
//                 folded into float64 every 16 pixels), reduces them with
//                 shuffles and writes the finished statistics.  No atomics, no
//                 scratch accumulators, deterministic; SLIC segments are
//                 compact so neighbouring boxes overlap in L2, not in HBM.
#include "common.cuh"
#include "bbox.cuh"

namespace obia {

constexpr int kZB = 8;  // bands per warp pass

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
//
// VEC: the ZB bands of this pass are contiguous and 16-byte aligned in the pixel record, so a
// pixel is fetched with ZB/4 128-bit loads.  Rows are processed four at a time so that a warp
// has four label loads and then up to eight pixel loads in flight instead of a chain of
// dependent round trips per row, and the trips are software-pipelined (labels of the next trip
// requested under the pixel loads of this one).  The kernel waits on those loads (ncu: long-scoreboard
// stalls, 22 % of the warp slots filled at 128 registers).  Measured on c2 (K4 = boxes + gather):
// 8 bands per warp 3.7 ms, 4 bands per warp at 64 registers 3.3 ms, pipelined 2.9 ms / 8 bands pipelined 2.8 ms.
#ifndef OBIA_ZG_ROWS
#define OBIA_ZG_ROWS 4        // rows per trip
#endif
#ifndef OBIA_ZG_BANDS
#define OBIA_ZG_BANDS 8       // bands per warp on the vector path (4: half the registers, twice the label reads)
#endif
template <bool VEC, int ZB>
#ifndef OBIA_ZG_CTAS8
#define OBIA_ZG_CTAS8 2
#endif
__global__ void __launch_bounds__(256, ZB == 4 ? 4 : OBIA_ZG_CTAS8)
zonal_gather_kernel(const int32_t *__restrict__ labels, const float *__restrict__ raw, ZonalWs w, int W,
                    int C, ZBands zb, int Cz, int64_t max_label, double resolution,
                    double *__restrict__ stats, int32_t label_lo, int32_t zero_row)
{
    constexpr int R = OBIA_ZG_ROWS;
    const int lane = threadIdx.x & 31;
    const int64_t L = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // table row
    if (L > max_label) return;
    const int32_t LV = (zero_row && L == 0) ? 0 : (int32_t)L + label_lo - zero_row;   // label value
    const int b0 = blockIdx.y * ZB;
    const int nb = min(ZB, Cz - b0);
    double *out = stats + (L * Cz + b0) * 8;
    const int cnt_total = w.count[L];
    const double NAND = __longlong_as_double(0x7ff8000000000000LL);
    if (cnt_total == 0) {
        for (int i = lane; i < nb * 8; i += 32) out[i] = ((i & 7) == 0 || (i & 7) == 7) ? 0.0 : NAND;
        return;
    }
    const int x0 = w.xmin[L], x1 = w.xmax[L], y0 = w.ymin[L], y1 = w.ymax[L];
    int bidx[ZB];
#pragma unroll
    for (int b = 0; b < ZB; ++b) bidx[b] = zb.band[min(b0 + b, Cz - 1)];

    auto load_px = [&](int64_t pix, float (&v)[ZB]) {
        const float *p = raw + pix * C;
        if (VEC) {
#pragma unroll
            for (int q = 0; q < ZB / 4; ++q) {
                const float4 a = *reinterpret_cast<const float4 *>(p + bidx[0] + 4 * q);
                v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
            }
        } else {
#pragma unroll
            for (int b = 0; b < ZB; ++b) v[b] = p[bidx[b]];
        }
    };

    // pivot per band = the first VALID (non-NaN) sample among the segment's pixels of the first 32-column
    // chunk of its first row that holds one (row y0 holds a pixel by construction); 0 when they are
    // all NaN in that band.  The pivot only conditions the power sums, any finite value is correct.
    float pivot[ZB];
    {
        bool found = false;
#pragma unroll
        for (int b = 0; b < ZB; ++b) pivot[b] = 0.0f;
        for (int xs = x0; xs <= x1 && !found; xs += 32) {
            const int x = xs + lane;
            const bool hit = (x <= x1) && (labels[(int64_t)y0 * W + x] == LV);
            if (__ballot_sync(0xffffffffu, hit)) {
                found = true;
                float v[ZB];
#pragma unroll
                for (int b = 0; b < ZB; ++b) v[b] = 0.0f;
                if (hit) load_px((int64_t)y0 * W + x, v);
#pragma unroll
                for (int b = 0; b < ZB; ++b) {
                    const unsigned m = __ballot_sync(0xffffffffu, hit && v[b] == v[b]);
                    const float pv = __shfl_sync(0xffffffffu, v[b], m ? (__ffs(m) - 1) : 0);
                    pivot[b] = m ? pv : 0.0f;
                }
            }
        }
    }

    const float INF = __int_as_float(0x7f800000);
    // ps[m*8 + b] = float32 partial of the (m+1)-th pivot-shifted power sum of band b (this lane's
    // pixels since the last fold); `tot` = float64 running total of ps[lane] over the whole warp.
    float ps[4 * ZB], mn[ZB], mx[ZB];
    int nvalid[ZB];   // NaN samples are dropped per band (segment_statistics.py:144-147)
    double tot = 0.0;
#pragma unroll
    for (int i = 0; i < 4 * ZB; ++i) ps[i] = 0.0f;
#pragma unroll
    for (int b = 0; b < ZB; ++b) {
        mn[b] = INF;
        mx[b] = -INF;
        nvalid[b] = 0;
    }
    // Fold: transposed butterfly.  At each step a lane keeps the half of the vector selected by
    // its lane bit and adds the partner's copy of that half, so after 5 steps lane l holds the
    // warp total of element l (31 shuffles instead of 32 x 5).  Sums continue in float64.
    auto fold = [&]() {
        // (the first two steps are taken element by element: at most eight float64 partials are live)
        double h8[8];
        {
            const bool up16 = lane & 16, up8 = lane & 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                double lo8, hi8;      // elements i and i + 8 of the 16-vector after the offset-16 step
                if constexpr (ZB == 8) {
                    const float s0 = up16 ? ps[i] : ps[i + 16], k0 = up16 ? ps[i + 16] : ps[i];
                    const float s1 = up16 ? ps[i + 8] : ps[i + 24], k1 = up16 ? ps[i + 24] : ps[i + 8];
                    lo8 = (double)k0 + (double)__shfl_xor_sync(0xffffffffu, s0, 16);
                    hi8 = (double)k1 + (double)__shfl_xor_sync(0xffffffffu, s1, 16);
                } else {
                    // 16 sums: both half-warps end with all of them, lane l with element l & 15
                    lo8 = (double)ps[i] + (double)__shfl_xor_sync(0xffffffffu, ps[i], 16);
                    hi8 = (double)ps[i + 8] + (double)__shfl_xor_sync(0xffffffffu, ps[i + 8], 16);
                }
                const double send = up8 ? lo8 : hi8;
                const double keep = up8 ? hi8 : lo8;
                h8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
        }
        double h4[4];
        {
            const bool up = lane & 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double send = up ? h8[i] : h8[i + 4];
                const double keep = up ? h8[i + 4] : h8[i];
                h4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
        }
        double h2[2];
        {
            const bool up = lane & 2;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double send = up ? h4[i] : h4[i + 2];
                const double keep = up ? h4[i + 2] : h4[i];
                h2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
        }
        {
            const bool up = lane & 1;
            const double send = up ? h2[0] : h2[1];
            const double keep = up ? h2[1] : h2[0];
            tot += keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
#pragma unroll
        for (int i = 0; i < 4 * ZB; ++i) ps[i] = 0.0f;
    };

    int since_fold = 0;
    // Software pipeline over the trips (32 columns x R rows each, row-major over the box): the labels of trip
    // t + 1 are requested while the pixel loads of trip t are in flight, so a trip waits for one memory round
    // trip instead of two dependent ones.
    auto load_hits = [&](int yb, int xs, bool (&h)[R]) {
        const int x = xs + lane;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int y = yb + r;
            h[r] = (x <= x1) && (y <= y1) && (labels[(int64_t)y * W + x] == LV);
        }
    };
    int yb = y0, xs = x0;
    bool hit[R];
    load_hits(yb, xs, hit);
    while (yb <= y1) {
        const int x = xs + lane;
        float v[R][ZB];
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (hit[r]) load_px((int64_t)(yb + r) * W + x, v[r]);
        int nxs = xs + 32, nyb = yb;
        if (nxs > x1) {
            nxs = x0;
            nyb = yb + R;
        }
        bool hnext[R];
#pragma unroll
        for (int r = 0; r < R; ++r) hnext[r] = false;
        if (nyb <= y1) load_hits(nyb, nxs, hnext);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (!hit[r]) continue;
#pragma unroll
            for (int b = 0; b < ZB; ++b) {
                const bool ok = v[r][b] == v[r][b];
                const float d = ok ? v[r][b] - pivot[b] : 0.0f;
                const float dd = d * d;
                nvalid[b] += ok;
                ps[b] += d;
                ps[ZB + b] += dd;
                ps[2 * ZB + b] = fmaf(dd, d, ps[2 * ZB + b]);
                ps[3 * ZB + b] = fmaf(dd, dd, ps[3 * ZB + b]);
                mn[b] = fminf(mn[b], v[r][b]);   // fminf / fmaxf return the non-NaN operand
                mx[b] = fmaxf(mx[b], v[r][b]);
            }
        }
        // warp-uniform: at most R pixels per lane per trip -> <= 32 float32 terms per fold
        if (++since_fold == 32 / R) {
            since_fold = 0;
            fold();
        }
#pragma unroll
        for (int r = 0; r < R; ++r) hit[r] = hnext[r];
        yb = nyb;
        xs = nxs;
    }
    fold();
#pragma unroll
    for (int b = 0; b < ZB; ++b) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            mn[b] = fminf(mn[b], __shfl_xor_sync(0xffffffffu, mn[b], o));
            mx[b] = fmaxf(mx[b], __shfl_xor_sync(0xffffffffu, mx[b], o));
        }
        nvalid[b] = __reduce_add_sync(0xffffffffu, nvalid[b]);
    }
    // lane m*ZB + b holds the m-th power sum of band b: hand the four sums of band b to lane b
    const int bsel = lane & (ZB - 1);
    const double S1 = __shfl_sync(0xffffffffu, tot, bsel);
    const double S2 = __shfl_sync(0xffffffffu, tot, ZB + bsel);
    const double S3 = __shfl_sync(0xffffffffu, tot, 2 * ZB + bsel);
    const double S4 = __shfl_sync(0xffffffffu, tot, 3 * ZB + bsel);
    // lane b finishes band b
#pragma unroll
    for (int b = 0; b < ZB; ++b) {
        if (lane == b && b < nb) {
            double *o = out + b * 8;
            if (nvalid[b] == 0) {   // every sample of the band is NaN: the reference returns NaN statistics
                o[0] = 0.0;
                for (int k = 1; k < 7; ++k) o[k] = NAND;
                o[7] = 0.0;
                continue;
            }
            const double n = (double)nvalid[b];
            const double md = S1 / n;
            const double e2 = S2 / n, e3 = S3 / n, e4 = S4 / n;
            const double mean = (double)pivot[b] + md;
            double m2 = e2 - md * md;
            if (m2 < 0.0) m2 = 0.0;
            const double m3 = e3 - 3.0 * md * e2 + 2.0 * md * md * md;
            const double m4 = e4 - 4.0 * md * e3 + 6.0 * md * md * e2 - 3.0 * md * md * md * md;
            // scipy.stats.skew/kurtosis: NaN when the data are (nearly) constant
            const double thr = resolution * mean;
            const bool degenerate = m2 <= thr * thr;
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

// Many-band variant (Cz >= 24): lanes across BANDS.  One warp per label; the 32 labels of a row
// segment are fetched with one coalesced load, then the warp visits only the matching pixels and
// every lane reads and accumulates its own BPL bands of that pixel (a 64-band pixel is one 256-byte
// coalesced request).  No shuffles: each lane owns its bands' sums from start to finish.
#ifndef OBIA_ZB_CTAS
#define OBIA_ZB_CTAS 3
#endif
template <int BPL>
__global__ void __launch_bounds__(256, OBIA_ZB_CTAS)
zonal_gather_bands_kernel(const int32_t *__restrict__ labels, const float *__restrict__ raw, ZonalWs w, int W,
                          int C, ZBands zb, int Cz, int64_t max_label, double resolution,
                          double *__restrict__ stats, int32_t label_lo, int32_t zero_row)
{
    const int lane = threadIdx.x & 31;
    const int64_t L = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (L > max_label) return;
    const int32_t LV = (zero_row && L == 0) ? 0 : (int32_t)L + label_lo - zero_row;
    const int cnt_total = w.count[L];
    const double NAND = __longlong_as_double(0x7ff8000000000000LL);
    double *out = stats + L * Cz * 8;
    if (cnt_total == 0) {
        for (int i = lane; i < Cz * 8; i += 32) out[i] = ((i & 7) == 0 || (i & 7) == 7) ? 0.0 : NAND;
        return;
    }
    const int x0 = w.xmin[L], x1 = w.xmax[L], y0 = w.ymin[L], y1 = w.ymax[L];
    int bidx[BPL];
    bool bok[BPL];
#pragma unroll
    for (int k = 0; k < BPL; ++k) {
        const int b = lane + 32 * k;            // band slot b of the list <-> lane b % 32
        bok[k] = b < Cz;
        bidx[k] = zb.band[bok[k] ? b : 0];
    }
    float pivot[BPL], s1[BPL], s2[BPL], s3[BPL], s4[BPL], mn[BPL], mx[BPL];
    double d1[BPL], d2[BPL], d3[BPL], d4[BPL];
    int nvalid[BPL];
    bool have_pivot[BPL];
    const float INF = __int_as_float(0x7f800000);
#pragma unroll
    for (int k = 0; k < BPL; ++k) {
        s1[k] = s2[k] = s3[k] = s4[k] = 0.0f;
        d1[k] = d2[k] = d3[k] = d4[k] = 0.0;
        mn[k] = INF;
        mx[k] = -INF;
        pivot[k] = 0.0f;
        nvalid[k] = 0;
        have_pivot[k] = false;
    }
    int pending = 0;
#ifndef OBIA_ZB_U
#define OBIA_ZB_U 4           // matching pixels per trip (c4: 4 at 3 CTAs per SM 11.2 ms, 8 at 2 CTAs 11.7 ms)
#endif
    // chunks of 32 columns, row-major over the box; the labels of the next chunk are requested before the
    // pixel trips of this one (software pipeline, as in zonal_gather_kernel)
    auto load_hit = [&](int y, int xs) { return (xs + lane <= x1) && (labels[(int64_t)y * W + xs + lane] == LV); };
    int y = y0, xs = x0;
    bool hit = load_hit(y, xs);
    while (y <= y1) {
        {
            const int64_t row = (int64_t)y * W;
            unsigned m = __ballot_sync(0xffffffffu, hit);
            int nxs = xs + 32, ny = y;
            if (nxs > x1) {
                nxs = x0;
                ny = y + 1;
            }
            const bool hnext = (ny <= y1) ? load_hit(ny, nxs) : false;
            while (m) {
                // up to U matching pixels per trip: all their loads are issued before any is used
                constexpr int U = OBIA_ZB_U;
                int bpos[U];
                int nu = 0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    bpos[u] = m ? (__ffs(m) - 1) : -1;
                    if (m) {
                        m &= m - 1;
                        ++nu;
                    }
                }
                float v[U][BPL];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (bpos[u] >= 0) {
                        const float *p = raw + (row + xs + bpos[u]) * C;
#pragma unroll
                        for (int k = 0; k < BPL; ++k) v[u][k] = bok[k] ? p[bidx[k]] : 0.0f;
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (bpos[u] < 0) continue;
#pragma unroll
                    for (int k = 0; k < BPL; ++k) {
                        const bool ok = v[u][k] == v[u][k];     // NaN samples are dropped per band
                        if (ok && !have_pivot[k]) {             // pivot = the band's first valid sample
                            have_pivot[k] = true;
                            pivot[k] = v[u][k];
                        }
                        nvalid[k] += ok;
                        const float d = ok ? v[u][k] - pivot[k] : 0.0f;
                        const float dd = d * d;
                        s1[k] += d;
                        s2[k] += dd;
                        s3[k] = fmaf(dd, d, s3[k]);
                        s4[k] = fmaf(dd, dd, s4[k]);
                        mn[k] = fminf(mn[k], v[u][k]);
                        mx[k] = fmaxf(mx[k], v[u][k]);
                    }
                }
                pending += nu;
                if (pending >= 8) {
                    pending = 0;
#pragma unroll
                    for (int k = 0; k < BPL; ++k) {
                        d1[k] += (double)s1[k]; d2[k] += (double)s2[k];
                        d3[k] += (double)s3[k]; d4[k] += (double)s4[k];
                        s1[k] = s2[k] = s3[k] = s4[k] = 0.0f;
                    }
                }
            }
            hit = hnext;
            y = ny;
            xs = nxs;
        }
    }
#pragma unroll
    for (int k = 0; k < BPL; ++k) {
        if (!bok[k]) continue;
        double *o = out + (lane + 32 * k) * 8;
        if (nvalid[k] == 0) {
            o[0] = 0.0;
            for (int q = 1; q < 7; ++q) o[q] = NAND;
            o[7] = 0.0;
            continue;
        }
        const double n = (double)nvalid[k];
        const double S1 = d1[k] + (double)s1[k], S2 = d2[k] + (double)s2[k];
        const double S3 = d3[k] + (double)s3[k], S4 = d4[k] + (double)s4[k];
        const double md = S1 / n, e2 = S2 / n, e3 = S3 / n, e4 = S4 / n;
        const double mean = (double)pivot[k] + md;
        double m2 = e2 - md * md;
        if (m2 < 0.0) m2 = 0.0;
        const double m3 = e3 - 3.0 * md * e2 + 2.0 * md * md * md;
        const double m4 = e4 - 4.0 * md * e3 + 6.0 * md * md * e2 - 3.0 * md * md * md * md;
        const double thr = resolution * mean;
        const bool degenerate = m2 <= thr * thr;
        o[0] = n;
        o[1] = mean;
        o[2] = m2;
        o[3] = (double)mn[k];
        o[4] = (double)mx[k];
        o[5] = degenerate ? NAND : m3 / (m2 * sqrt(m2));
        o[6] = degenerate ? NAND : m4 / (m2 * m2) - 3.0;
        o[7] = mean * n;
    }
}

}  // namespace obia

using namespace obia;

extern "C" int64_t obia_b200_zonal_workspace_bytes(int64_t max_label, int32_t Cz)
{
    if (max_label < 0 || Cz <= 0) return -1;
    return zonal_ws_layout(nullptr, max_label).bytes;
}

static int zonal_stats_impl(const int32_t *labels, const float *raw, int64_t H, int64_t W,
                            int32_t C, const int32_t *bands_host, int32_t Cz, int64_t label_lo, int64_t max_label,
                            double resolution, double *stats, void *workspace, void *stream, int32_t zero_row = 0)
{
    if (!labels || !raw || !bands_host || !stats || !workspace || H <= 0 || W <= 0 || C <= 0 || Cz <= 0 ||
        max_label < 0 || label_lo < 0)
        return set_err(OBIA_B200_ERR_ARG, "zonal_stats: bad argument");
    if (Cz > OBIA_B200_MAX_BANDS)
        return set_err(OBIA_B200_ERR_UNSUPPORTED, "zonal_stats: at most %d statistics bands per call",
                       OBIA_B200_MAX_BANDS);
    if (H * W >= 0x7fffffffLL || max_label + label_lo >= 0x7fffffffLL)
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
    const int32_t lo = (int32_t)label_lo;
    zonal_init_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(w, n);
    OBIA_LAUNCH_CHECK();
    zonal_bbox_launch(labels, w, H, W, max_label, lo, zero_row, st);
    OBIA_LAUNCH_CHECK();
    if (Cz >= 24) {
        // many bands: lanes across bands, one pass over the raster for all of them
        const unsigned g = (unsigned)ceil_div(n, 8);
        if (Cz <= 32)
            zonal_gather_bands_kernel<1><<<g, 256, 0, st>>>(labels, raw, w, (int)W, C, zb, Cz, max_label, resolution, stats, lo, zero_row);
        else
            zonal_gather_bands_kernel<2><<<g, 256, 0, st>>>(labels, raw, w, (int)W, C, zb, Cz, max_label, resolution, stats, lo, zero_row);
        OBIA_LAUNCH_CHECK();
        return OBIA_B200_OK;
    }
    // vector path: every pass reads its kZB contiguous, 16-byte aligned floats of the pixel record
    auto vec_ok = [&](int zbn) {
        bool vec = (C % 4 == 0) && (Cz % zbn == 0) && ((reinterpret_cast<uintptr_t>(raw) & 15) == 0);
        for (int b = 0; b < Cz && vec; ++b)
            vec = (b % zbn == 0) ? (zb.band[b] % 4 == 0) : (zb.band[b] == zb.band[b - 1] + 1);
        return vec;
    };
#if OBIA_ZG_BANDS == 4
    if (vec_ok(4)) {
        dim3 grid4((unsigned)ceil_div(n, 8), (unsigned)ceil_div(Cz, 4));
        zonal_gather_kernel<true, 4><<<grid4, 256, 0, st>>>(labels, raw, w, (int)W, C, zb, Cz, max_label, resolution,
                                                            stats, lo, zero_row);
        OBIA_LAUNCH_CHECK();
        return OBIA_B200_OK;
    }
#endif
    dim3 grid((unsigned)ceil_div(n, 8), (unsigned)ceil_div(Cz, kZB));
    if (vec_ok(kZB))
        zonal_gather_kernel<true, kZB><<<grid, 256, 0, st>>>(labels, raw, w, (int)W, C, zb, Cz, max_label,
                                                             resolution, stats, lo, zero_row);
    else
        zonal_gather_kernel<false, kZB><<<grid, 256, 0, st>>>(labels, raw, w, (int)W, C, zb, Cz, max_label,
                                                              resolution, stats, lo, zero_row);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

extern "C" int obia_b200_zonal_stats(const int32_t *labels, const float *raw, int64_t H, int64_t W,
                                     int32_t C, const int32_t *bands_host, int32_t Cz, int64_t max_label,
                                     double resolution, double *stats, void *workspace, void *stream)
{
    return zonal_stats_impl(labels, raw, H, W, C, bands_host, Cz, 0, max_label, resolution, stats, workspace, stream);
}

extern "C" int obia_b200_zonal_stats_range(const int32_t *labels, const float *raw, int64_t H, int64_t W,
                                           int32_t C, const int32_t *bands_host, int32_t Cz, int64_t label_lo,
                                           int64_t n_rows, int32_t zero_row, double resolution, double *stats,
                                           void *workspace, void *stream)
{
    if (n_rows <= 0 || (zero_row != 0 && zero_row != 1) || (zero_row && label_lo <= 0))
        return set_err(OBIA_B200_ERR_ARG, "zonal_stats_range: bad argument");
    return zonal_stats_impl(labels, raw, H, W, C, bands_host, Cz, label_lo, n_rows - 1, resolution, stats, workspace,
                            stream, zero_row);
}

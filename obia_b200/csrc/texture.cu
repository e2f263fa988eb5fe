// K5: per-segment GLCM texture features (replaces calculate_textural_stats,
// obia/segmentation/segment_statistics.py:179-298, as create_objects applies it per segment,
// :496-508).  Reference semantics per (segment, band):
//   crop = bounding box of the segment, pixels outside the segment (and NaN samples) -> 0   (:214-247)
//   q    = uint8(((crop - min) / (max - min)) * 255), min/max over the whole crop, zeros included;
//          constant crop -> all zeros                                                     (:251-258)
//   GLCM = graycomatrix(q, distances=[2], angles=[0, pi/4, pi/2, 3pi/4], levels=256,
//                       symmetric=True, normed=True)                                       (:261-268)
//        = pixel offsets (0,2), (1,1), (2,0), (1,-1)  [UPSTREAM skimage _glcm_loop]
//   out  = mean over the four angles of graycoprops contrast, dissimilarity, homogeneity, ASM,
//          energy, correlation                                                            (:285-296)
// (The reference indexes the band-first crop with [:, :, band]; that axis slip is fixed here, see
//  DESIGN.md / SURVEY.md 8a a11.)
//
// Nothing here needs the 256x256 matrix itself.  With N pairs per angle and the symmetric,
// normalised matrix P:
//   contrast      = sum (i-j)^2 / N          dissimilarity = sum |i-j| / N
//   homogeneity   = sum 1/(1+(i-j)^2) / N
//   mean          = sum (i+j) / 2N           var = sum (i^2+j^2) / 2N - mean^2
//   correlation   = (sum ij / N - mean^2) / var      (1 when var == 0, like skimage's std < 1e-15)
//   ASM           = (sum_{i<j} u_ij^2 + 2 sum_i u_ii^2) / (2 N^2),  u = count of the UNORDERED pair
// All sums are integers (exact, order-independent).  Only ASM needs multiplicities: a shared-memory
// histogram over the 32 896 unordered level pairs whose atomicAdd returns the old count c, so that
// sum u^2 = sum over increments of (2c + 1).  The histogram is cleared by replaying the pairs, so a
// persistent CTA zeroes its 64 KB once.
//
// One CTA per segment, all bands of a chunk of 8 share two passes over the crop (min/max, then
// quantised levels staged in shared memory); 16-bit counters (2 CTAs / SM) for bounding boxes
// below 65 536 pixels, a second launch with 32-bit counters (1 CTA / SM) for the rare larger ones.
#include <type_traits>

#include "common.cuh"
#include "bbox.cuh"

namespace obia {

constexpr int kBins = 256 * 257 / 2;   // unordered level pairs
constexpr int kTileBytes = 16384;       // shared memory for the staged uint8 levels of one band chunk
constexpr int kTileMax = 4096;          // largest staged crop (pixels); 16384 / bands-in-chunk if smaller
constexpr int kTexFields = 6;

struct TexBands {
    int32_t band[OBIA_B200_MAX_BANDS];
};

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int kBC = 8;                 // bands processed per pass over the crop

struct TexSmem {                       // fixed-size part of the dynamic shared memory, after the histogram
    double lut[256];                   // 1 / (1 + d^2)
    unsigned long long red[32][8];     // per-warp partial sums of one angle
    unsigned long long sums[kBC][4][8];   // per (band, angle): s1 s2 sa sb sc sdiag soff | homogeneity (double bits)
    double props[kBC][4][kTexFields];
    float wmn[32][kBC], wmx[32][kBC];
    unsigned wany[32];
    float mn[kBC], mx[kBC];
    unsigned any;
};

// The four angle passes of one band: accumulate the pair sums (and the multiplicity histogram),
// publish them, clear the histogram by replaying the pairs.  STAGED crops (<= kTileMax pixels, levels
// in shared memory) keep every per-thread and per-warp sum in 32 bits.
template <bool WIDE, bool STAGED, int NW, typename LevelFn>
__device__ __forceinline__ void band_passes(unsigned *hist, TexSmem &S, unsigned short *clr, int b, int h, int w,
                                            int warp, int lane, int tid, LevelFn level)
{
    constexpr int NT = NW * 32;
    using acc_t = typename std::conditional<STAGED, unsigned, unsigned long long>::type;
#pragma unroll 1
    for (int a = 0; a < 4; ++a) {
        const int dr = (a == 0) ? 0 : (a == 2) ? 2 : 1;
        const int dc = (a == 0) ? 2 : (a == 1) ? 1 : (a == 2) ? 0 : -1;
        const int nr = h - dr, ncw = w - (dc < 0 ? -dc : dc), c_lo = dc < 0 ? -dc : 0;
        acc_t s1 = 0, s2 = 0, sa = 0, sb = 0, sc = 0, sdiag = 0, soff = 0;
        double sh = 0.0;
        for (int r = warp; r < nr; r += NW)
            for (int c = c_lo + lane; c < c_lo + ncw; c += 32) {
                const int i = level(r, c), j = level(r + dr, c + dc);
                const int d = i > j ? i - j : j - i;
                s1 += (unsigned)d;
                s2 += (unsigned)(d * d);
                sh += S.lut[d];
                sa += (unsigned)(i + j);
                sb += (unsigned)(i * i + j * j);
                sc += (unsigned)(i * j);
                const int lo = min(i, j), hi = max(i, j);
                const int bin = hi * (hi + 1) / 2 + lo;
                unsigned old;
                if (WIDE) {
                    old = atomicAdd(&hist[bin], 1u);
                } else {
                    const int sft = (bin & 1) * 16;
                    old = (atomicAdd(&hist[bin >> 1], 1u << sft) >> sft) & 0xffffu;
                    // staged crops remember the touched word: clearing needs no second look at the levels
                    if (STAGED) clr[r * ncw + (c - c_lo)] = (unsigned short)(bin >> 1);
                }
                if (d == 0)
                    sdiag += (acc_t)2 * old + 1;
                else
                    soff += (acc_t)2 * old + 1;
            }
        unsigned long long t1, t2, ta, tb_, tc, td, to;
        if (STAGED) {   // one REDUX each
            t1 = __reduce_add_sync(0xffffffffu, (unsigned)s1);
            t2 = __reduce_add_sync(0xffffffffu, (unsigned)s2);
            ta = __reduce_add_sync(0xffffffffu, (unsigned)sa);
            tb_ = __reduce_add_sync(0xffffffffu, (unsigned)sb);
            tc = __reduce_add_sync(0xffffffffu, (unsigned)sc);
            td = __reduce_add_sync(0xffffffffu, (unsigned)sdiag);
            to = __reduce_add_sync(0xffffffffu, (unsigned)soff);
        } else {
            t1 = warp_sum_u64(s1);
            t2 = warp_sum_u64(s2);
            ta = warp_sum_u64(sa);
            tb_ = warp_sum_u64(sb);
            tc = warp_sum_u64(sc);
            td = warp_sum_u64(sdiag);
            to = warp_sum_u64(soff);
        }
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) sh += __shfl_xor_sync(0xffffffffu, sh, s);
        if (lane == 0) {
            unsigned long long *rw = S.red[warp];
            rw[0] = t1; rw[1] = t2; rw[2] = ta; rw[3] = tb_; rw[4] = tc; rw[5] = td; rw[6] = to;
            rw[7] = (unsigned long long)__double_as_longlong(sh);
        }
        __syncthreads();   // partial sums published; every increment of this angle has landed
        if (tid < 7) {
            unsigned long long t = 0;
            for (int i = 0; i < NW; ++i) t += S.red[i][tid];
            S.sums[b][a][tid] = t;
        } else if (tid == 7) {
            double t = 0.0;
            for (int i = 0; i < NW; ++i) t += __longlong_as_double((long long)S.red[i][7]);
            S.sums[b][a][7] = (unsigned long long)__double_as_longlong(t);
        }
        // clear the histogram: all non-zero words were touched by this angle's pairs
        if (STAGED) {
            const int npairs = (nr > 0 && ncw > 0) ? nr * ncw : 0;
            for (int t = tid; t < npairs; t += NT) hist[clr[t]] = 0u;
        } else {
            for (int r = warp; r < nr; r += NW)
                for (int c = c_lo + lane; c < c_lo + ncw; c += 32) {
                    const int i = level(r, c), j = level(r + dr, c + dc);
                    const int lo = min(i, j), hi = max(i, j);
                    const int bin = hi * (hi + 1) / 2 + lo;
                    hist[WIDE ? bin : (bin >> 1)] = 0u;
                }
        }
        __syncthreads();   // histogram clean, S.red free
    }
}

template <bool WIDE, int NT>
__global__ void __launch_bounds__(NT)
texture_kernel(const int32_t *__restrict__ labels, const float *__restrict__ raw, ZonalWs bb, int W, int C,
               TexBands tb, int nb, int64_t max_label, int f64, double *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int kHistWords = WIDE ? kBins : (kBins + 1) / 2;
    constexpr int NW = NT / 32;
    unsigned *hist = reinterpret_cast<unsigned *>(smem);
    TexSmem &S = *reinterpret_cast<TexSmem *>(smem + (size_t)kHistWords * 4);
    unsigned char *tiles = smem + (size_t)kHistWords * 4 + sizeof(TexSmem);   // [bands in chunk][cap], 16-bit launch only
    unsigned short *clr = reinterpret_cast<unsigned short *>(tiles + kTileBytes);   // [kTileMax] touched words

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kHistWords; i += NT) hist[i] = 0u;
    for (int i = tid; i < 256; i += NT) S.lut[i] = 1.0 / (1.0 + (double)i * (double)i);
    __syncthreads();

    const double NAN_D = __longlong_as_double(0x7ff8000000000000LL);
    for (int64_t L = blockIdx.x; L <= max_label; L += gridDim.x) {
        const int cnt = bb.count[L];
        double *o = out + L * (int64_t)nb * kTexFields;
        if (cnt <= 0) {
            if (!WIDE)
                for (int i = tid; i < nb * kTexFields; i += NT) o[i] = NAN_D;
            continue;
        }
        const int x0 = bb.xmin[L], y0 = bb.ymin[L];
        const int w = bb.xmax[L] - x0 + 1, h = bb.ymax[L] - y0 + 1;
        const int64_t area = (int64_t)w * h;
        if ((area >= 65536) != WIDE) continue;   // the other launch owns this segment
        const int lab = (int)L;

        for (int k0 = 0; k0 < nb; k0 += kBC) {
            const int nbc = min(kBC, nb - k0);
            // fewer bands in the chunk leave room for larger crops (3 bands: 4096 px, 8 bands: 2048 px)
            const int cap = min(kTileBytes / nbc, kTileMax);
            const bool staged = !WIDE && area <= cap;
            // ---- pass 1 over the crop: min / max of every band of this chunk ----------------------
            // sample = raster value inside the segment, 0 outside the segment and for NaN samples
            float mn[kBC], mx[kBC];
            unsigned anym = 0;
#pragma unroll
            for (int b = 0; b < kBC; ++b) {
                mn[b] = __int_as_float(0x7f800000);
                mx[b] = __int_as_float(0xff800000);
            }
            for (int r = warp; r < h; r += NW)
                for (int c = lane; c < w; c += 32) {
                    const int64_t p = (int64_t)(y0 + r) * W + (x0 + c);
                    const bool inside = labels[p] == lab;
#pragma unroll
                    for (int b = 0; b < kBC; ++b) {
                        if (b < nbc) {
                            float v = inside ? raw[p * C + tb.band[k0 + b]] : 0.0f;
                            const bool valid = inside && (v == v);
                            if (!valid) v = 0.0f;
                            mn[b] = fminf(mn[b], v);
                            mx[b] = fmaxf(mx[b], v);
                            anym |= (valid ? 1u : 0u) << b;
                        }
                    }
                }
#pragma unroll
            for (int b = 0; b < kBC; ++b) {
#pragma unroll
                for (int s = 16; s >= 1; s >>= 1) {
                    mn[b] = fminf(mn[b], __shfl_xor_sync(0xffffffffu, mn[b], s));
                    mx[b] = fmaxf(mx[b], __shfl_xor_sync(0xffffffffu, mx[b], s));
                }
                if (lane == 0) {
                    S.wmn[warp][b] = mn[b];
                    S.wmx[warp][b] = mx[b];
                }
            }
            anym = __reduce_or_sync(0xffffffffu, anym);
            if (lane == 0) S.wany[warp] = anym;
            __syncthreads();
            if (tid < kBC) {
                float a = S.wmn[0][tid], z = S.wmx[0][tid];
                for (int i = 1; i < NW; ++i) {
                    a = fminf(a, S.wmn[i][tid]);
                    z = fmaxf(z, S.wmx[i][tid]);
                }
                S.mn[tid] = a;
                S.mx[tid] = z;
            }
            if (tid == 32) {
                unsigned m = 0;
                for (int i = 0; i < NW; ++i) m |= S.wany[i];
                S.any = m;
            }
            __syncthreads();
            // the reference's arithmetic in the dtype of its masked crop (float32 rasters stay
            // float32, integer rasters become float64), truncated by astype(uint8)
            auto quantise = [&](float v, float lo, float hi) -> int {
                if (hi == lo) return 0;
                if (f64) return (int)(((double)v - (double)lo) / ((double)hi - (double)lo) * 255.0);
                return (int)__fmul_rn(__fdiv_rn(__fsub_rn(v, lo), __fsub_rn(hi, lo)), 255.0f);
            };
            // ---- pass 2 over the crop: quantised levels of every band into shared memory ----------
            if (staged) {
                for (int r = warp; r < h; r += NW)
                    for (int c = lane; c < w; c += 32) {
                        const int64_t p = (int64_t)(y0 + r) * W + (x0 + c);
                        const bool inside = labels[p] == lab;
#pragma unroll
                        for (int b = 0; b < kBC; ++b) {
                            if (b < nbc) {
                                float v = inside ? raw[p * C + tb.band[k0 + b]] : 0.0f;
                                if (!(v == v)) v = 0.0f;
                                tiles[b * cap + r * w + c] = (unsigned char)quantise(v, S.mn[b], S.mx[b]);
                            }
                        }
                    }
                __syncthreads();
            }
            const unsigned any_bands = S.any;

            for (int b = 0; b < nbc; ++b) {
                if (!((any_bands >> b) & 1u)) continue;   // no valid sample: NaN features (:217-231)
                if (staged) {
                    const unsigned char *tile = tiles + b * cap;
                    band_passes<WIDE, true, NW>(hist, S, clr, b, h, w, warp, lane, tid,
                                                [&](int r, int c) -> int { return tile[r * w + c]; });
                } else {
                    const int band = tb.band[k0 + b];
                    const float lo_v = S.mn[b], hi_v = S.mx[b];
                    band_passes<WIDE, false, NW>(hist, S, clr, b, h, w, warp, lane, tid, [&](int r, int c) -> int {
                        const int64_t p = (int64_t)(y0 + r) * W + (x0 + c);
                        float v = (labels[p] == lab) ? raw[p * C + band] : 0.0f;
                        if (!(v == v)) v = 0.0f;
                        return quantise(v, lo_v, hi_v);
                    });
                }
            }
            // ---- features of the chunk: one thread per (band, angle), then the mean over the angles ----
            if (tid < nbc * 4) {
                const int b = tid >> 2, a = tid & 3;
                const int dr = (a == 0) ? 0 : (a == 2) ? 2 : 1;
                const int dc = (a == 0) ? 2 : (a == 1) ? 1 : (a == 2) ? 0 : -1;
                const int nr = h - dr, ncw = w - (dc < 0 ? -dc : dc);
                const unsigned long long npairs = (nr > 0 && ncw > 0) ? (unsigned long long)nr * ncw : 0ull;
                double *pr = S.props[b][a];
                if (npairs == 0) {
                    // empty matrix: every weighted sum is 0 and correlation is 1 (std < 1e-15)
                    pr[0] = pr[1] = pr[2] = pr[3] = pr[4] = 0.0;
                    pr[5] = 1.0;
                } else {
                    const unsigned long long *q = S.sums[b][a];
                    const unsigned long long S1 = q[0], S2 = q[1], SA = q[2], SB = q[3], SC = q[4], SD = q[5], SO = q[6];
                    const double SH = __longlong_as_double((long long)q[7]);
                    const double N = (double)npairs;
                    const double asm_ = ((double)SO + 2.0 * (double)SD) / (2.0 * N * N);
                    pr[0] = (double)S2 / N;
                    pr[1] = (double)S1 / N;
                    pr[2] = SH / N;
                    pr[3] = asm_;
                    pr[4] = sqrt(asm_);
                    // var = (2N SB - SA^2) / 4N^2, cov = (4N SC - SA^2) / 4N^2: exact in 128 bits
                    const unsigned __int128 sa2 = (unsigned __int128)SA * SA;
                    const unsigned __int128 vnum = (unsigned __int128)(2ull * npairs) * SB - sa2;
                    if (vnum == 0) {
                        pr[5] = 1.0;
                    } else {
                        const unsigned __int128 c4 = (unsigned __int128)(4ull * npairs) * SC;
                        const double cnum = (c4 >= sa2) ? (double)(c4 - sa2) : -(double)(sa2 - c4);
                        pr[5] = cnum / (double)vnum;
                    }
                }
            }
            __syncthreads();
            if (tid < nbc * kTexFields) {
                const int b = tid / kTexFields, f = tid % kTexFields;
                double v = NAN_D;
                if ((any_bands >> b) & 1u)
                    v = (((S.props[b][0][f] + S.props[b][1][f]) + S.props[b][2][f]) + S.props[b][3][f]) / 4.0;
                o[(k0 + b) * kTexFields + f] = v;
            }
            __syncthreads();   // S.* and the tiles are reused by the next chunk / segment
        }
    }
}

template <bool WIDE, int NT> static size_t texture_smem_bytes()
{
    const size_t hist_words = WIDE ? kBins : (kBins + 1) / 2;
    return hist_words * 4 + sizeof(TexSmem) + (WIDE ? 16 : (size_t)kTileBytes + (size_t)kTileMax * 2);
}

}  // namespace obia

using namespace obia;

extern "C" int64_t obia_b200_texture_workspace_bytes(int64_t max_label)
{
    if (max_label < 0) return -1;
    return zonal_ws_layout(nullptr, max_label).bytes;
}

extern "C" int obia_b200_texture_stats(const int32_t *labels, const float *raw, int64_t H, int64_t W, int32_t C,
                                       const int32_t *bands_host, int32_t n_bands, int64_t max_label,
                                       int32_t quantise_f64, double *features, void *workspace, void *stream)
{
    if (!labels || !raw || !bands_host || !features || !workspace || H <= 0 || W <= 0 || C <= 0 ||
        n_bands <= 0 || max_label < 0)
        return set_err(OBIA_B200_ERR_ARG, "texture_stats: bad argument");
    if (n_bands > OBIA_B200_MAX_BANDS)
        return set_err(OBIA_B200_ERR_UNSUPPORTED, "texture_stats: at most %d bands per call", OBIA_B200_MAX_BANDS);
    if (H * W >= 0x7fffffffLL || max_label >= 0x7fffffffLL)
        return set_err(OBIA_B200_ERR_UNSUPPORTED, "texture_stats: H*W exceeds int32");
    TexBands tb;
    memset(&tb, 0, sizeof(tb));
    for (int b = 0; b < n_bands; ++b) {
        if (bands_host[b] < 0 || bands_host[b] >= C)
            return set_err(OBIA_B200_ERR_ARG, "texture_stats: band %d out of range", bands_host[b]);
        tb.band[b] = bands_host[b];
    }
    cudaStream_t st = (cudaStream_t)stream;
    ZonalWs w = zonal_ws_layout(workspace, max_label);
    const int64_t n = max_label + 1;
    zonal_init_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(w, n);
    OBIA_LAUNCH_CHECK();
    zonal_bbox_launch(labels, w, H, W, max_label, 0, 0, st);
    OBIA_LAUNCH_CHECK();

    int dev = 0, sms = 0;
    OBIA_CUDA_CHECK(cudaGetDevice(&dev));
    OBIA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    constexpr int NT_S = 256, NT_W = 512;
    const size_t sm_s = texture_smem_bytes<false, NT_S>(), sm_w = texture_smem_bytes<true, NT_W>();
    OBIA_CUDA_CHECK(cudaFuncSetAttribute(texture_kernel<false, NT_S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_s));
    OBIA_CUDA_CHECK(cudaFuncSetAttribute(texture_kernel<true, NT_W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_w));
    // persistent CTAs: as many as fit per SM (shared-memory bound: 2 with 16-bit counters, 1 with 32-bit)
    int occ_s = 1, occ_w = 1;
    OBIA_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, texture_kernel<false, NT_S>, NT_S, sm_s));
    OBIA_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_w, texture_kernel<true, NT_W>, NT_W, sm_w));
    const int64_t cap_s = (int64_t)sms * (occ_s > 0 ? occ_s : 1), cap_w = (int64_t)sms * (occ_w > 0 ? occ_w : 1);
    const unsigned g_s = (unsigned)(n < cap_s ? n : cap_s);
    const unsigned g_w = (unsigned)(n < cap_w ? n : cap_w);
    texture_kernel<false, NT_S><<<g_s, NT_S, sm_s, st>>>(labels, raw, w, (int)W, C, tb, n_bands, max_label,
                                                         quantise_f64, features);
    OBIA_LAUNCH_CHECK();
    texture_kernel<true, NT_W><<<g_w, NT_W, sm_w, st>>>(labels, raw, w, (int)W, C, tb, n_bands, max_label,
                                                        quantise_f64, features);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

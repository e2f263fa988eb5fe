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
// One CTA per segment; 16-bit counters (3 CTAs / SM) for bounding boxes below 65 536 pixels, a
// second launch with 32-bit counters (1 CTA / SM) for the rare larger ones.
#include "common.cuh"
#include "bbox.cuh"

namespace obia {

constexpr int kBins = 256 * 257 / 2;   // unordered level pairs
constexpr int kTile = 4096;            // crop pixels staged in shared memory as uint8
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

template <bool WIDE, int NT>
__global__ void __launch_bounds__(NT)
texture_kernel(const int32_t *__restrict__ labels, const float *__restrict__ raw, ZonalWs bb, int W, int C,
               TexBands tb, int nb, int64_t max_label, int f64, double *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int kHistWords = WIDE ? kBins : (kBins + 1) / 2;
    constexpr int NW = NT / 32;
    unsigned *hist = reinterpret_cast<unsigned *>(smem);
    double *lut = reinterpret_cast<double *>(smem + (size_t)kHistWords * 4);
    unsigned long long *red = reinterpret_cast<unsigned long long *>(lut + 256);   // [NW][8]
    double *redd = reinterpret_cast<double *>(red + NW * 8);                         // [NW]
    float *redf = reinterpret_cast<float *>(redd + NW);                              // [NW][2] + flags
    unsigned char *tile = reinterpret_cast<unsigned char *>(redf + NW * 4);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kHistWords; i += NT) hist[i] = 0u;
    for (int i = tid; i < 256; i += NT) lut[i] = 1.0 / (1.0 + (double)i * (double)i);
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
        const bool staged = !WIDE && area <= kTile;
        const int lab = (int)L;

        for (int k = 0; k < nb; ++k) {
            const int band = tb.band[k];
            // sample of the crop at (r, c): the raster value inside the segment, 0 outside / NaN
            auto sample = [&](int r, int c, bool &valid) -> float {
                const int64_t p = (int64_t)(y0 + r) * W + (x0 + c);
                float v = 0.0f;
                valid = false;
                if (labels[p] == lab) {
                    v = raw[p * C + band];
                    valid = (v == v);
                    if (!valid) v = 0.0f;
                }
                return v;
            };
            // ---- min / max of the crop --------------------------------------------------------
            float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
            int any = 0;
            for (int64_t p = tid; p < area; p += NT) {
                bool valid;
                const float v = sample((int)(p / w), (int)(p % w), valid);
                mn = fminf(mn, v);
                mx = fmaxf(mx, v);
                any |= valid ? 1 : 0;
            }
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
                mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, s));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
                any |= __shfl_xor_sync(0xffffffffu, any, s);
            }
            if (lane == 0) {
                redf[warp * 4] = mn;
                redf[warp * 4 + 1] = mx;
                redf[warp * 4 + 2] = __int_as_float(any);
            }
            __syncthreads();
            any = 0;
            for (int i = 0; i < NW; ++i) {
                mn = fminf(mn, redf[i * 4]);
                mx = fmaxf(mx, redf[i * 4 + 1]);
                any |= __float_as_int(redf[i * 4 + 2]);
            }
            __syncthreads();   // redf is reused by the next band
            if (!any) {        // no valid sample: every feature is NaN (:217-231)
                for (int i = tid; i < kTexFields; i += NT) o[k * kTexFields + i] = NAN_D;
                continue;
            }
            const bool flat = (mx == mn);
            const float span32 = __fsub_rn(mx, mn);
            const double span64 = (double)mx - (double)mn;
            // the reference's arithmetic in the dtype of its masked crop (float32 rasters stay
            // float32, integer rasters become float64), truncated by astype(uint8)
            auto quantise = [&](float v) -> int {
                if (flat) return 0;
                if (f64) return (int)(((double)v - (double)mn) / span64 * 255.0);
                return (int)__fmul_rn(__fdiv_rn(__fsub_rn(v, mn), span32), 255.0f);
            };
            if (staged) {
                for (int p = tid; p < (int)area; p += NT) {
                    bool valid;
                    tile[p] = (unsigned char)quantise(sample(p / w, p % w, valid));
                }
                __syncthreads();
            }
            auto level = [&](int r, int c) -> int {
                if (staged) return tile[r * w + c];
                bool valid;
                return quantise(sample(r, c, valid));
            };

            double feat[kTexFields] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};   // sums over the angles (thread 0)
#pragma unroll 1
            for (int a = 0; a < 4; ++a) {
                const int dr = (a == 0) ? 0 : (a == 2) ? 2 : 1;
                const int dc = (a == 0) ? 2 : (a == 1) ? 1 : (a == 2) ? 0 : -1;
                const int nr = h - dr, ncw = w - (dc < 0 ? -dc : dc), c_lo = dc < 0 ? -dc : 0;
                const int64_t npairs = (nr > 0 && ncw > 0) ? (int64_t)nr * ncw : 0;
                unsigned long long s1 = 0, s2 = 0, sa = 0, sb = 0, sc = 0, sdiag = 0, soff = 0;
                double sh = 0.0;
                for (int64_t t = tid; t < npairs; t += NT) {
                    const int r = (int)(t / ncw), c = c_lo + (int)(t % ncw);
                    const int i = level(r, c), j = level(r + dr, c + dc);
                    const int d = i > j ? i - j : j - i;
                    s1 += (unsigned)d;
                    s2 += (unsigned)(d * d);
                    sh += lut[d];
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
                    }
                    if (d == 0)
                        sdiag += 2ull * old + 1ull;
                    else
                        soff += 2ull * old + 1ull;
                }
                s1 = warp_sum_u64(s1);
                s2 = warp_sum_u64(s2);
                sa = warp_sum_u64(sa);
                sb = warp_sum_u64(sb);
                sc = warp_sum_u64(sc);
                sdiag = warp_sum_u64(sdiag);
                soff = warp_sum_u64(soff);
#pragma unroll
                for (int s = 16; s >= 1; s >>= 1) sh += __shfl_xor_sync(0xffffffffu, sh, s);
                if (lane == 0) {
                    unsigned long long *rw = red + warp * 8;
                    rw[0] = s1; rw[1] = s2; rw[2] = sa; rw[3] = sb; rw[4] = sc; rw[5] = sdiag; rw[6] = soff;
                    redd[warp] = sh;
                }
                __syncthreads();   // sums published; every increment of this angle has landed
                // clear the histogram by replaying the pairs (all non-zero words were touched)
                for (int64_t t = tid; t < npairs; t += NT) {
                    const int r = (int)(t / ncw), c = c_lo + (int)(t % ncw);
                    const int i = level(r, c), j = level(r + dr, c + dc);
                    const int lo = min(i, j), hi = max(i, j);
                    const int bin = hi * (hi + 1) / 2 + lo;
                    hist[WIDE ? bin : (bin >> 1)] = 0u;
                }
                if (tid == 0) {
                    unsigned long long S1 = 0, S2 = 0, SA = 0, SB = 0, SC = 0, SD = 0, SO = 0;
                    double SH = 0.0;
                    for (int i = 0; i < NW; ++i) {
                        const unsigned long long *rw = red + i * 8;
                        S1 += rw[0]; S2 += rw[1]; SA += rw[2]; SB += rw[3]; SC += rw[4]; SD += rw[5]; SO += rw[6];
                        SH += redd[i];
                    }
                    if (npairs == 0) {
                        feat[5] += 1.0;   // empty matrix: all sums 0, correlation 1 (std < 1e-15)
                    } else {
                        const double N = (double)npairs;
                        const double asm_ = ((double)SO + 2.0 * (double)SD) / (2.0 * N * N);
                        feat[0] += (double)S2 / N;
                        feat[1] += (double)S1 / N;
                        feat[2] += SH / N;
                        feat[3] += asm_;
                        feat[4] += sqrt(asm_);
                        // var = (2N SB - SA^2) / 4N^2, cov = (4N SC - SA^2) / 4N^2, exact in 128 bits
                        const unsigned __int128 sa2 = (unsigned __int128)SA * SA;
                        const unsigned __int128 vnum = (unsigned __int128)(2ull * (unsigned long long)npairs) * SB - sa2;
                        if (vnum == 0) {
                            feat[5] += 1.0;
                        } else {
                            const unsigned __int128 c4 = (unsigned __int128)(4ull * (unsigned long long)npairs) * SC;
                            const double cnum = (c4 >= sa2) ? (double)(c4 - sa2) : -(double)(sa2 - c4);
                            feat[5] += cnum / (double)vnum;
                        }
                    }
                }
                __syncthreads();   // histogram clean, reduction scratch free
            }
            if (tid == 0) {
#pragma unroll
                for (int i = 0; i < kTexFields; ++i) o[k * kTexFields + i] = feat[i] / 4.0;
            }
        }
    }
}

template <bool WIDE, int NT> static size_t texture_smem_bytes()
{
    const size_t hist_words = WIDE ? kBins : (kBins + 1) / 2;
    const size_t nw = NT / 32;
    return hist_words * 4 + 256 * 8 + nw * 8 * 8 + nw * 8 + nw * 4 * 4 + (WIDE ? 16 : kTile);
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
    const int64_t n = max_label + 1, N = H * W;
    zonal_init_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(w, n);
    OBIA_LAUNCH_CHECK();
    zonal_bbox_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(labels, w, N, (int)W, max_label);
    OBIA_LAUNCH_CHECK();

    int dev = 0, sms = 0;
    OBIA_CUDA_CHECK(cudaGetDevice(&dev));
    OBIA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    constexpr int NT_S = 256, NT_W = 512;
    const size_t sm_s = texture_smem_bytes<false, NT_S>(), sm_w = texture_smem_bytes<true, NT_W>();
    OBIA_CUDA_CHECK(cudaFuncSetAttribute(texture_kernel<false, NT_S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_s));
    OBIA_CUDA_CHECK(cudaFuncSetAttribute(texture_kernel<true, NT_W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_w));
    // persistent CTAs: 3 per SM with 16-bit counters, 1 per SM with 32-bit counters
    const unsigned g_s = (unsigned)(n < (int64_t)sms * 3 ? n : (int64_t)sms * 3);
    const unsigned g_w = (unsigned)(n < (int64_t)sms ? n : (int64_t)sms);
    texture_kernel<false, NT_S><<<g_s, NT_S, sm_s, st>>>(labels, raw, w, (int)W, C, tb, n_bands, max_label,
                                                         quantise_f64, features);
    OBIA_LAUNCH_CHECK();
    texture_kernel<true, NT_W><<<g_w, NT_W, sm_w, st>>>(labels, raw, w, (int)W, C, tb, n_bands, max_label,
                                                        quantise_f64, features);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

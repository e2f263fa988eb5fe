/*
 * obia_b200.h -- C ABI of the B200-native SLIC + zonal-statistics hot path.
 *
 * The reference (iosefa/obia) is pure Python and has no FFI of its own; the
 * native work on this path is done by third-party wheels it calls.  Each entry
 * point below names the reference call site (file:line under /root/reference)
 * and the third-party routine it replaces.  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`
 *   - the caller allocates and owns all buffers (sizes given per function;
 *     `obia_b200_*_workspace_bytes` for scratch); the library never allocates
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*), no
 *     hidden synchronisation unless stated
 *   - return 0 on success, negative on error; `obia_b200_last_error()` gives a
 *     thread-local message.  Nothing throws across the ABI.
 *   - rasters are row-major; `raw` is pixel-interleaved (H, W, C) float32
 *     exactly like `Image.img_data` (obia/handlers/geotif.py:100); SLIC
 *     features are band-planar [Cf][H][pitch] float32; labels are (H, W) int32.
 */
#ifndef OBIA_B200_H
#define OBIA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OBIA_B200_OK 0
#define OBIA_B200_ERR_ARG (-1)
#define OBIA_B200_ERR_CUDA (-2)
#define OBIA_B200_ERR_UNSUPPORTED (-3)

#define OBIA_B200_MAX_BANDS 64 /* selected bands per call (kernel-parameter table) */

const char *obia_b200_last_error(void);
int obia_b200_version(void);

/* Measurement hooks (bench.py): number of kernels this library has launched in
 * the process, and CUDA-event timing of the dominant kernel (SLIC
 * assign+update) on the stream it is launched on.  `profile_read` waits for the
 * recorded events, returns their summed duration and count, and clears them. */
int64_t obia_b200_launch_count(void);
int obia_b200_profile_enable(int on);
int obia_b200_profile_read(double *total_ms, int64_t *launches);

/* ---------------------------------------------------------------- K1 ----
 * Per-band min / max over a pixel-interleaved raster, plus the same over the
 * masked pixels only, plus a non-finite flag per band.
 * Replaces numpy's `np.min/np.max` in `normalize_band`
 * (obia/segmentation/segment_boundaries.py:11-16, applied to every band at
 * :31-33) and the `image_values.min()/max()` of skimage.segmentation.slic
 * reached from segment_boundaries.py:51.
 *   raw        [n_pixels][C] float32
 *   mask       [n_pixels] uint8 or NULL
 *   out        [C][4] float32 : min, max, masked min, masked max
 *   nonfinite  [C] int32      : bit0 = NaN seen, bit1 = +-inf seen
 */
int obia_b200_band_minmax(const float *raw, int64_t n_pixels, int32_t C,
                          const uint8_t *mask, float *out, int32_t *nonfinite,
                          void *stream);

/* In-place per-band min-max normalisation of the caller's raster: the side
 * effect of segment_boundaries.py:31-33 (`img_data[:,:,i] = normalize_band`).
 *   minmax : the [C][4] device array written by obia_b200_band_minmax
 */
int obia_b200_normalize_inplace(float *raw, int64_t n_pixels, int32_t C,
                                const float *minmax, void *stream);

/* Same arithmetic, written to another buffer (the raw raster stays intact for
 * the statistics).  `out` may be page-locked HOST memory that the device can
 * address (cudaHostAlloc / cudaHostRegister under UVA): the kernel then streams
 * the normalised raster straight into the caller's `img_data` over PCIe, with
 * no staging copy and without occupying the copy engine the small read-backs
 * of the SLIC path use.  max_ctas > 0 caps the grid (a handful of CTAs
 * saturates PCIe and leaves the SMs to the SLIC kernels running concurrently).
 */
int obia_b200_normalize_to(const float *raw, float *out, int64_t n_pixels,
                           int32_t C, const float *minmax, int32_t max_ctas,
                           void *stream);

/* Fused band select + obia min-max normalise + skimage global rescale +
 * optional RGB->CIELAB + multiply by 1/compactness, written band-planar.
 * Replaces segment_boundaries.py:35-43 and, inside skimage.segmentation.slic
 * (segment_boundaries.py:51): `image -= imin; image /= (imax-imin)`,
 * `rgb2lab`, `image * ratio`.
 *   bands_host  [Cs] selected band indices; Cs <= OBIA_B200_MAX_BANDS
 *   features    [Cf][H][pitch] float32, Cf = 3 if to_lab else Cs
 *   pitch       elements per feature row, multiple of 4, >= W
 *   ratio       float32(1/compactness); pass 1.0f when a Gaussian pass follows
 */
int obia_b200_slic_features(const float *raw, int64_t H, int64_t W, int32_t C,
                            const int32_t *bands_host, int32_t Cs,
                            const float *band_min_host,
                            const float *band_max_host, float imin, float imax,
                            int32_t to_lab, float ratio, float *features,
                            int64_t pitch, void *stream);

/* Separable Gaussian pre-smoothing of planar features (skimage's `sigma>0`
 * branch: skimage.filters.gaussian -> scipy.ndimage.gaussian_filter,
 * mode='reflect', truncate=4; float64 accumulation, float32 storage between
 * the row and column passes, like scipy).  Output is multiplied by `ratio`.
 *   weights_host [2*radius+1] float64 normalised kernel (host)
 *   tmp, out     [Cf][H][pitch] float32 (tmp may not alias in or out)
 * With out != in and radii <= 63 (sigma <= 15.8) both passes run in ONE shared-memory tiled kernel,
 * the taps travel as kernel parameters and the call is asynchronous (tmp is not touched).  Wider
 * kernels, or out == in, take two global passes with the taps in constant memory; that path
 * synchronises the stream once.
 */
int obia_b200_gaussian_planar(const float *in, float *tmp, float *out,
                              int64_t H, int64_t W, int64_t pitch, int32_t Cf,
                              const double *weights_y_host, int32_t radius_y,
                              const double *weights_x_host, int32_t radius_x,
                              float ratio, void *stream);

/* maskSLIC initialisation, the k-means part: replaces
 * `scipy.cluster.vq.kmeans2(coord[idx_dense], coord[idx], iter=5)` inside skimage's
 * `_get_mask_centroids` (reached from segment_boundaries.py:51 when `mask=` is passed, i.e. for
 * every tile of obia/utils/tiling.py:137-143 and :275-281).  Bit-identical to scipy for pixel
 * coordinates (float64 distances in feature order, lowest index wins ties, exact integer sums).
 *   points_yx    [m][2] int32 (y, x) of the dense sample, device
 *   centroids_yx [n][2] float64 in/out, device
 *   extent_y/x   all coordinates lie in [0, extent): the raster (tile) shape
 * More than 1024 centroids: the nearest centroid is found through a uniform
 * grid over the centroids, visited ring by ring with an exact stopping bound,
 * so the assignment (ties included) is the brute-force one at O(1) distance
 * evaluations per point instead of n.
 */
int64_t obia_b200_mask_kmeans_workspace_bytes(int64_t n, int64_t extent_y,
                                              int64_t extent_x);
int obia_b200_mask_kmeans(const int32_t *points_yx, int64_t m, double *centroids_yx,
                          int64_t n, int32_t iters, int64_t extent_y,
                          int64_t extent_x, void *workspace, void *stream);

/* Nearest OTHER centroid of every centroid: skimage's
 * `dist = squareform(pdist(centroids)); np.fill_diagonal(dist, np.inf);
 * closest = dist.argmin(-1)` in `_get_mask_centroids` (Euclidean float64
 * distances, lowest index among equal minima), without the n x n matrix.
 *   closest   [n] int32 out, device
 *   workspace obia_b200_mask_kmeans_workspace_bytes(n, extent_y, extent_x)
 */
int obia_b200_nearest_centroid(const double *centroids_yx, int64_t n,
                               int64_t extent_y, int64_t extent_x,
                               int32_t *closest, void *workspace, void *stream);

/* HOST helper (no CUDA): the two `RandomState(123)` sample draws of skimage's
 * `_get_mask_centroids` -- `np.sort(rng.choice(np.arange(n_coord), k, replace=False))` for
 * k = min(n_segments, n_coord) and then k = min(100 * n_segments, n_coord) -- numpy's legacy
 * MT19937 + masked-rejection Fisher-Yates restated bit for bit.  Thread-safe and GIL-free, so the
 * tiled driver runs it for many tiles on host threads while the GPU segments earlier tiles.
 *   idx        [min(n_segments, n_coord)]       int64 out (host)
 *   idx_dense  [min(100 * n_segments, n_coord)] int64 out (host)
 */
int obia_b200_mask_sample_indices(int64_t n_coord, int64_t n_segments, int64_t *idx,
                                  int64_t *idx_dense);

/* ---------------------------------------------------------------- K2 ----
 * SLIC iterations: replaces Cython `_slic_cython`
 * (skimage/segmentation/_slic.pyx) reached from segment_boundaries.py:51.
 * Semantics kept: +-2*step scatter window per centre with C truncation,
 * float32 distance in the reference's operation order (no FMA contraction),
 * strict-less update with the lowest centre index winning exact ties, exactly
 * `max_num_iter` iterations, empty centre -> NaN centre that never wins.
 * Centre sums are accumulated in 64-bit fixed point so the result does not
 * depend on thread scheduling.
 *
 *   features  [Cf][H][pitch] float32
 *   mask      [H][W] uint8 or NULL
 *   centres   [n][2+Cf] float32 rows (cy, cx, colour...) in/out
 *   labels    [H][W] int32 out (pre-filled by the call with start_label-1)
 *   workspace obia_b200_slic_workspace_bytes(...) bytes
 *   step      float: max of the init steps (spatial weight 1/step^2)
 *   step_y/x  integer window half sizes (regular_grid steps)
 *   fix_scale power-of-two scale of the fixed-point colour sums; must satisfy
 *             max|feature| * fix_scale * min(H*W, (4*step_y+1)*(4*step_x+1)) <= 2^62
 *   status    [4] int32 device words (zeroed by the call): [0] != 0 when a
 *             tile's candidate list overflowed (result invalid)
 */
int64_t obia_b200_slic_workspace_bytes(int64_t H, int64_t W, int32_t Cf,
                                       int64_t n, int32_t step_y,
                                       int32_t step_x);
int obia_b200_slic_iterate(const float *features, const uint8_t *mask,
                           float *centres, int32_t *labels, void *workspace,
                           int64_t H, int64_t W, int64_t pitch, int32_t Cf,
                           int64_t n, float step, int32_t step_y,
                           int32_t step_x, int32_t max_num_iter,
                           int32_t start_label, int32_t ignore_color,
                           int32_t slic_zero, double fix_scale, int32_t *status,
                           void *stream);
/* obia_b200_slic_iterate with skimage's `spacing=(spacing_y, spacing_x)` (obia forwards **kwargs to
 * skimage.segmentation.slic, segment_boundaries.py:51): the spatial term of `_slic_cython` becomes
 * ((sy * (cy - y))^2 + (sx * (cx - x))^2) / step^2; windows, grid and connectivity stay in pixel units.
 * Exact kernel only; (1, 1) is bit-identical to obia_b200_slic_iterate. */
int obia_b200_slic_iterate_spacing(const float *features, const uint8_t *mask,
                                   float *centres, int32_t *labels, void *workspace,
                                   int64_t H, int64_t W, int64_t pitch, int32_t Cf,
                                   int64_t n, float step, int32_t step_y,
                                   int32_t step_x, int32_t max_num_iter,
                                   int32_t start_label, int32_t ignore_color,
                                   int32_t slic_zero, double fix_scale, float spacing_y,
                                   float spacing_x, int32_t *status, void *stream);

/* The same iteration, one sweep at a time, for a raster sharded by ROW STRIPS across GPUs (global
 * multi-GPU SLIC): `features`/`mask`/`labels` hold the strip [y_offset, y_offset + H) of a raster with
 * H_total rows; `centres` is the full (replicated) centre table.  A sweep leaves this strip's
 * contribution to the centre sums in the accumulator table at the START of `workspace`
 * ([n][3+Cf] int64: count, sum y, sum x, fixed-point colour sums).  The caller sums that table over
 * the ranks (ncclAllReduce on int64: exact, order-independent) and calls `finish_sweep`, so every rank
 * derives identical centres and the result is bit-identical to the single-GPU run.
 * `slic_iterate` == begin; { sweep(y_offset=0, H_total=H); finish_sweep } x max_num_iter.
 * workspace: obia_b200_slic_workspace_bytes(H_total, W, Cf, n, step_y, step_x). */
int obia_b200_slic_begin(int32_t *labels, void *workspace, int64_t H, int64_t W,
                         int64_t H_total, int32_t Cf, int64_t n, int32_t step_y,
                         int32_t step_x, int32_t start_label, int32_t *status, void *stream);
int obia_b200_slic_sweep(const float *features, const uint8_t *mask, const float *centres,
                         int32_t *labels, void *workspace, int64_t H, int64_t W, int64_t pitch,
                         int32_t Cf, int64_t n, float step, int32_t step_y, int32_t step_x,
                         int32_t start_label, int32_t ignore_color, int32_t slic_zero,
                         double fix_scale, int64_t y_offset, int64_t H_total, int32_t *status,
                         void *stream);
int obia_b200_slic_finish_sweep(float *centres, void *workspace, int64_t H_total, int64_t W,
                                int32_t Cf, int64_t n, int32_t step_y, int32_t step_x,
                                double fix_scale, void *stream);
/* Tolerance mode of the two calls above (same arguments, same workspace, same fused deterministic
 * centre update): the squared distance is expanded so that one fused multiply-add per channel
 * ranks a candidate centre (see csrc/slic_fast.cu); every tile works in local coordinates and
 * colours relative to one of its candidate centres, so the float32 result differs from the
 * reference's operation order only between candidates that are within rounding of each other.
 * This is the mode north_star's ">= 99.5 % label agreement" tolerance allows and the default of
 * the Python host (`exact=False`); the exact calls above reproduce `_slic_cython` bit for bit
 * given the centres.  slic_zero != 0 falls through to the exact kernel.  For a sharded run to be
 * bit-identical to the single-GPU one, y_offset must be a multiple of 64 (tile height).
 * `obia_b200_slic_fast_variant(warps)` selects the CTA shape (8 or 4 warps; tuning knob). */
int obia_b200_slic_iterate_fast(const float *features, const uint8_t *mask,
                                float *centres, int32_t *labels, void *workspace,
                                int64_t H, int64_t W, int64_t pitch, int32_t Cf,
                                int64_t n, float step, int32_t step_y,
                                int32_t step_x, int32_t max_num_iter,
                                int32_t start_label, int32_t ignore_color,
                                int32_t slic_zero, double fix_scale, int32_t *status,
                                void *stream);
int obia_b200_slic_sweep_fast(const float *features, const uint8_t *mask, const float *centres,
                              int32_t *labels, void *workspace, int64_t H, int64_t W, int64_t pitch,
                              int32_t Cf, int64_t n, float step, int32_t step_y, int32_t step_x,
                              int32_t start_label, int32_t ignore_color, int32_t slic_zero,
                              double fix_scale, int64_t y_offset, int64_t H_total, int32_t *status,
                              void *stream);
int obia_b200_slic_fast_variant(int32_t warps_per_cta);

/* Sharded runs may exchange only the BANDS of the accumulator table whose centres can receive pixels
 * from two ranks (centre indices [up_lo, up_hi) at the upper strip boundary, [down_lo, down_hi) at the
 * lower one) instead of all-reducing the whole table.  After finish_sweep this check raises status[1]
 * when a live centre outside those bands has a window that reaches beyond rows [row_lo, row_hi) of
 * this rank (a centre drifted further than the band allows): the caller then repeats the run with
 * the full all-reduce.  Asynchronous. */
int obia_b200_slic_band_check(const float *centres, int64_t n, int32_t Cf, int64_t row_lo, int64_t row_hi,
                              int64_t H_total, int64_t up_lo, int64_t up_hi, int64_t down_lo,
                              int64_t down_hi, int32_t step_y, int32_t *status, void *stream);

/* slic_zero (SLICO, `_slic.pyx` "update the color distance maxima"): with slic_zero != 0 a sweep
 * divides every colour distance by the centre's running maximum (initialised to 1 by slic_begin);
 * after finish_sweep this call raises each centre's maximum to the largest colour distance of the
 * pixels currently assigned to it.  Sharded runs take the element-wise maximum of the float table
 * (the last n floats of the workspace) over the strips. */
int obia_b200_slic_update_max_color(const float *features, const uint8_t *mask, const float *centres,
                                    const int32_t *labels, void *workspace, int64_t H, int64_t W,
                                    int64_t H_total, int64_t pitch, int32_t Cf, int64_t n,
                                    int32_t step_y, int32_t step_x, int32_t start_label, void *stream);

/* method="quickshift": replaces `skimage.segmentation.quickshift(img_to_segment, **kwargs)`
 * (obia/segmentation/segment_boundaries.py:48-49; scikit-image's `_quickshift_cython`): density of every
 * pixel over its clipped (2w+1)^2 window, w = ceil(3 * kernel_size) (float32 distances, double exp, float32
 * running sum, row-major order), + `noise`; parent = window pixel of strictly higher density at the smallest
 * distance, cut at `max_dist`; labels = raster-order rank of every pixel's root, 0..n-1.
 *   features  [Cf][H][pitch] float32 from obia_b200_slic_features (imin = 0, imax = 1: no global rescale;
 *             to_lab = convert2lab; ratio = quickshift's `ratio`), smoothed first when sigma > 0
 *   noise     [H][W] float64: numpy `default_rng(rng).normal(scale=1e-5, size=(H, W))`, drawn on the host
 *   labels    [H][W] int32 out
 *   n_labels_host  optional out (synchronises the stream when given)
 * workspace: obia_b200_quickshift_workspace_bytes(H, W). */
int64_t obia_b200_quickshift_workspace_bytes(int64_t H, int64_t W);
int obia_b200_quickshift(const float *features, const double *noise, int32_t *labels, void *workspace,
                         int64_t H, int64_t W, int64_t pitch, int32_t Cf, float kernel_size, float max_dist,
                         int64_t *n_labels_host, void *stream);

/* ---------------------------------------------------------------- K3 ----
 * Enforce connectivity: replaces Cython `_enforce_label_connectivity_cython`
 * (sequential raster-scan BFS) reached from segment_boundaries.py:51.
 * Union-find connected components + exact data-parallel replay of the
 * reference's small-segment merge, BFS size cap and raster-order numbering
 * (see DESIGN.md, K3).  Synchronises the stream once at the end (the read-back
 * of the segment count; more only when the rare label-0 chains need extra rounds).
 *   labels_in  [H][W] int32 (values start_label-1 = masked)
 *   labels_out [H][W] int32
 *   n_labels_host  out: number of kept segments (labels are
 *                  start_label .. start_label+n-1, plus possibly 0)
 */
int64_t obia_b200_connectivity_workspace_bytes(int64_t H, int64_t W);
int obia_b200_enforce_connectivity(const int32_t *labels_in,
                                   int32_t *labels_out, void *workspace,
                                   int64_t H, int64_t W, int64_t min_size,
                                   int64_t max_size, int32_t start_label,
                                   int64_t *n_labels_host, void *stream);

/* The same, for ONE raster sharded by row strips across GPUs (no reference counterpart: the reference is
 * single-process; result = the rows the single-raster call above would give, bit for bit).
 * `labels_ext` holds the rank's core rows [core_row0, core_row0 + core_rows) preceded / followed by halo
 * rows of its neighbours' SLIC labels (H_ext rows in all; `top_open` / `bottom_open` = the halo edge is
 * not the raster edge).  SLIC components are bounded by the +-2*step windows, so with a halo of a few
 * window heights every piece that reaches the core is complete inside the strip.
 *   strip_begin   components, size-cap split, small-piece merge targets; returns on the host
 *                 counts[0] = kept pieces that start above the core rows (inside the strip),
 *                 counts[1] = kept pieces that start in the core rows, counts[2] = all kept pieces,
 *                 counts[3] = rounds of the small-piece fixed point that were launched (diagnostic),
 *                 counts[4] = kept pieces that start above the core rows but below the outer third of the
 *                 upper halo, which the kernel books unknown (the rows of the statistics table a rank
 *                 shares with the rank above).  `counts_host` holds 5 values.
 *                 The caller all-gathers counts[1] over the ranks; `label_offset` of this rank is the
 *                 exclusive prefix sum minus its counts[0] (north_star: "exclusive scan of per-tile
 *                 label offsets").
 *   strip_finish  numbers the pieces (start_label + label_offset + local rank) and writes
 *                 the core rows.  `incomplete_host[0]` != 0 when a core pixel's label depends on pixels
 *                 outside the strip (cut component or unknown merge chain): the caller retries with a
 *                 taller halo (or gathers the whole raster); `incomplete_host[1]` != 0 when the strip
 *                 holds a merged piece without an earlier neighbour: it carries label 0 whatever the
 *                 rank (SURVEY.md defect 7), outside the rank's label range.  Both calls synchronise
 *                 the stream.
 * workspace: obia_b200_connectivity_workspace_bytes(H_ext, W), unchanged between the two calls. */
int obia_b200_connectivity_strip_begin(const int32_t *labels_ext, void *workspace, int64_t H_ext,
                                       int64_t W, int64_t core_row0, int64_t core_rows,
                                       int32_t top_open, int32_t bottom_open, int64_t min_size,
                                       int64_t max_size, int32_t start_label, int64_t *counts_host,
                                       void *stream);
int obia_b200_connectivity_strip_finish(const int32_t *labels_ext, int32_t *labels_out_core,
                                        void *workspace, int64_t H_ext, int64_t W, int64_t core_row0,
                                        int64_t core_rows, int64_t min_size, int64_t max_size,
                                        int32_t start_label, int64_t label_offset,
                                        int32_t *incomplete_host, void *stream);

/* ---------------------------------------------------------------- K4 ----
 * Per-segment, per-band zonal statistics in one pass over the raster:
 * replaces the per-segment loop of `create_objects`
 * (obia/segmentation/segment_statistics.py:475-508): crop_image_to_bbox +
 * mask_image_with_polygon (obia/utils/utils.py:37-67) +
 * calculate_spectral_stats (segment_statistics.py:113-176: np.mean, np.var,
 * np.min, np.max, scipy.stats.skew, scipy.stats.kurtosis).
 *   labels     [H][W] int32; pixels with label < 0 or > max_label are skipped
 *   raw        [H][W][C] float32
 *   bands_host [Cz] band indices
 *   resolution np.finfo(dtype).resolution of the dtype the reference would
 *              compute in (1e-6 float32 rasters, 1e-15 integer rasters):
 *              scipy returns NaN skew/kurtosis when m2 <= (resolution*mean)^2
 *   stats      [max_label+1][Cz][8] float64 out:
 *              count, mean, variance, min, max, skewness, kurtosis, sum
 *              NaN samples are dropped per band like `band_data[~np.isnan(band_data)]`
 *              (:144-147): count = valid samples of that band; count==0 -> NaN statistics
 */
int64_t obia_b200_zonal_workspace_bytes(int64_t max_label, int32_t Cz);
int obia_b200_zonal_stats(const int32_t *labels, const float *raw, int64_t H,
                          int64_t W, int32_t C, const int32_t *bands_host,
                          int32_t Cz, int64_t max_label, double resolution,
                          double *stats, void *workspace, void *stream);

/* Same statistics over the label range [label_lo, label_lo + n_rows): row r of `stats` = label
 * label_lo + r, every other label is skipped.  Used by the sharded path, where a rank's strip holds a
 * contiguous range of the raster-order label numbering.  zero_row = 1 (label_lo > 0): row 0 holds
 * label 0 -- merged pieces without an earlier neighbour carry it on any rank -- and row r >= 1 holds
 * label label_lo + r - 1 (n_rows counts the extra row).
 * workspace: obia_b200_zonal_workspace_bytes(n_rows - 1, Cz). */
int obia_b200_zonal_stats_range(const int32_t *labels, const float *raw, int64_t H,
                                int64_t W, int32_t C, const int32_t *bands_host,
                                int32_t Cz, int64_t label_lo, int64_t n_rows, int32_t zero_row,
                                double resolution, double *stats, void *workspace,
                                void *stream);

/* Polygons -> label raster, pixel-centre rule: replaces the per-segment
 * `rasterio.features.geometry_mask([polygon], transform, invert=True)` of
 * obia/utils/utils.py:53-67 (create_objects, segment_statistics.py:479-484) and
 * `rasterio.features.rasterize` of obia/utils/tiling.py:248-255.  All pointers are device pointers.
 *   xy              [n_vertices][2] float64 vertex (x, y) in PIXEL coordinates (column, row; pixel
 *                   (r, c) covers [c, c+1) x [r, r+1)); rings need not repeat their first vertex
 *   ring_start      [n_rings + 1] first vertex of every ring
 *   poly_ring_start [n_polygons + 1] first ring of every polygon (exterior, holes, further parts:
 *                   even-odd over all of them)
 *   poly_label      [n_polygons] label written for the polygon's pixels (larger label wins overlaps)
 *   bbox            [n_polygons][4] int32 x0, y0, x1, y1 (inclusive, clipped to the raster)
 *   labels          [H][W] int32 in/out, pre-filled by the caller (-1 = no polygon) */
int obia_b200_rasterize_polygons(const double *xy, const int32_t *ring_start,
                                 const int32_t *poly_ring_start, const int32_t *poly_label,
                                 const int32_t *bbox, int64_t n_polygons, int32_t *labels,
                                 int64_t H, int64_t W, void *stream);

/* ---------------------------------------------------------------- K5 ----
 * Per-segment GLCM texture features: replaces calculate_textural_stats
 * (obia/segmentation/segment_statistics.py:179-298) as create_objects calls it
 * for every segment (:496-508).  Per (segment, band): the bounding-box crop
 * with pixels outside the segment (and NaN samples) set to 0 (:214-247),
 * min/max-scaled to uint8 over the whole crop (:251-258), skimage
 * graycomatrix(distances=[2], angles=[0,pi/4,pi/2,3pi/4], levels=256,
 * symmetric=True, normed=True) (:261-268), then the mean over the angles of
 * graycoprops contrast, dissimilarity, homogeneity, ASM, energy, correlation
 * (:285-296).  The band axis slip at :214 is fixed (band-first crop indexed by
 * band), see DESIGN.md.
 *   labels, raw, bands_host, max_label   as for obia_b200_zonal_stats
 *   quantise_f64  0: scale in float32 (float32 rasters); 1: in float64 (the
 *                 reference's masked crop of an integer raster is float64)
 *   features      [max_label+1][n_bands][6] float64 out, order as above;
 *                 labels without pixels / bands without a valid sample -> NaN
 */
int64_t obia_b200_texture_workspace_bytes(int64_t max_label);
int obia_b200_texture_stats(const int32_t *labels, const float *raw, int64_t H,
                            int64_t W, int32_t C, const int32_t *bands_host,
                            int32_t n_bands, int64_t max_label,
                            int32_t quantise_f64, double *features,
                            void *workspace, void *stream);

/* ------------------------------------------------------ tiled driver ----
 * `create_tiled_segments` (obia/utils/tiling.py:62-291) with every window of a BATCH -- all black tiles of a
 * chunk (:103-153) or all white windows of one tile-row (:156-287), which are independent of each other --
 * handled by one launch per stage instead of one Python iteration per tile.  The windows of a batch are
 * stacked into a SLAB: window i occupies slab rows [i * win_rows, i * win_rows + h_i), columns [0, w_i);
 * win_rows = max h + 1 leaves one masked row between windows.  `descs` is a device array of B window
 * descriptors (136 bytes each, layout in obia_b200/csrc/batch.cuh == numpy dtype WIN_DESC in
 * obia_b200/batch.py).  Every entry is the batched form of the single-raster entry named beside it and gives,
 * inside each window, exactly that entry's result on the window alone. */

/* obia_b200_band_minmax per window.  raw (H, Wl, C) interleaved float32; out (B, C, 4) = min, max, masked
 * min, masked max; nonfinite (B, C); mask_counts (B) = mask pixels per window (0 without mask_slab). */
int obia_b200_window_stats(const float *raw, int64_t Wl, int32_t C, const void *descs, int64_t B, int32_t hmax,
                           const uint8_t *mask_slab, int32_t slab_w, float *out, int32_t *nonfinite,
                           int32_t *mask_counts, void *stream);
/* window of a (H, Wm) uint8 mask raster -> slab (tiling.py:121-123: the mask tile of a black tile) */
int obia_b200_window_mask_copy(const uint8_t *mask, int64_t Wm, const void *descs, int64_t B, int32_t hmax,
                               int32_t wmax, uint8_t *mask_slab, int32_t slab_w, void *stream);
/* obia_b200_slic_features per window: bands (device [Cs]); band_min / band_diff (device [B][Cs]); the global
 * rescale (imin, idiff, rescale) per window from the descriptors; features (Cf, slab_rows, pitch). */
int obia_b200_window_features(const float *raw, int64_t Wl, int32_t C, const int32_t *bands, int32_t Cs,
                              const float *band_min, const float *band_diff, const void *descs, int64_t B,
                              int32_t hmax, int32_t wmax, int32_t to_lab, float ratio, float *features,
                              int64_t slab_rows, int64_t pitch, void *stream);
/* maskSLIC initialisation per window (obia_b200_mask_kmeans + obia_b200_nearest_centroid + the `steps` mean):
 * points_pos / seed_pos = slab pixel positions of coord[idx_dense] / coord[idx] (dense_mask_slab != NULL instead
 * of points_pos: the points of every window are all of its mask pixels, assigned tile by tile); cwin (n_total) =
 * window of every centroid.  Writes centroids (n_total, 2) float64, the SLIC centre rows (n_total, 2 + Cf) and, into the
 * descriptors, sw / inv_w (valid = 0 where the step is not positive: the reference's ValueError). */
int64_t obia_b200_mask_kmeans_batch_workspace_bytes(int64_t n_total, int64_t km_cells_total);
int obia_b200_mask_kmeans_batch(const int32_t *points_pos, int64_t m_total, const uint8_t *dense_mask_slab,
                                int32_t hmax, int32_t wmax, const int32_t *seed_pos, const int32_t *cwin, void *descs,
                                int64_t B, int64_t n_total, int64_t km_cells_total, int32_t slab_w, int32_t win_rows,
                                int32_t iters, int32_t Cf, double *centroids, float *centres, void *workspace,
                                void *stream);
/* obia_b200_slic_iterate_fast per window.  `prepare` fills the kernel-variant dependent fixed-point fields of
 * HOST descriptors before upload (`pad` = strip phases of the tile variant that serves the window; `variants` =
 * OR of it over the usable windows); status: (4 + B) int32, status[4 + i] != 0 = window i overflowed the
 * candidate staging (CandidateOverflowError of the single-raster path). */
int64_t obia_b200_slic_batch_workspace_bytes(int64_t n_total, int64_t cells_total, int32_t Cf);
int obia_b200_slic_batch_prepare(void *descs_host, int64_t B, int32_t Cf);
int obia_b200_slic_iterate_batch(const float *features, const uint8_t *mask, float *centres, int32_t *labels,
                                 void *workspace, const void *descs, const int32_t *cwin, int64_t B,
                                 int64_t n_total, int64_t cells_total, int32_t hmax, int32_t wmax,
                                 int64_t slab_rows, int32_t slab_w, int64_t pitch, int32_t Cf, int32_t max_num_iter,
                                 int32_t start_label, int32_t ignore_color, int32_t variants, int32_t *status,
                                 void *stream);
/* obia_b200_enforce_connectivity on the slab with (min_size, max_size) per window (device int32 pairs);
 * kept pieces are numbered start_label.. in slab raster order, i.e. window by window.
 * workspace: obia_b200_connectivity_workspace_bytes(H, W). */
int obia_b200_enforce_connectivity_windows(const int32_t *labels_in, int32_t *labels_out, void *workspace,
                                           int64_t H, int64_t W, const int32_t *window_sizes, int64_t win_rows,
                                           int32_t start_label, int64_t *n_labels_host, void *stream);
/* Segment bookkeeping over a global handle raster G (H, GW) int32, -1 = no segment (tiled.cu).
 * paint: slab labels of the usable windows -> handles handle_base + (label - start_label) (label 0 of
 * start_label 1 -> handle_zero + window; the -2 marker of start_label 0 -> first_label[window], the first kept
 * label of the window, device [B], < 0 when it kept none), pixel counts accumulated into sizes.
 * white_prepare: tiling.py:182-260 for every white window of a tile-row: segments entirely inside the window
 * polygon are erased from G (live = 0), segments straddling its edge and the two bottom corner squares are
 * masked out; the windows' SLIC masks are written into mask_slab. */
int obia_b200_tiled_paint(const int32_t *labels_slab, const uint8_t *mask_slab, int32_t slab_w, const void *descs,
                          const uint8_t *usable, const int32_t *first_label, int64_t B, int32_t hmax, int32_t wmax,
                          int32_t start_label, int32_t handle_base, int32_t handle_zero, int32_t *G, int64_t GW,
                          int32_t *sizes, void *stream);
int obia_b200_tiled_white_prepare(int32_t *G, int64_t GW, const void *descs, const uint8_t *parity, int64_t B,
                                  int32_t hmax, int32_t wmax, int32_t corner, int32_t *counts, int64_t capacity,
                                  const int32_t *sizes, uint8_t *live, int32_t *any_segment,
                                  const uint8_t *user_mask, int64_t Wm, uint8_t *mask_slab, int32_t slab_w,
                                  void *stream);

/* Seam exchange of the column-block sharded driver: import a neighbour's version of a boundary band (three
 * int64 planes per pixel: creation key or -1, segment size, home = owner rank << 32 | handle there) into the
 * rows x cols block of G at G_band.  Segments homed on this rank keep their handle, neighbour segments are
 * translated through `mirror` (neighbour handle -> local handle, -1 unknown; new ones take slots slot_base +
 * [0, rows * cols), reserved by the caller).  retire != 0 (white rows): every handle in the band that does not
 * come back is dead.  No host synchronisation. */
int obia_b200_tiled_seam_import(int32_t *G_band, int64_t GW, int32_t rows, int32_t cols, const int64_t *planes,
                                int32_t my_rank, int32_t retire, int32_t *mirror, int64_t mirror_cap,
                                int32_t slot_base, int32_t *counter, int32_t *err, int64_t *keys, int32_t *sizes,
                                uint8_t *live, int64_t *homes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* OBIA_B200_H */

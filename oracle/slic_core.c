/*
 * oracle/slic_core.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, never shipped).
 *
 * Plain-C restatement of the two single-threaded loops that do the SLIC
 * arithmetic behind obia's `create_segments`
 * (/root/reference/obia/segmentation/segment_boundaries.py:51 calls
 * `skimage.segmentation.slic`).  The arithmetic lives in scikit-image
 * (constraint `scikit-image>=0.23.2`, /root/reference/pyproject.toml:23, no
 * lock file), which is NOT vendored in /root/reference and NOT installable in
 * this image.  The algorithm below is therefore restated from the published
 * scikit-image sources (`skimage/segmentation/_slic.pyx`:
 * `_slic_cython` and `_enforce_label_connectivity_cython`), following
 * SURVEY.md section 3.4 steps 8 and 9.
 *
 * PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors
 * for this path and scikit-image cannot be run here, so
 * this restatement is checked only against recalled known-answer micro-cases
 * and structural properties (tests/test_oracle_*.py).
 *
 * Loop order, float32 arithmetic, comparison direction and truncation are
 * kept identical to the Cython loops.  Build with -ffp-contract=off so no FMA
 * contraction changes the float32 rounding (x86-64 scikit-image wheels are
 * built for the baseline ISA, which has no FMA).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* The file is compiled twice: REAL=float (what obia exercises: float32 rasters,
 * geotif.py:100) and REAL=double (only to check the recalled scikit-image
 * known-answer cases, which use float64 images). */
#ifndef REAL
#define REAL float
#define SUFFIX(n) n
#endif

/* Colour accumulation `dist_color += t * t`.
 * Default: separate multiply and add (x86-64 scikit-image wheels: baseline ISA,
 * no FMA).  With -DORACLE_FMA the statement is contracted to one fused
 * multiply-add, which is what compilers emit for it on targets whose baseline
 * has FMA (arm64 wheels: clang -ffp-contract=on / gcc -ffp-contract=fast).  The
 * CUDA kernel uses the fused form (packed FFMA2), so the `_fma` build is the
 * bit-exact twin of the GPU assignment; the default build is the x86-64 twin. */
#ifdef ORACLE_FMA
#define COLOR_ACC(acc, t) fmaf((t), (t), (acc))
#else
#define COLOR_ACC(acc, t) ((acc) + (t) * (t))
#endif

/* Cython's max()/min() lower to a ternary; NaN therefore propagates the way
 * `(b > a) ? b : a` does.  Keep that shape. */
static inline REAL cy_maxf(REAL a, REAL b) { return (b > a) ? b : a; }
static inline REAL cy_minf(REAL a, REAL b) { return (b < a) ? b : a; }

/* C cast of a float to a signed 64-bit integer; NaN / out-of-range give
 * INT64_MIN on x86-64 (cvttss2si), which makes the window empty. */
static inline int64_t cy_trunc(REAL v)
{
    if (!(v == v)) return INT64_MIN;
    if (v >= (REAL)9.2e18 || v <= (REAL)-9.2e18) return INT64_MIN;
    return (int64_t)v;
}

/*
 * _slic_cython restated for depth == 1 (obia only ever passes 2-D rasters:
 * segment_boundaries.py:43,51).
 *
 *   image     [H][W][C] float32, already scaled by 1/compactness
 *   mask      [H][W] uint8 or NULL
 *   segments  [n][3+C] float32, rows = (z, y, x, colour...) ; updated in place
 *   step      float (max of the init steps) -> spatial_weight = 1/step^2
 *   step_y/x  integer window half-sizes = regular_grid((1,H,W), n) steps
 *   nearest   [H][W] int64 out; must be pre-filled by the caller with
 *             start_label-1 (the Cython code allocates it that way)
 * Returns the number of iterations executed.
 */
int64_t SUFFIX(obia_oracle_slic_core)(const REAL *image, const uint8_t *mask,
                              REAL *segments, int64_t H, int64_t W, int64_t C,
                              int64_t n, float step, int64_t max_num_iter,
                              const REAL *spacing, int slic_zero,
                              int64_t start_label, int ignore_color,
                              int64_t step_y, int64_t step_x, int64_t *nearest)
{
    const int64_t nf = 3 + C;
    const REAL sy = spacing[1], sx = spacing[2];
    REAL *distance = (REAL *)malloc(sizeof(REAL) * (size_t)(H * W));
    int64_t *n_elems = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    REAL *max_dist_color = (REAL *)malloc(sizeof(REAL) * (size_t)n);
    if (!distance || !n_elems || !max_dist_color) return -1;
    for (int64_t k = 0; k < n; ++k) max_dist_color[k] = (REAL)1;

    /* `float step` in the Cython signature: float product, double division,
     * np_floats store */
    const float step_sq = step * step;
    const REAL spatial_weight = (REAL)(1.0 / (double)step_sq);

    int64_t it = 0;
    for (it = 0; it < max_num_iter; ++it) {
        int change = 0;
        for (int64_t i = 0; i < H * W; ++i) distance[i] = INFINITY; /* DBL_MAX -> REAL */

        for (int64_t k = 0; k < n; ++k) {
            const REAL *seg = segments + k * nf;
            const REAL cy = seg[1], cx = seg[2];
            /* z window is [0,1) for depth 1 and cz == 0 -> dz == 0 */
            const REAL cz = seg[0];
            const REAL tz = spacing[0] * (cz - (REAL)0);
            const REAL dz = tz * tz;
            const int64_t y_min = cy_trunc(cy_maxf(cy - (REAL)(2 * step_y), (REAL)0));
            const int64_t y_max = cy_trunc(cy_minf(cy + (REAL)(2 * step_y) + (REAL)1, (REAL)H));
            const int64_t x_min = cy_trunc(cy_maxf(cx - (REAL)(2 * step_x), (REAL)0));
            const int64_t x_max = cy_trunc(cy_minf(cx + (REAL)(2 * step_x) + (REAL)1, (REAL)W));
            if (cy_trunc(cz) == INT64_MIN) continue; /* NaN centre: empty z window */

            for (int64_t y = y_min; y < y_max; ++y) {
                const REAL ty = sy * (cy - (REAL)y);
                const REAL dy = ty * ty;
                for (int64_t x = x_min; x < x_max; ++x) {
                    if (mask && !mask[y * W + x]) continue;
                    const REAL tx = sx * (cx - (REAL)x);
                    const REAL dx = tx * tx;
                    REAL dist_center = (dz + dy + dx) * spatial_weight;
                    if (!ignore_color) {
                        REAL dist_color = (REAL)0;
                        const REAL *px = image + (y * W + x) * C;
                        for (int64_t c = 0; c < C; ++c) {
                            const REAL t = px[c] - seg[3 + c];
                            dist_color = COLOR_ACC(dist_color, t);
                        }
                        if (slic_zero) dist_color /= max_dist_color[k];
                        dist_center += dist_color;
                    }
                    if (distance[y * W + x] > dist_center) {
                        nearest[y * W + x] = k + start_label;
                        distance[y * W + x] = dist_center;
                        change = 1;
                    }
                }
            }
        }
        if (!change) break;

        /* recompute centres: float32 sequential sums in raster order */
        memset(n_elems, 0, sizeof(int64_t) * (size_t)n);
        memset(segments, 0, sizeof(REAL) * (size_t)(n * nf));
        for (int64_t y = 0; y < H; ++y) {
            for (int64_t x = 0; x < W; ++x) {
                if (mask && !mask[y * W + x]) continue;
                const int64_t k = nearest[y * W + x] - start_label;
                if (k < 0 || k >= n) continue; /* Cython would write out of bounds */
                REAL *seg = segments + k * nf;
                n_elems[k] += 1;
                seg[0] += (REAL)0;
                seg[1] += (REAL)y;
                seg[2] += (REAL)x;
                const REAL *px = image + (y * W + x) * C;
                for (int64_t c = 0; c < C; ++c) seg[3 + c] += px[c];
            }
        }
        for (int64_t k = 0; k < n; ++k) {
            const REAL cnt = (REAL)n_elems[k];
            for (int64_t c = 0; c < nf; ++c) segments[k * nf + c] /= cnt; /* 0/0 -> NaN */
        }

        if (slic_zero) {
            for (int64_t y = 0; y < H; ++y) {
                for (int64_t x = 0; x < W; ++x) {
                    if (mask && !mask[y * W + x]) continue;
                    const int64_t k = nearest[y * W + x] - start_label;
                    if (k < 0 || k >= n) continue;
                    const REAL *seg = segments + k * nf;
                    const REAL *px = image + (y * W + x) * C;
                    REAL dist_color = (REAL)0;
                    for (int64_t c = 0; c < C; ++c) {
                        const REAL t = px[c] - seg[3 + c];
                        dist_color = COLOR_ACC(dist_color, t);
                    }
                    if (max_dist_color[k] < dist_color) max_dist_color[k] = dist_color;
                }
            }
        }
    }
    free(distance);
    free(n_elems);
    free(max_dist_color);
    return it;
}

/*
 * One assignment sweep only (no centre update): the scatter half of a
 * _slic_cython iteration, used by the parity tests to compare the GPU
 * assignment kernel bit-for-bit against identical centres.
 */
void SUFFIX(obia_oracle_slic_assign_once)(const REAL *image, const uint8_t *mask,
                                  const REAL *segments, int64_t H, int64_t W,
                                  int64_t C, int64_t n, float step,
                                  int64_t start_label, int ignore_color,
                                  int64_t step_y, int64_t step_x,
                                  int64_t *nearest, REAL *distance)
{
    const int64_t nf = 3 + C;
    const float step_sq = step * step;
    const REAL spatial_weight = (REAL)(1.0 / (double)step_sq);
    for (int64_t i = 0; i < H * W; ++i) distance[i] = INFINITY;
    for (int64_t k = 0; k < n; ++k) {
        const REAL *seg = segments + k * nf;
        const REAL cy = seg[1], cx = seg[2];
        if (cy_trunc(seg[0]) == INT64_MIN) continue;
        const int64_t y_min = cy_trunc(cy_maxf(cy - (REAL)(2 * step_y), (REAL)0));
        const int64_t y_max = cy_trunc(cy_minf(cy + (REAL)(2 * step_y) + (REAL)1, (REAL)H));
        const int64_t x_min = cy_trunc(cy_maxf(cx - (REAL)(2 * step_x), (REAL)0));
        const int64_t x_max = cy_trunc(cy_minf(cx + (REAL)(2 * step_x) + (REAL)1, (REAL)W));
        for (int64_t y = y_min; y < y_max; ++y) {
            const REAL ty = (REAL)1 * (cy - (REAL)y);
            const REAL dy = ty * ty;
            for (int64_t x = x_min; x < x_max; ++x) {
                if (mask && !mask[y * W + x]) continue;
                const REAL tx = (REAL)1 * (cx - (REAL)x);
                const REAL dx = tx * tx;
                REAL dist_center = ((REAL)0 + dy + dx) * spatial_weight;
                if (!ignore_color) {
                    REAL dist_color = (REAL)0;
                    const REAL *px = image + (y * W + x) * C;
                    for (int64_t c = 0; c < C; ++c) {
                        const REAL t = px[c] - seg[3 + c];
                        dist_color = COLOR_ACC(dist_color, t);
                    }
                    dist_center += dist_color;
                }
                if (distance[y * W + x] > dist_center) {
                    nearest[y * W + x] = k + start_label;
                    distance[y * W + x] = dist_center;
                }
            }
        }
    }
}

/*
 * _enforce_label_connectivity_cython restated for depth == 1 (4-connected).
 * Sequential raster scan; BFS capped at max_size; components smaller than
 * min_size take the LAST-SEEN already-labelled different neighbour
 * (`adjacent`, initial value 0); output labels are consecutive in raster order
 * of each kept component's first pixel, starting at start_label
 * (SURVEY.md 3.4 step 9).  Neighbour order: x+1, x-1, y+1, y-1.
 * Returns 0, or -1 on a bad argument / allocation failure.
 */
#ifndef OBIA_ORACLE_NO_CC
int obia_oracle_enforce_connectivity(const int64_t *segments, int64_t H,
                                     int64_t W, int64_t min_size,
                                     int64_t max_size, int64_t start_label,
                                     int64_t *connected)
{
    if (max_size < 1) return -1; /* coord_list would have no room for the seed */
    static const int64_t ddx[4] = {1, -1, 0, 0};
    static const int64_t ddy[4] = {0, 0, 1, -1};
    const int64_t mask_label = start_label - 1;
    int64_t current_new_label = start_label;
    int64_t *coord = (int64_t *)malloc(sizeof(int64_t) * 2 * (size_t)max_size);
    if (!coord) return -1;
    for (int64_t i = 0; i < H * W; ++i) connected[i] = mask_label;

    for (int64_t y = 0; y < H; ++y) {
        for (int64_t x = 0; x < W; ++x) {
            if (segments[y * W + x] == mask_label) continue;
            if (connected[y * W + x] > mask_label) continue;
            int64_t adjacent = 0;
            const int64_t label = segments[y * W + x];
            connected[y * W + x] = current_new_label;
            int64_t current_segment_size = 1;
            int64_t bfs_visited = 0;
            coord[0] = y;
            coord[1] = x;
            while (bfs_visited < current_segment_size && current_segment_size < max_size) {
                for (int i = 0; i < 4; ++i) {
                    const int64_t yy = coord[2 * bfs_visited] + ddy[i];
                    const int64_t xx = coord[2 * bfs_visited + 1] + ddx[i];
                    if (0 <= xx && xx < W && 0 <= yy && yy < H) {
                        const int64_t q = yy * W + xx;
                        if (segments[q] == label && connected[q] == mask_label) {
                            connected[q] = current_new_label;
                            coord[2 * current_segment_size] = yy;
                            coord[2 * current_segment_size + 1] = xx;
                            current_segment_size += 1;
                            if (current_segment_size >= max_size) break;
                        } else if (connected[q] > mask_label && connected[q] != current_new_label) {
                            adjacent = connected[q];
                        }
                    }
                }
                bfs_visited += 1;
            }
            if (current_segment_size < min_size) {
                for (int64_t i = 0; i < current_segment_size; ++i)
                    connected[coord[2 * i] * W + coord[2 * i + 1]] = adjacent;
            } else {
                current_new_label += 1;
            }
        }
    }
    free(coord);
    return 0;
}
#endif

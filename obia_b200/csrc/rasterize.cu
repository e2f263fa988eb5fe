// Polygons -> label raster with the pixel-centre rule (a pixel belongs to a polygon when its centre
// lies inside, even-odd over all rings: holes and multi-part geometries).  Replaces the per-segment
// `rasterio.features.geometry_mask(..., invert=True)` of obia/utils/utils.py:53-67 (called for every
// row of `create_objects`, obia/segmentation/segment_statistics.py:479-484) and the
// `rasterio.features.rasterize` of obia/utils/tiling.py:248-255: with the label raster the statistics
// of ALL polygons come from one pass of the zonal kernel instead of a crop + mask per segment.
// One CTA per polygon walks its clipped bounding box; every thread tests its pixel centres against all
// edges (crossing number, float64, half-open edge rule so that polygons sharing an edge partition
// the pixels).  Where polygons overlap the one with the larger table row wins.
#include "common.cuh"

namespace obia {

__global__ void __launch_bounds__(128)
rasterize_polygons_kernel(const double *__restrict__ xy, const int32_t *__restrict__ ring_start,
                          const int32_t *__restrict__ poly_ring_start, const int32_t *__restrict__ poly_label,
                          const int32_t *__restrict__ bbox, int32_t *labels, int H, int W)
{
    const int p = blockIdx.x;
    const int x0 = bbox[4 * p], y0 = bbox[4 * p + 1], x1 = bbox[4 * p + 2], y1 = bbox[4 * p + 3];   // inclusive
    if (x1 < x0 || y1 < y0) return;
    const int bw = x1 - x0 + 1;
    const int64_t npx = (int64_t)bw * (y1 - y0 + 1);
    const int r0 = poly_ring_start[p], r1 = poly_ring_start[p + 1];
    const int32_t label = poly_label[p];
    for (int64_t i = threadIdx.x; i < npx; i += blockDim.x) {
        const int x = x0 + (int)(i % bw), y = y0 + (int)(i / bw);
        const double px = x + 0.5, py = y + 0.5;
        bool inside = false;
        for (int r = r0; r < r1; ++r) {
            const int v0 = ring_start[r], v1 = ring_start[r + 1];
            if (v1 - v0 < 3) continue;
            double xj = xy[2 * (int64_t)(v1 - 1)], yj = xy[2 * (int64_t)(v1 - 1) + 1];
            for (int v = v0; v < v1; ++v) {
                const double xi = xy[2 * (int64_t)v], yi = xy[2 * (int64_t)v + 1];
                if ((yi > py) != (yj > py)) {
                    const double xc = (xj - xi) * (py - yi) / (yj - yi) + xi;
                    if (px < xc) inside = !inside;
                }
                xj = xi;
                yj = yi;
            }
        }
        if (inside) atomicMax(labels + (int64_t)y * W + x, label);
    }
}

}  // namespace obia

using namespace obia;

extern "C" int obia_b200_rasterize_polygons(const double *xy, const int32_t *ring_start,
                                            const int32_t *poly_ring_start, const int32_t *poly_label,
                                            const int32_t *bbox, int64_t n_polygons, int32_t *labels, int64_t H,
                                            int64_t W, void *stream)
{
    if (!xy || !ring_start || !poly_ring_start || !poly_label || !bbox || !labels || n_polygons < 0 || H <= 0 || W <= 0)
        return set_err(OBIA_B200_ERR_ARG, "rasterize_polygons: bad argument");
    if (H * W >= 0x7fffffffLL || n_polygons >= 0x7fffffffLL)
        return set_err(OBIA_B200_ERR_UNSUPPORTED, "rasterize_polygons: H*W exceeds int32");
    if (n_polygons == 0) return OBIA_B200_OK;
    rasterize_polygons_kernel<<<(unsigned)n_polygons, 128, 0, (cudaStream_t)stream>>>(
        xy, ring_start, poly_ring_start, poly_label, bbox, labels, (int)H, (int)W);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

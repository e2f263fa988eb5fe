// Window descriptors of the batched tiled driver (obia_b200/utils/tiling.py): all windows of one batch --
// every black tile of a chunk, or every white window of one tile-row (reference
// obia/utils/tiling.py:103-153 and :156-287) -- are stacked into one SLAB raster (window i occupies slab
// rows [row0, row0 + h), columns [0, w); one masked gap row between windows) and every stage of the path
// runs ONCE over the slab with `blockIdx.z` / a per-pixel lookup selecting the window's parameters.
// The layout mirrors the numpy dtype WIN_DESC in obia_b200/batch.py field by field.
#pragma once
#include <stdint.h>

namespace obia {

struct WinDesc {
    int32_t y0, x0, h, w;                     //   0  window in the source raster (row, LOCAL column)
    int32_t row0, valid, n, c0;               //  16  slab row, usable flag, centres, first centre (batch-wide index)
    int32_t cell0, ncy, ncx, step_y;          //  32  SLIC cell grid of the window inside the batch-wide `head`
    int32_t step_x, min_size, max_size, n_mask;   //  48
    float sw, inv_w, fix_scale32, imin;       //  64  1/step^2, step^2 (score units), 32-bit fixed-point scale, global rescale
    float idiff;                              //  80
    int32_t rescale, p0, m;                   //  84  k-means points of the window: [p0, p0 + m)
    double fix_scale;                         //  96
    int64_t fix_ratio;                        // 104
    double km_cs;                             // 112  k-means search grid: cell size, shape, first cell
    int32_t km_ncy, km_ncx;                   // 120
    int32_t km_cell0, pad;                    // 128
};
static_assert(sizeof(WinDesc) == 136, "WinDesc layout is shared with obia_b200/batch.py");

}  // namespace obia

"""Optional host step: label raster -> polygons.

The reference polygonises with one full-raster `rasterio.features.shapes` call PER LABEL
(/root/reference/obia/segmentation/segment_boundaries.py:62-70, O(n_segments * H * W)).  Here the
whole label raster is traced ONCE and the polygons are grouped by value.  Needs rasterio + shapely,
which are not part of the GPU path (and not installed in the build image): without them this
raises ImportError and the segment tables keep `geometry = None`.
"""
from __future__ import annotations

import numpy as np


def polygons_from_labels(raster, row_values, affine_transformation=None):
    """One (multi)polygon per entry of `row_values` (the raster value of each table row)."""
    try:
        from rasterio.features import shapes
        from shapely.affinity import affine_transform
        from shapely.geometry import shape
        from shapely.ops import unary_union
    except ImportError as e:  # pragma: no cover - GDAL-family libraries are absent from this image
        raise ImportError("polygonize=True needs rasterio and shapely (host-side step)") from e
    raster = np.ascontiguousarray(raster, dtype=np.int32)
    parts = {}
    for geom, value in shapes(raster, mask=raster >= 0, connectivity=4):
        parts.setdefault(int(value), []).append(shape(geom))
    out = []
    for v in np.asarray(row_values).tolist():
        polys = parts.get(int(v), [])
        g = polys[0] if len(polys) == 1 else (unary_union(polys) if polys else None)
        if g is not None and affine_transformation is not None:
            g = affine_transform(g, affine_transformation)     # segment_boundaries.py:69
        out.append(g)
    return out

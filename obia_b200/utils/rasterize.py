"""Polygon geometries -> label raster on the GPU (pixel-centre rule).

Host glue of `obia_b200_rasterize_polygons`: replaces the per-segment
`rasterio.features.geometry_mask` of /root/reference/obia/utils/utils.py:53-67 (reached from
`create_objects`, segment_statistics.py:479-484): the reference crops the raster to the polygon's
bounding box and masks the pixels whose CENTRE lies outside the polygon (GDAL's default,
`all_touched=False`); here every polygon is burnt into one label raster, which the zonal kernel then
reduces in a single pass.

Geometries are anything with `__geo_interface__` (shapely Polygon / MultiPolygon, geopandas rows,
`utils.polygonize.SimplePolygon`) or GeoJSON-like dicts.  Coordinates are in the raster's CRS and
are mapped to pixel space with the inverse of `Image.affine_transformation`
([a, b, d, e, xoff, yoff], shapely order: x = a*col + b*row + xoff, y = d*col + e*row + yoff).
"""
from __future__ import annotations

import ctypes

import numpy as np


def _rings_of(geom):
    """List of rings (each an (n, 2) float64 array) of a Polygon / MultiPolygon-like geometry."""
    gi = geom if isinstance(geom, dict) else getattr(geom, "__geo_interface__", None)
    if gi is None:
        raise TypeError(f"geometry of type {type(geom).__name__} has no __geo_interface__")
    t = gi["type"]
    if t == "Polygon":
        polys = [gi["coordinates"]]
    elif t == "MultiPolygon":
        polys = gi["coordinates"]
    elif t == "GeometryCollection":
        out = []
        for g in gi["geometries"]:
            out += _rings_of(g)
        return out
    else:
        return []       # points / lines cover no pixel centre
    rings = []
    for poly in polys:
        for r in poly:
            a = np.asarray(r, dtype=np.float64)[:, :2]
            if len(a) >= 2 and np.array_equal(a[0], a[-1]):
                a = a[:-1]
            if len(a) >= 3:
                rings.append(a)
    return rings


def world_to_pixel(affine_transformation):
    """Inverse of the shapely-order affine [a, b, d, e, xoff, yoff] as a function on (n, 2) arrays."""
    if affine_transformation is None:
        return lambda xy: xy
    a, b, d, e, xoff, yoff = (float(v) for v in affine_transformation)
    det = a * e - b * d
    if det == 0:
        raise ValueError("singular affine transformation")

    def inv(xy):
        x, y = xy[:, 0] - xoff, xy[:, 1] - yoff
        return np.stack([(e * x - b * y) / det, (-d * x + a * y) / det], axis=1)
    return inv


def rasterize_polygons(geometries, H, W, affine_transformation=None, labels=None, device="cuda", out=None):
    """Label raster (H, W) int32 CUDA tensor: pixel -> `labels[i]` (default i + 1) of the polygon whose
    interior holds the pixel centre, -1 where there is none (`out`: burn into an existing raster)."""
    import torch

    from .. import _lib
    from ..pipeline import _p, _stream_ptr

    lib = _lib.load()
    inv = world_to_pixel(affine_transformation)
    n = len(geometries)
    vals = np.arange(1, n + 1, dtype=np.int32) if labels is None else np.asarray(labels, dtype=np.int32)
    verts, ring_start, poly_ring_start, bbox = [], [0], [0], np.empty((n, 4), dtype=np.int32)
    nv = 0
    for i, g in enumerate(geometries):
        rings = [] if g is None else _rings_of(g)
        lo = np.array([np.inf, np.inf])
        hi = -lo
        for r in rings:
            px = inv(r)
            verts.append(px)
            nv += len(px)
            ring_start.append(nv)
            lo, hi = np.minimum(lo, px.min(0)), np.maximum(hi, px.max(0))
        poly_ring_start.append(len(ring_start) - 1)
        if rings:
            # pixel centres (c + 0.5, r + 0.5) inside [lo, hi]
            x0, y0 = int(np.ceil(lo[0] - 0.5)), int(np.ceil(lo[1] - 0.5))
            x1, y1 = int(np.floor(hi[0] - 0.5)), int(np.floor(hi[1] - 0.5))
            bbox[i] = (max(x0, 0), max(y0, 0), min(x1, W - 1), min(y1, H - 1))
        else:
            bbox[i] = (0, 0, -1, -1)
    if out is None:
        out = torch.full((H, W), -1, dtype=torch.int32, device=device)
    if n == 0 or nv == 0:
        return out
    dev = out.device
    xy = torch.from_numpy(np.ascontiguousarray(np.concatenate(verts, axis=0))).to(dev)
    rs = torch.from_numpy(np.asarray(ring_start, dtype=np.int32)).to(dev)
    ps = torch.from_numpy(np.asarray(poly_ring_start, dtype=np.int32)).to(dev)
    lv = torch.from_numpy(vals).to(dev)
    bb = torch.from_numpy(bbox).to(dev)
    _lib.check(lib.obia_b200_rasterize_polygons(_p(xy), _p(rs), _p(ps), _p(lv), _p(bb), n, _p(out), H, W,
                                                _stream_ptr()), "rasterize_polygons")
    return out

"""CPU oracle for `create_tiled_segments` -- TEST INFRASTRUCTURE ONLY.

Restates /root/reference/obia/utils/tiling.py:62-291 on the CPU, following the reference's own
control flow (two raster-order loops, a growing list of segments, per-tile predicates), with a
segment represented by its pixel set instead of a shapely polygon:
    within(tile_polygon)   <=>  all pixels inside the polygon        (tiling.py:205-231)
    overlaps(tile_polygon) <=>  some but not all pixels inside
    rasterize(geometry)    <=>  the pixels themselves                (tiling.py:248-255)
The per-tile segmentation is the SLIC oracle (slic_oracle.create_segments_labels), i.e. what
`create_segments(image, mask=..., n_segments=..., method="slic", **kwargs)` produces up to the
label raster (tiling.py:137-143, :275-281).

Same documented deviations as the product (n_segments popped from kwargs, non-inverted mask when
no input mask is given, corner squares in pixels): see obia_b200/utils/tiling.py.

PARITY UNPINNED against the reference itself (GDAL / rasterio / geopandas are not installable
here); the product and this oracle are two independent restatements checked against each other.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may import this module.
"""
from __future__ import annotations

import math

import numpy as np

import slic_oracle as so


def _segment(raw, mask, n_segments, kw):
    """create_segments up to labels; list of boolean pixel masks (ascending label), or ValueError."""
    if n_segments <= 0:
        raise ValueError("n_segments must be positive")
    labels = so.create_segments_labels(np.array(raw, dtype=np.float32, copy=True), None,
                                       n_segments=n_segments, mask=mask, **kw)
    return [labels == i for i in np.unique(labels) if i != -1]


def create_tiled_segments(raw, input_mask=None, tile_size=200, buffer=30, crown_radius=5,
                          pixel_area=1.0, **kwargs):
    """Returns (labels (H, W) int32 with ids 1..N, -1 elsewhere; N)."""
    raw = np.asarray(raw, dtype=np.float32)
    height, width = raw.shape[:2]
    kw = dict(kwargs)
    fixed_n = kw.pop("n_segments", None)
    mask_full = None if input_mask is None else (np.asarray(input_mask) != 0)

    def n_for(mask):
        if fixed_n is not None:
            return int(fixed_n)
        if mask is None:
            raise ValueError("need input_mask or n_segments")
        return int(round(mask.sum() * pixel_area / (math.pi * crown_radius ** 2)))

    # every segment: dict(pix=(ys, xs) global coordinates, alive=bool)
    black, white = [], []

    def add(store, segs, y0, x0):
        for m in segs:
            ys, xs = np.nonzero(m)
            store.append(dict(ys=ys + y0, xs=xs + x0, alive=True))

    # ---- pass 1: black tiles (tiling.py:103-153)
    for j in range(0, height, tile_size):
        for i in range(0, width, tile_size):
            if (i // tile_size + j // tile_size) % 2 != 0:
                continue
            w, h = min(tile_size, width - i), min(tile_size, height - j)
            if w == 0 or h == 0:
                continue
            mask = None if mask_full is None else mask_full[j:j + h, i:i + w].copy()
            try:
                add(black, _segment(raw[j:j + h, i:i + w], mask, n_for(mask), kw), j, i)
            except ValueError:
                pass

    # ---- pass 2: white tiles (tiling.py:156-287)
    for j in range(0, height, tile_size):
        for i in range(0, width, tile_size):
            if (i // tile_size + j // tile_size) % 2 == 0:
                continue
            i0 = max(0, i - buffer)
            w = min(width, i + tile_size + buffer) - i0
            j0 = max(0, j - buffer)
            h = min(height, j + tile_size + buffer) - j0
            w, h = max(0, min(w, width - i0)), max(0, min(h, height - j0))
            if w == 0 or h == 0:
                continue
            mask = None if mask_full is None else mask_full[j0:j0 + h, i0:i0 + w].copy()
            # window polygon = box minus the two bottom corner squares (pixel-centre rule)
            poly = np.ones((h, w), bool)
            c = buffer / 2.0
            nc = max(0, min(int(math.ceil(c - 0.5)) if c > 0 else 0, h, w))
            if nc:
                poly[h - nc:, :nc] = False
                poly[h - nc:, w - nc:] = False
            corners = ~poly
            hits = []
            for s in black + white:
                if not s["alive"]:
                    continue
                ys, xs = s["ys"] - j0, s["xs"] - i0
                inwin = (ys >= 0) & (ys < h) & (xs >= 0) & (xs < w)
                inside = np.zeros(len(ys), bool)
                inside[inwin] = poly[ys[inwin], xs[inwin]]
                a = int(inside.sum())
                if a > 0:
                    hits.append((s, a == len(ys), inwin))
            if hits:
                rasterized = corners.copy()
                for s, is_within, inwin in hits:
                    if is_within:
                        s["alive"] = False                                   # :220-231
                    else:
                        rasterized[s["ys"][inwin] - j0, s["xs"][inwin] - i0] = True   # :233-255
                mask = (~rasterized) if mask is None else (mask & ~rasterized)
            try:
                add(white, _segment(raw[j0:j0 + h, i0:i0 + w], mask, n_for(mask), kw), j0, i0)
            except ValueError:
                pass

    # ---- concat + renumber (tiling.py:289-290)
    out = np.full((height, width), -1, np.int32)
    n = 0
    for s in black + white:
        if s["alive"]:
            n += 1
            out[s["ys"], s["xs"]] = n
    return out, n

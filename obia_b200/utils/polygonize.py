"""Host step after the GPU path: label raster -> polygons.

The reference polygonises with one full-raster `rasterio.features.shapes` call PER LABEL
(/root/reference/obia/segmentation/segment_boundaries.py:59-77, O(n_segments * H * W)) and
applies the raster's affine transform to every polygon (:69).  Here the whole label raster is
traced ONCE, vectorised in numpy, with the same geometric convention (pixel-edge polygons,
4-connectivity: regions that only touch at a pixel corner stay separate rings, exterior ring plus
one interior ring per hole, vertices only where the outline turns):

1. every grid edge between two different values yields one directed half-edge per side, oriented
   so that its own region lies on the right;
2. at a vertex an outline continues with the tightest right turn (right, straight, left), which
   is what keeps corner-touching pixels in separate rings;
3. the rings are the cycles of that successor map; they are ordered by pointer jumping (list
   ranking), so there is no per-edge Python loop;
4. a ring that passes twice through a vertex (two holes, or a hole and the outside, meeting at a
   pixel corner) is split there into simple rings, as OGC validity requires;
5. a ring with positive signed area (x right, y down) is an exterior, a negative one a hole of the
   same region (every table row is ONE 4-connected region, so holes need no point-in-polygon test).

Geometry objects are shapely Polygons when shapely is importable, otherwise `SimplePolygon`
(exterior / interiors / bounds / area / wkt / __geo_interface__).  Pure host code: it is optional
(`polygonize=True`) and timed separately from the GPU path; memory is O(outline length).
"""
from __future__ import annotations

import numpy as np


class SimplePolygon:
    """Minimal polygon container used when shapely is not installed."""

    def __init__(self, exterior, interiors=()):
        self.exterior = np.asarray(exterior, dtype=np.float64)
        self.interiors = [np.asarray(r, dtype=np.float64) for r in interiors]

    @staticmethod
    def _ring_area(r):
        x, y = r[:, 0], r[:, 1]
        return 0.5 * float(np.sum(x[:-1] * y[1:] - x[1:] * y[:-1]))

    @property
    def area(self):
        return abs(self._ring_area(self.exterior)) - sum(abs(self._ring_area(r)) for r in self.interiors)

    @property
    def bounds(self):
        e = self.exterior
        return float(e[:, 0].min()), float(e[:, 1].min()), float(e[:, 0].max()), float(e[:, 1].max())

    @property
    def wkt(self):
        def ring(r):
            return "(" + ", ".join(f"{x:.15g} {y:.15g}" for x, y in r) + ")"
        return "POLYGON (" + ", ".join(ring(r) for r in [self.exterior] + self.interiors) + ")"

    @property
    def __geo_interface__(self):
        return {"type": "Polygon",
                "coordinates": [[tuple(p) for p in r] for r in [self.exterior] + self.interiors]}

    def __repr__(self):
        return f"<SimplePolygon {len(self.exterior) - 1} vertices, {len(self.interiors)} holes>"


_DX = np.array([1, 0, -1, 0], dtype=np.int64)   # direction codes 0:+x 1:+y 2:-x 3:-y (y grows downwards)
_DY = np.array([0, 1, 0, -1], dtype=np.int64)


def _half_edges(raster):
    """(value, start vertex, direction) of every outline half-edge; the region is on the right."""
    H, W = raster.shape
    L = np.full((H + 2, W + 2), -1, dtype=np.int64)
    L[1:-1, 1:-1] = raster
    W1 = W + 1
    # horizontal grid edges: row boundary y in [0, H], between L[y, x+1] (above) and L[y+1, x+1] (below)
    up, down = L[:-1, 1:-1], L[1:, 1:-1]                      # (H+1, W)
    ys, xs = np.nonzero(up != down)
    u, d = up[ys, xs], down[ys, xs]
    v0 = ys * W1 + xs                                          # vertex (x, y)
    val = [d, u]
    start = [v0, v0 + 1]                                       # below: (x,y)->(x+1,y); above: (x+1,y)->(x,y)
    direc = [np.zeros_like(v0), np.full_like(v0, 2)]
    # vertical grid edges: column boundary x in [0, W], between L[y+1, x] (left) and L[y+1, x+1] (right)
    left, right = L[1:-1, :-1], L[1:-1, 1:]                    # (H, W+1)
    ys, xs = np.nonzero(left != right)
    l, r = left[ys, xs], right[ys, xs]
    v0 = ys * W1 + xs
    val += [l, r]
    start += [v0, v0 + W1]                                     # left: (x,y)->(x,y+1); right: (x,y+1)->(x,y)
    direc += [np.ones_like(v0), np.full_like(v0, 3)]
    val, start, direc = np.concatenate(val), np.concatenate(start), np.concatenate(direc)
    keep = val >= 0
    return val[keep], start[keep], direc[keep]


def trace_rings(raster):
    """All outline rings of a label raster (values < 0 are background).

    Returns (ring_value, ring_area, ring_ptr, xs, ys): ring k belongs to value ring_value[k], has
    signed area ring_area[k] (> 0 exterior, < 0 hole) and the CLOSED vertex sequence
    xs/ys[ring_ptr[k]:ring_ptr[k+1]] (first vertex repeated at the end), corners only.  A ring may
    pass twice through a vertex where it pinches (see `split_ring`).
    """
    raster = np.asarray(raster)
    H, W = raster.shape
    W1 = W + 1
    val, start, direc = _half_edges(raster)
    E = val.size
    if E == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, np.zeros(0), np.zeros(1, dtype=np.int64), z, z
    end = start + _DX[direc] + _DY[direc] * W1
    nvert = (H + 1) * W1
    # successor: same value, starts where this edge ends, tightest right turn first
    key = (val * nvert + start) * 4 + direc
    order = np.argsort(key, kind="stable")
    skey = key[order]
    nxt = np.full(E, -1, dtype=np.int64)
    base = (val * nvert + end) * 4
    for turn in (1, 0, 3):
        want = base + (direc + turn) % 4
        pos = np.searchsorted(skey, want)
        pos[pos >= E] = E - 1
        hit = (skey[pos] == want) & (nxt < 0)
        nxt[hit] = order[pos[hit]]
    assert (nxt >= 0).all(), "open outline: label raster edges do not form closed rings"
    # ring id = smallest edge index on the cycle (min-propagation with pointer doubling)
    rid = np.arange(E, dtype=np.int64)
    jump = nxt.copy()
    while True:
        new = np.minimum(rid, rid[jump])
        jump = jump[jump]
        if np.array_equal(new, rid):
            break
        rid = new
    # position along the ring: cut every cycle before its representative and rank the lists
    is_rep = rid == np.arange(E)
    dist = np.where(is_rep[nxt], 0, 1).astype(np.int64)       # steps until the representative is reached again
    jump = np.where(is_rep[nxt], np.arange(E), nxt)            # the edge entering the representative is terminal
    while True:
        nd = dist + np.where(jump == np.arange(E), 0, dist[jump])
        nj = jump[jump]
        if np.array_equal(nj, jump):
            dist = nd
            break
        dist, jump = nd, nj
    ring_len = np.bincount(rid, minlength=E)[rid]
    pos_in_ring = (ring_len - 1 - dist) % ring_len             # representative -> 0, then along `nxt`
    seq = np.lexsort((pos_in_ring, rid))
    rid_s, dir_s, start_s, val_s = rid[seq], direc[seq], start[seq], val[seq]
    # keep only corners: an edge's start vertex is a corner iff the previous edge had another direction
    first = np.r_[True, rid_s[1:] != rid_s[:-1]]
    last = np.r_[first[1:], True]
    prev_dir = np.empty_like(dir_s)
    prev_dir[1:] = dir_s[:-1]
    prev_dir[first] = dir_s[last]                              # cyclic predecessor of a ring's first edge
    corner = dir_s != prev_dir
    cx, cy, crid, cval = start_s[corner] % W1, start_s[corner] // W1, rid_s[corner], val_s[corner]
    cfirst = np.r_[True, crid[1:] != crid[:-1]]
    starts = np.flatnonzero(cfirst)
    counts = np.diff(np.r_[starts, crid.size])
    # close the rings (repeat the first vertex) and compute signed areas
    nring = starts.size
    out_ptr = np.r_[0, np.cumsum(counts + 1)]
    xs = np.empty(out_ptr[-1], dtype=np.int64)
    ys = np.empty(out_ptr[-1], dtype=np.int64)
    dest = np.arange(crid.size) + np.repeat(np.arange(nring), counts)
    xs[dest], ys[dest] = cx, cy
    xs[out_ptr[1:] - 1], ys[out_ptr[1:] - 1] = cx[starts], cy[starts]
    cross = xs[:-1] * ys[1:] - xs[1:] * ys[:-1]
    cross[out_ptr[1:-1] - 1] = 0                               # no term across two rings
    area = 0.5 * np.add.reduceat(cross, out_ptr[:-1])
    return cval[starts], area, out_ptr, xs, ys


def split_ring(xs, ys):
    """Split a closed ring that touches itself at vertices into simple closed rings."""
    path, seen, out = [], {}, []
    for p in zip(xs[:-1].tolist(), ys[:-1].tolist()):
        if p in seen:                      # the loop since the first visit is a ring of its own
            i = seen[p]
            loop = path[i:]
            for q in loop[1:]:
                del seen[q]
            del path[i + 1:]
            out.append(loop + [p])
        else:
            seen[p] = len(path)
            path.append(p)
    out.append(path + [path[0]])
    return [np.asarray(r, dtype=np.int64) for r in out]


def _signed_area(r):
    return 0.5 * float(np.sum(r[:-1, 0] * r[1:, 1] - r[1:, 0] * r[:-1, 1]))


def polygons_from_labels(raster, row_values, affine_transformation=None):
    """One polygon per entry of `row_values` (the raster value of each table row; every value is one
    4-connected region).  `affine_transformation` = [a, b, d, e, xoff, yoff] as in
    `shapely.affinity.affine_transform` (obia's `Image.affine_transformation`, segment_boundaries.py:69)."""
    try:
        from shapely.geometry import Polygon as _ShapelyPolygon
    except ImportError:
        _ShapelyPolygon = None
    ring_val, area, ptr, xs, ys = trace_rings(raster)
    W1 = np.asarray(raster).shape[1] + 1
    # rings that touch themselves at a vertex (rare: pinch points) are split into simple rings
    vid = ys * W1 + xs
    closing = np.zeros(vid.size, dtype=bool)
    closing[ptr[1:] - 1] = True
    ring_of = np.repeat(np.arange(ring_val.size), np.diff(ptr))
    open_key = np.sort(ring_of[~closing] * (vid.max() + 1 if vid.size else 1) + vid[~closing])
    dup_rings = set((open_key[1:][open_key[1:] == open_key[:-1]] // (vid.max() + 1 if vid.size else 1)).tolist())

    def transform(r):
        r = r.astype(np.float64)
        if affine_transformation is None:
            return r
        a, b, d, e, xoff, yoff = (float(v) for v in affine_transformation)
        return np.stack([a * r[:, 0] + b * r[:, 1] + xoff, d * r[:, 0] + e * r[:, 1] + yoff], axis=1)

    by_value = {}
    for k in np.argsort(ring_val, kind="stable").tolist():
        ring = np.stack([xs[ptr[k]:ptr[k + 1]], ys[ptr[k]:ptr[k + 1]]], axis=1)
        parts = [(ring, float(area[k]))] if k not in dup_rings else \
            [(r, _signed_area(r)) for r in split_ring(ring[:, 0], ring[:, 1])]
        by_value.setdefault(int(ring_val[k]), []).extend(parts)
    out = []
    for v in np.asarray(row_values).tolist():
        rings = by_value.get(int(v))
        if not rings:
            out.append(None)
            continue
        ext = [r for r, a in rings if a > 0]
        holes = [transform(r) for r, a in rings if a < 0]
        if len(ext) != 1:
            raise ValueError(f"value {v} is not one 4-connected region ({len(ext)} exterior rings)")
        shell = transform(ext[0])
        out.append(_ShapelyPolygon(shell, holes) if _ShapelyPolygon is not None else SimplePolygon(shell, holes))
    return out

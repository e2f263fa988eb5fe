"""CPU tests of the vectorised label-raster outline tracer (obia_b200/utils/polygonize.py):
the host step that replaces the per-label rasterio.features.shapes loop of
obia/segmentation/segment_boundaries.py:59-77."""
import numpy as np
import pytest
from scipy import ndimage

from obia_b200.utils.polygonize import SimplePolygon, polygons_from_labels, trace_rings


def _rings(p):
    if isinstance(p, SimplePolygon):
        return p.exterior, p.interiors
    return np.asarray(p.exterior.coords), [np.asarray(r.coords) for r in p.interiors]


def _rasterise(p, H, W):
    """Even-odd fill at pixel centres from the vertical edges of all rings (rectilinear polygons)."""
    ext, holes = _rings(p)
    out = np.zeros((H, W), dtype=np.int64)
    yc = np.arange(H) + 0.5
    for r in [ext] + list(holes):
        for (x0, y0), (x1, y1) in zip(r[:-1], r[1:]):
            if x0 == x1 and y0 != y1:
                rows = (yc > min(y0, y1)) & (yc < max(y0, y1))
                out[rows, int(x0):] += 1            # crossing to the left of every pixel centre with x > x0
    return (out % 2).astype(bool)


def test_single_pixel_and_rectangle_exact_rings():
    r = np.full((3, 4), -1, np.int32)
    r[1, 2] = 5
    val, area, ptr, xs, ys = trace_rings(r)
    assert val.tolist() == [5] and area.tolist() == [1.0]
    assert sorted(zip(xs[:-1].tolist(), ys[:-1].tolist())) == [(2, 1), (2, 2), (3, 1), (3, 2)]
    assert (xs[0], ys[0]) == (xs[-1], ys[-1])
    r[:] = 7                                          # one rectangle: corners only, no collinear vertices
    (p,) = polygons_from_labels(r, [7])
    ext, holes = _rings(p)
    assert len(ext) == 5 and not holes and p.area == 12.0 and tuple(p.bounds) == (0.0, 0.0, 4.0, 3.0)


def test_hole_and_corner_touching_pixels():
    r = np.zeros((5, 5), np.int32)
    r[1:4, 1:4] = 1
    r[2, 2] = 2                                       # label 1 is a ring with a one-pixel hole
    p0, p1, p2 = polygons_from_labels(r, [0, 1, 2])
    assert p1.area == 8.0 and len(_rings(p1)[1]) == 1 and p2.area == 1.0
    assert p0.area == 16.0 and len(_rings(p0)[1]) == 1
    # 4-connectivity: two pixels of one value touching at a corner are two exterior rings
    d = np.full((2, 2), -1, np.int32)
    d[0, 0] = d[1, 1] = 3
    val, area, *_ = trace_rings(d)
    assert val.tolist() == [3, 3] and area.tolist() == [1.0, 1.0]
    with pytest.raises(ValueError):
        polygons_from_labels(d, [3])
    # a hole that touches the outline's inner corner diagonally stays a hole of the same ring
    e = np.ones((4, 4), np.int32)
    e[1, 1] = 0
    e[2, 2] = 0
    (pe,) = polygons_from_labels(e, [1])
    assert pe.area == 14.0 and len(_rings(pe)[1]) == 2


@pytest.mark.parametrize("seed,shape", [(0, (40, 50)), (1, (64, 64)), (2, (17, 90)), (3, (1, 1)), (4, (1, 23)),
                                        (5, (31, 2)), (6, (5, 5)), (7, (128, 96))])
def test_random_regions_fill_back_exactly(seed, shape):
    """Every 4-connected region of a random label raster: polygon area == pixel count, the
    polygon rasterises back to exactly the region, rings are closed and axis-aligned."""
    rs = np.random.RandomState(seed)
    H, W = shape
    coarse = rs.randint(0, 4, size=(H // 4 + 1, W // 4 + 1))
    lab = np.kron(coarse, np.ones((4, 4), int))[:H, :W]
    lab[rs.rand(H, W) < 0.08] = 9                     # speckle: holes, pinches, single pixels
    lab[:3, :5] = -1                                   # background / masked
    region = np.full((H, W), -1, np.int64)
    n = 0
    for v in np.unique(lab[lab >= 0]):
        cc, k = ndimage.label(lab == v)               # 4-connectivity
        region[cc > 0] = cc[cc > 0] + n - 1
        n += k
    polys = polygons_from_labels(region, np.arange(n))
    assert len(polys) == n
    counts = np.bincount(region[region >= 0], minlength=n)
    for i, p in enumerate(polys):
        assert p.area == counts[i]
        ext, holes = _rings(p)
        for r in [ext] + list(holes):
            assert (r[0] == r[-1]).all()
            step = np.abs(np.diff(r, axis=0))
            assert ((step[:, 0] == 0) ^ (step[:, 1] == 0)).all()          # axis-aligned, no zero-length edge
            turn = np.diff(np.r_[step[:, 0] == 0, step[0, 0] == 0].astype(int))
            assert (turn != 0).all()                                      # corners only
    for i in rs.choice(n, size=min(n, 60), replace=False):
        np.testing.assert_array_equal(_rasterise(polys[i], H, W), region == i)


def test_affine_and_missing_values():
    r = np.zeros((2, 3), np.int32)
    (p, q) = polygons_from_labels(r, [0, 4], affine_transformation=[10.0, 0.0, 0.0, -10.0, 500.0, 900.0])
    assert q is None
    assert tuple(p.bounds) == (500.0, 880.0, 530.0, 900.0) and p.area == 600.0
    assert p.wkt.startswith("POLYGON ((") and p.__geo_interface__["type"] == "Polygon"


def test_segments_frame_geojson_roundtrip_on_cpu(tmp_path):
    """SegmentsFrame.materialize_geometry + to_file (GeoJSON, no geopandas) from a CPU label raster."""
    import json
    import pandas as pd
    import torch
    from obia_b200.segmentation.segment_boundaries import SegmentsFrame
    lab = np.zeros((6, 8), np.int32)
    lab[:, 4:] = 1
    lab[2:4, 1:3] = 2
    f = SegmentsFrame({"geometry": None, "segment_id": [1, 2, 3]}, index=pd.RangeIndex(3))
    f.label_raster = torch.from_numpy(lab)
    f.segment_labels = np.array([0, 1, 2])
    f.crs = "EPSG:32702"
    with pytest.raises(ValueError):
        f.to_file(str(tmp_path / "x.geojson"))                 # no geometries yet
    f.materialize_geometry([1.0, 0.0, 0.0, -1.0, 10.0, 20.0])
    f["b0_mean"] = [1.5, float("nan"), 3.0]
    f.to_file(str(tmp_path / "x.geojson"))
    doc = json.loads((tmp_path / "x.geojson").read_text())
    assert doc["crs"]["properties"]["name"] == "EPSG:32702" and len(doc["features"]) == 3
    assert doc["features"][1]["properties"] == {"segment_id": 2, "b0_mean": None}
    assert len(doc["features"][0]["geometry"]["coordinates"]) == 2            # label 0 has a hole (label 2)
    assert [g.area for g in f["geometry"]] == [20.0, 24.0, 4.0]
    with pytest.raises(NotImplementedError):
        f.to_file(str(tmp_path / "x.gpkg"))


def test_image_stats_dtype_rule():
    """Image.stats_in_float64: integer rasters are reduced in float64 by the reference (utils.py:64)."""
    import torch
    from obia_b200.handlers.geotif import Image
    assert not Image(np.zeros((2, 2, 1), np.float32), None, None, None, None).stats_in_float64()
    assert Image(np.zeros((2, 2, 1), np.uint8), None, None, None, None).stats_in_float64()
    assert Image(np.zeros((2, 2, 1), np.float64), None, None, None, None).stats_in_float64()
    assert not Image(torch.zeros((2, 2, 1)), None, None, None, None).stats_in_float64()
    assert Image(torch.zeros((2, 2, 1), dtype=torch.int16), None, None, None, None).stats_in_float64()

    class _File:           # rasterio dataset stand-in: the FILE dtype decides, not img_data (always float32)
        dtypes = ("uint8", "uint8")
    assert Image(np.zeros((2, 2, 2), np.float32), None, None, None, _File()).stats_in_float64()

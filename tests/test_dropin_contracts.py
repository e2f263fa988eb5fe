"""GPU: drop-in contracts around the hot path -- create_objects on arbitrary polygon tables (the
CUDA scanline rasteriser vs a numpy point-in-polygon reference), filtered segment tables,
label_segments + a classify-shaped consumer of the feature columns, repeated calls on one Image."""
import numpy as np
import pandas as pd
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _pnpoly_mask(rings, H, W):
    """numpy reference: pixel centres inside (even-odd over all rings), half-open edge rule."""
    yy, xx = np.mgrid[:H, :W]
    px, py = xx + 0.5, yy + 0.5
    inside = np.zeros((H, W), bool)
    for r in rings:
        xj, yj = r[-1]
        for xi, yi in r:
            cond = (yi > py) != (yj > py)
            with np.errstate(divide="ignore", invalid="ignore"):
                xc = (xj - xi) * (py - yi) / (yj - yi) + xi
            inside ^= cond & (px < xc)
            xj, yj = xi, yi
    return inside


def test_rasterize_polygons_pixel_centre_rule():
    from obia_b200.utils.polygonize import SimplePolygon
    from obia_b200.utils.rasterize import rasterize_polygons
    H, W = 60, 80
    rng = np.random.RandomState(5)
    polys, ref = [], np.full((H, W), -1, np.int32)
    for i in range(12):
        cx, cy = rng.uniform(5, W - 5), rng.uniform(5, H - 5)
        k = rng.randint(3, 9)
        ang = np.sort(rng.uniform(0, 2 * np.pi, k))
        rad = rng.uniform(3, 14, k)
        ext = np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], 1)
        holes = []
        if i % 3 == 0:
            holes = [np.stack([cx + 1.3 * np.cos(ang[::-1]), cy + 1.3 * np.sin(ang[::-1])], 1)]
        polys.append(SimplePolygon(np.vstack([ext, ext[:1]]), [np.vstack([h, h[:1]]) for h in holes]))
        m = _pnpoly_mask([ext] + holes, H, W)
        ref[m] = np.maximum(ref[m], i + 1)           # larger row wins overlaps
    got = rasterize_polygons(polys, H, W).cpu().numpy()
    np.testing.assert_array_equal(got, ref)
    # affine: x = 10 + 2*col, y = 500 - 2*row (north-up raster with 2-unit pixels)
    aff = [2.0, 0.0, 0.0, -2.0, 10.0, 500.0]
    world = []
    for p in polys:
        f = lambda r: np.stack([10 + 2 * r[:, 0], 500 - 2 * r[:, 1]], 1)
        world.append(SimplePolygon(f(p.exterior), [f(h) for h in p.interiors]))
    got2 = rasterize_polygons(world, H, W, aff).cpu().numpy()
    np.testing.assert_array_equal(got2, ref)


def test_create_objects_on_arbitrary_polygon_table():
    """`create_objects(any table with geometry + segment_id, image)` like the reference
    (segment_statistics.py:392-511): the polygons of a previous segmentation, stripped of everything
    obia_b200-specific, give the same feature table as the label-raster path."""
    from obia_b200.handlers.geotif import Image
    from obia_b200.segmentation.segment_boundaries import create_segments
    from obia_b200.segmentation.segment_statistics import create_objects
    from gpu_helpers import synth_raster
    raw = synth_raster(96, 112, 4, seed=9)
    aff = [0.5, 0.0, 0.0, -0.5, 1000.0, 2000.0]
    img = Image(raw.copy(), "EPSG:32702", aff, None, None)
    segs = create_segments(img, None, "slic", n_segments=60, compactness=0.3, polygonize=True, mutate_image=False)
    want = create_objects(segs, img, calculate_textural=False)
    plain = pd.DataFrame({"geometry": list(segs["geometry"]), "segment_id": list(segs["segment_id"])})
    got = create_objects(plain, img, calculate_textural=False)
    assert list(got.columns) == list(want.columns)
    num = [c for c in want.columns if c not in ("geometry",)]
    np.testing.assert_allclose(got[num].to_numpy(dtype=float), want[num].to_numpy(dtype=float), rtol=1e-12,
                               equal_nan=True)
    # a reordered subset of the rows (user-edited table): each row keeps its own polygon's statistics
    sub = plain.iloc[[7, 2, 11]].reset_index(drop=True)
    got_sub = create_objects(sub, img, calculate_textural=False)
    np.testing.assert_allclose(got_sub[num].to_numpy(dtype=float), want[num].iloc[[7, 2, 11]].to_numpy(dtype=float),
                               rtol=1e-12, equal_nan=True)


def test_create_objects_on_filtered_segments_frame():
    """`create_objects(segments[filter], image)` -- rows follow their segment_id."""
    from obia_b200.handlers.geotif import Image
    from obia_b200.segmentation.segment_boundaries import create_segments
    from obia_b200.segmentation.segment_statistics import create_objects
    from gpu_helpers import synth_raster
    raw = synth_raster(80, 90, 3, seed=4, quantize=True)
    img = Image(raw.copy(), "EPSG:32702", [1, 0, 0, -1, 0, 0], None, None)
    segs = create_segments(img, None, "slic", n_segments=40, compactness=10, mutate_image=False)
    full = create_objects(segs, img, calculate_textural=False)
    keep = segs["segment_id"] % 3 == 1
    part = create_objects(segs[keep], img, calculate_textural=False)
    assert list(part["segment_id"]) == list(segs["segment_id"][keep])
    num = [c for c in full.columns if c != "geometry"]
    np.testing.assert_array_equal(part[num].to_numpy(dtype=float), full[num][keep.to_numpy()].to_numpy(dtype=float))


def test_label_segments_and_classify_shaped_consumer():
    """The interop chain of the reference: feature table -> label_segments (utils.py:12-34) ->
    classify's feature matrix `x = table.drop(['feature_class', 'geometry', 'segment_id'])`
    (classification/classify.py:83, :125) -> sklearn RandomForest."""
    from sklearn.ensemble import RandomForestClassifier
    from obia_b200.handlers.geotif import Image
    from obia_b200.segmentation.segment import segment
    from obia_b200.utils.utils import label_segments
    from gpu_helpers import synth_raster
    raw = synth_raster(120, 120, 3, seed=2, quantize=True)
    aff = [1.0, 0.0, 0.0, -1.0, 300.0, 900.0]
    img = Image(raw.copy(), "EPSG:32702", aff, None, None)
    seg = segment(img, [0, 1, 2], None, "slic", n_segments=300, compactness=10, calc_contrast=False,
                  calc_dissimilarity=False, calc_homogeneity=False, calc_ASM=False, calc_energy=False,
                  calc_correlation=False)
    table = seg.segments
    lab = table.label_raster.cpu().numpy()
    rows = np.asarray(table.segment_labels)
    # one labelled point at an inner pixel of 30 segments; two points of different classes in one more
    pts, cls = [], []
    rng = np.random.RandomState(0)
    for i in range(31):
        ys, xs = np.nonzero(lab == rows[i])
        k = rng.randint(len(ys))
        pts.append((300.0 + xs[k] + 0.5, 900.0 - (ys[k] + 0.5)))
        cls.append("tree" if table["b0_mean"].iloc[i] > table["b0_mean"].median() else "ground")
    ys, xs = np.nonzero(lab == rows[30])
    pts.append((300.0 + xs[0] + 0.5, 900.0 - (ys[0] + 0.5)))
    cls.append("ground" if cls[30] == "tree" else "tree")
    points = pd.DataFrame({"geometry": [{"type": "Point", "coordinates": p} for p in pts], "class": cls})
    labelled, mixed = label_segments(table, points)
    assert mixed == [int(table["segment_id"].iloc[30])]
    assert len(labelled) == 30 and list(labelled["feature_class"]) == cls[:30]
    # classify's feature matrix: every remaining column is a feature (NaN point-cloud columns dropped
    # by the user in the reference's notebooks; sklearn needs finite input)
    x = labelled.drop(["feature_class", "geometry", "segment_id"], axis=1).dropna(axis=1, how="all")
    assert list(x.columns)[:6] == ["b0_mean", "b0_variance", "b0_min", "b0_max", "b0_skewness", "b0_kurtosis"]
    clf = RandomForestClassifier(n_estimators=20, random_state=0).fit(x.to_numpy(), labelled["feature_class"])
    x_all = table.drop(["feature_class", "geometry", "segment_id"], axis=1, errors="ignore").dropna(axis=1, how="all")
    pred = clf.predict(x_all.to_numpy())
    assert len(pred) == len(table) and set(pred) <= {"tree", "ground"}
    assert (pred[:30] == np.asarray(cls[:30])).mean() > 0.9


def test_create_segments_twice_on_a_cuda_image():
    """img_data resident on the GPU is normalised from the cached RAW values on every call
    (the reference's re-normalisation of normalised data is the identity), never rescaled twice."""
    from obia_b200.handlers.geotif import Image
    from obia_b200.segmentation.segment_boundaries import create_segments
    from gpu_helpers import synth_raster
    raw = synth_raster(64, 72, 3, seed=1, quantize=True)
    img = Image(_cuda(raw), "EPSG:32702", [1, 0, 0, -1, 0, 0], None, None)
    a = create_segments(img, None, "slic", n_segments=30, compactness=10)
    first = img.img_data.clone()
    b = create_segments(img, None, "slic", n_segments=30, compactness=10)
    want = raw.copy()
    for c in range(3):
        want[:, :, c] = (raw[:, :, c] - raw[:, :, c].min()) / (raw[:, :, c].max() - raw[:, :, c].min())
    np.testing.assert_array_equal(first.cpu().numpy(), want)
    np.testing.assert_array_equal(img.img_data.cpu().numpy(), want)
    assert torch.equal(a.label_raster, b.label_raster)


def test_raw_cache_follows_img_data():
    """assigning a new raster to `image.img_data` invalidates the cached raw values."""
    from obia_b200.handlers.geotif import Image
    from obia_b200.segmentation.segment_boundaries import create_segments
    from gpu_helpers import synth_raster
    r1, r2 = synth_raster(48, 56, 3, seed=1), synth_raster(48, 56, 3, seed=2)
    img = Image(r1.copy(), "EPSG:32702", [1, 0, 0, -1, 0, 0], None, None)
    a = create_segments(img, None, "slic", n_segments=20, compactness=1.0, mutate_image=False)
    img.img_data = r2.copy()
    b = create_segments(img, None, "slic", n_segments=20, compactness=1.0, mutate_image=False)
    fresh = create_segments(Image(r2.copy(), "EPSG:32702", [1, 0, 0, -1, 0, 0], None, None), None, "slic",
                            n_segments=20, compactness=1.0, mutate_image=False)
    assert torch.equal(b.label_raster, fresh.label_raster) and not torch.equal(a.label_raster, b.label_raster)
    # the in-place normalisation done by create_segments itself does not invalidate the raw values
    img2 = Image(r1.copy(), "EPSG:32702", [1, 0, 0, -1, 0, 0], None, None)
    create_segments(img2, None, "slic", n_segments=20, compactness=1.0)
    raw_cached = img2.device_raw()
    np.testing.assert_array_equal(raw_cached.cpu().numpy(), r1)


def test_candidate_overflow_is_a_value_error():
    from obia_b200 import _lib
    assert issubclass(_lib.CandidateOverflowError, ValueError) and issubclass(_lib.CandidateOverflowError,
                                                                              _lib.ObiaB200Error)


@pytest.mark.parametrize("C,bands,kw", [
    (3, None, dict(kernel_size=3, max_dist=6)),                                       # Lab path (skimage default)
    (4, [0, 2], dict(kernel_size=2, max_dist=5, convert2lab=False, ratio=0.5)),
    (5, None, dict(kernel_size=3, max_dist=8, convert2lab=False, sigma=1.0, rng=7)),
    (3, None, dict(kernel_size=1, max_dist=2, ratio=0.2, random_seed=3)),
])
def test_quickshift_against_the_oracle(C, bands, kw):
    """method="quickshift" (segment_boundaries.py:48-49): densities, parents and the raster-order root
    numbering of scikit-image's `_quickshift_cython` as restated in oracle/quickshift_oracle.py."""
    import quickshift_oracle as qo
    from obia_b200 import pipeline
    from gpu_helpers import synth_raster
    H, W = 57, 83
    raw = synth_raster(H, W, C, seed=C, quantize=(C == 3))
    okw = {("rng" if k == "random_seed" else k): v for k, v in kw.items()}
    want = qo.create_segments_labels(raw, bands, **okw)
    got, n = pipeline.quickshift_labels(_cuda(raw), bands, **kw)
    got = got.cpu().numpy()
    agree = float((got == want).mean())
    print(f"quickshift agreement {agree:.5f}, segments gpu={n} oracle={want.max() + 1}")
    assert agree >= 0.995 and abs(n - (want.max() + 1)) <= max(1, 0.01 * n)


def test_create_segments_quickshift_table():
    from obia_b200.handlers.geotif import Image
    from obia_b200.segmentation.segment_boundaries import create_segments
    from gpu_helpers import synth_raster
    raw = synth_raster(60, 70, 3, seed=5, quantize=True)
    img = Image(raw.copy(), "EPSG:32702", [1, 0, 0, -1, 0, 0], None, None)
    segs = create_segments(img, None, "quickshift", kernel_size=3, max_dist=6, ratio=0.5)
    assert list(segs["segment_id"]) == list(range(1, len(segs) + 1)) and len(segs) >= 2
    assert (np.diff(np.asarray(segs.segment_labels)) >= 0).all()
    want = raw.copy()
    for c in range(3):
        want[:, :, c] = (raw[:, :, c] - raw[:, :, c].min()) / (raw[:, :, c].max() - raw[:, :, c].min())
    np.testing.assert_array_equal(img.img_data, want)          # the in-place normalisation side effect
    with pytest.raises(ValueError, match="Lab"):
        create_segments(Image(raw[:, :, :2].copy(), "EPSG:32702", [1, 0, 0, -1, 0, 0], None, None), None, "quickshift")
    with pytest.raises(Exception, match="unknown segmentation method"):
        create_segments(img, None, "watershed")

"""CPU: host-side logic of the product (grid geometry, centre initialisation, Gaussian taps,
fixed-point scale, feature-column contract) against the oracle / the real scipy."""
import numpy as np
import pytest

import slic_oracle as so
from obia_b200 import slic_host


@pytest.mark.parametrize("H,W,n", [(2048, 2048, 3000), (10000, 10000, 200000), (200, 200, 100), (37, 911, 50),
                                   (64, 64, 5000), (8, 8, 100), (300, 20, 7), (1, 50, 5)])
def test_grid_matches_oracle(H, W, n):
    yx, steps = slic_host.grid_centroids(H, W, n)
    c, s = so._get_grid_centroids((1, H, W), n)
    np.testing.assert_array_equal(yx, c[:, 1:])
    np.testing.assert_array_equal(steps, s)
    assert slic_host.window_steps(H, W, len(yx)) == so.grid_steps((1, H, W), len(c))[1:]


def test_mask_centroids_match_oracle():
    yy, xx = np.mgrid[:90, :120]
    mask = ((yy - 40) ** 2 + (xx - 70) ** 2) < 35 ** 2
    yx, steps = slic_host.mask_centroids(mask, 25)
    c, s = so._get_mask_centroids(mask[np.newaxis].astype(np.uint8), 25, True)
    np.testing.assert_allclose(yx, c[:, 1:], rtol=0, atol=0)
    np.testing.assert_allclose(steps, s)
    with pytest.raises(ValueError):
        slic_host.mask_centroids(np.zeros((5, 5), bool), 3)


@pytest.mark.parametrize("sigma", [0.5, 1.0, 1.3, 2.7])
def test_gaussian_taps_reproduce_scipy(sigma):
    from scipy import ndimage
    w, r = slic_host.gaussian_taps(np.float32(sigma))
    assert r == int(4.0 * float(np.float32(sigma)) + 0.5) and len(w) == 2 * r + 1
    x = np.random.RandomState(0).rand(64).astype(np.float32)
    want = ndimage.gaussian_filter1d(x, float(np.float32(sigma)), mode="reflect", truncate=4.0)
    pad = np.concatenate([x[::-1], x, x[::-1]]).astype(np.float64)      # half-sample symmetric
    got = np.array([np.dot(w, pad[64 + i - r:64 + i + r + 1]) for i in range(64)]).astype(np.float32)
    np.testing.assert_allclose(got, want, rtol=2e-6)


def test_fixed_point_scale_leaves_headroom():
    for (H, W, sy, sx, mx) in [(10000, 10000, 22, 22, 40.0), (200, 200, 20, 20, 0.1), (64, 64, 1, 1, 1e4),
                               (40000, 40000, 2000, 2000, 2560.0)]:
        s = slic_host.fixed_point_scale(mx, H, W, sy, sx)
        reach = min(H * W, (4 * sy + 1) * (4 * sx + 1))
        assert mx * s * reach <= 2.0 ** 62 and s == 2.0 ** round(np.log2(s))


def test_feature_column_contract():
    """Column names / order consumed by `classify` (obia/classification/classify.py:83)."""
    from obia_b200.segmentation.segment_statistics import _create_empty_stats_columns
    cols = _create_empty_stats_columns([0, 2], [0, 2], True, True, True, True, True, True,
                                       True, True, True, True, True, True, True, True, True, True, True)
    assert cols == (["segment_id"]
                    + [f"b{b}_{s}" for b in (0, 2) for s in ("mean", "variance", "min", "max", "skewness", "kurtosis")]
                    + [f"b{b}_{s}" for b in (0, 2) for s in ("contrast", "dissimilarity", "homogeneity", "ASM", "energy", "correlation")]
                    + ["pai", "fhd", "ch", "mean_intensity", "variance_intensity", "geometry"])
    cols = _create_empty_stats_columns([1], [], True, False, True, False, False, True,
                                       True, True, True, True, True, True, False, False, False, False, False)
    assert cols == ["segment_id", "b1_mean", "b1_min", "b1_kurtosis", "geometry"]


def test_stats_oracle_is_numpy_scipy():
    """The statistics oracle is literally np.mean/np.var/scipy.stats on the segment's pixels."""
    import stats_oracle
    from scipy.stats import kurtosis, skew
    rng = np.random.RandomState(0)
    labels = rng.randint(0, 5, size=(30, 40)).astype(np.int32)
    raw = rng.rand(30, 40, 3).astype(np.float32) * 100
    out, counts = stats_oracle.zonal_stats(labels, raw, [0, 2], np.arange(5))
    for i in range(5):
        px = raw[labels == i]
        assert counts[i] == len(px)
        np.testing.assert_allclose(out[i, 1, 0], np.mean(px[:, 2]), rtol=1e-6)
        np.testing.assert_allclose(out[i, 1, 1], np.var(px[:, 2]), rtol=1e-5)
        np.testing.assert_allclose(out[i, 0, 4], skew(px[:, 0].astype(np.float64)), rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(out[i, 0, 5], kurtosis(px[:, 0].astype(np.float64)), rtol=1e-3, atol=1e-3)


def test_combine_stats_matches_numpy_scipy():
    """Merging per-strip statistics (sharded multi-GPU path) == statistics of the concatenation."""
    import torch
    from scipy.stats import kurtosis, skew
    from obia_b200.sharded import combine_stats, split_rows
    rng = np.random.RandomState(0)

    def table(x):
        if len(x) == 0:
            return [0] + [np.nan] * 6 + [0]
        return [len(x), x.mean(), x.var(), x.min(), x.max(), skew(x), kurtosis(x), x.sum()]

    for parts in ([rng.rand(50) * 3 + 1, rng.rand(7) + 5, np.array([]), rng.rand(1) * 2],
                  [np.array([]), np.array([])], [np.full(5, 2.5), np.full(3, 2.5)],
                  [rng.normal(100, 1, 1000), rng.normal(100, 1, 3), rng.normal(100, 1, 40)]):
        T = [torch.tensor(table(p), dtype=torch.float64).view(1, 1, 8) for p in parts]
        out = combine_stats(T, resolution=1e-15)[0, 0].numpy()
        allx = np.concatenate(parts)
        ref = np.array(table(allx))
        if len(allx) and allx.var() == 0:
            ref[5] = ref[6] = np.nan          # scipy: (nearly) constant data -> NaN
        np.testing.assert_allclose(out, ref, rtol=1e-8, atol=1e-12, equal_nan=True)
    # strips start on multiples of the kernels' tile height and cover the raster exactly
    assert split_rows(10, 3, align=1) == [(0, 4), (4, 3), (7, 3)]
    assert split_rows(20000, 2) == [(0, 10112), (10112, 9888)]
    for H, world in ((80000, 8), (1101, 4), (40000, 8), (512, 4)):
        rows = split_rows(H, world)
        assert rows[0][0] == 0 and sum(h for _, h in rows) == H and all(r % 128 == 0 for r, _ in rows)
        assert all(rows[i][0] + rows[i][1] == rows[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        split_rows(300, 4)


def test_mask_sample_indices_equals_numpy_legacy_choice():
    """The C++ restatement of numpy's legacy RandomState(123).choice(replace=False) (MT19937, masked
    rejection, backward Fisher-Yates) is bit-identical to numpy, also through the background prefetch."""
    from obia_b200 import slic_host
    cases = [(1, 1), (2, 1), (3, 5), (10, 3), (400, 4), (1000, 7), (40000, 199), (65536, 10), (65537, 700),
             (123457, 1000), (623, 2), (624, 3), (625, 6), (67600, 100), (67600, 860), (250000, 40)]
    slic_host._CHOICE_CACHE.clear()
    slic_host.prefetch_mask_samples(cases[::2] + [(0, 3), (5, 0)])        # half of them in the background
    for n, k in cases:
        rng = np.random.RandomState(123)
        full = np.arange(n, dtype=int)
        want = (np.sort(rng.choice(full, min(k, n), replace=False)),
                np.sort(rng.choice(full, min(100 * k, n), replace=False)))
        got = slic_host.mask_sample_indices(n, k)
        np.testing.assert_array_equal(got[0], want[0])
        if 100 * k >= n:
            assert got[1] is None                      # every pixel: nothing drawn (sorted = arange)
            np.testing.assert_array_equal(want[1], np.arange(n))
        else:
            np.testing.assert_array_equal(got[1], want[1])
        assert slic_host.mask_sample_indices(n, k) is got                  # cached


def test_window_descriptor_layout_matches_the_cuda_header():
    """obia_b200/batch.py::WIN_DESC mirrors struct WinDesc of csrc/batch.cuh field by field (names, order, types)."""
    import os
    import re
    from obia_b200 import batch
    src = open(os.path.join(os.path.dirname(batch.__file__), "csrc", "batch.cuh")).read()
    body = src[src.index("struct WinDesc {"):src.index("};", src.index("struct WinDesc {"))]
    ctype = {"int32_t": "<i4", "float": "<f4", "double": "<f8", "int64_t": "<i8"}
    fields = []
    for line in body.splitlines()[1:]:
        line = line.split("//")[0].strip()
        m = re.match(r"(int32_t|int64_t|float|double)\s+([^;]+);", line)
        if m:
            fields += [(name.strip(), ctype[m.group(1)]) for name in m.group(2).split(",")]
    assert fields == [(n, batch.WIN_DESC.fields[n][0].str) for n in batch.WIN_DESC.names]
    assert batch.WIN_DESC.itemsize == 136 and "sizeof(WinDesc) == 136" in src
    offs = [batch.WIN_DESC.fields[n][1] for n in batch.WIN_DESC.names]
    assert offs == sorted(offs) and batch.WIN_DESC.fields["fix_scale"][1] % 8 == 0      # packed, doubles aligned


def test_batched_tiled_driver_option_gate():
    """Which SLIC options the batched tiled driver reproduces (the others take the per-tile driver)."""
    from obia_b200 import batch
    assert batch.supports({})
    assert batch.supports(dict(compactness=0.2, max_num_iter=5, start_label=0, convert2lab=False, sigma=0))
    for kw in (dict(sigma=1.0), dict(sigma=(1, 1)), dict(slic_zero=True), dict(exact=True), dict(enforce_connectivity=False),
               dict(spacing=(1, 2)), dict(start_label=2), dict(mask=None), dict(unknown_option=1)):
        assert not batch.supports(kw)


def test_parse_spacing():
    """slic's `spacing` handling for 2-D images (obia forwards **kwargs to skimage.segmentation.slic)."""
    from obia_b200 import slic_host
    assert slic_host.parse_spacing(None) == (1.0, 1.0)
    assert slic_host.parse_spacing([500, 1]) == (500.0, 1.0)
    assert slic_host.parse_spacing(np.array([0.5, 2.0])) == (0.5, 2.0)
    assert slic_host.parse_spacing((7, 0.25, 3)) == (0.25, 3.0)          # (z, y, x): z ignored for a 2-D image
    assert all(isinstance(v, np.float32) for v in slic_host.parse_spacing((1, 2)))
    with pytest.raises(TypeError):
        slic_host.parse_spacing(2.0)
    with pytest.raises(TypeError):
        slic_host.parse_spacing("12")
    for bad in ((1,), (1, 2, 3, 4), (0, 1), (1, -2), (np.inf, 1), (np.nan, 1)):
        with pytest.raises(ValueError):
            slic_host.parse_spacing(bad)

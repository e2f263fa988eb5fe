"""CPU tests of the GLCM texture oracle (oracle/texture_oracle.py).

scikit-image is absent here, so the restated graycomatrix / graycoprops are checked against
(i) the worked example of skimage's `graycomatrix` docstring and (ii) the property values of
skimage's own test-suite (skimage/feature/tests/test_texture.py) -- both recalled, see the
oracle header ("parity unpinned") -- and (iii) closed-form values of the direct pair sums the
CUDA kernel uses, so the algebra in csrc/texture.cu is pinned on the CPU as well.
"""
import numpy as np
import pytest

import texture_oracle as T

IMG = np.array([[0, 0, 1, 1], [0, 0, 1, 1], [0, 2, 2, 2], [2, 2, 3, 3]], dtype=np.uint8)


def test_graycomatrix_docstring_example():
    r = T.graycomatrix(IMG, [1], [0, np.pi / 4, np.pi / 2, 3 * np.pi / 4], levels=4)
    assert r[:, :, 0, 0].tolist() == [[2, 2, 1, 0], [0, 2, 0, 0], [0, 0, 3, 1], [0, 0, 0, 1]]
    assert r[:, :, 0, 1].tolist() == [[1, 1, 3, 0], [0, 1, 1, 0], [0, 0, 0, 2], [0, 0, 0, 0]]
    assert r[:, :, 0, 2].tolist() == [[3, 0, 2, 0], [0, 2, 2, 0], [0, 0, 1, 2], [0, 0, 0, 0]]
    assert r[:, :, 0, 3].tolist() == [[2, 0, 0, 0], [1, 1, 2, 0], [0, 0, 2, 1], [0, 0, 0, 0]]


def test_graycoprops_known_values():
    g = T.graycomatrix(IMG, [1, 2], [0], 4, normed=True, symmetric=True)
    np.testing.assert_almost_equal(T.graycoprops(g, "contrast")[0, 0], 0.58333333)
    np.testing.assert_almost_equal(T.graycoprops(g, "dissimilarity")[0, 0], 0.41666667)
    np.testing.assert_almost_equal(T.graycoprops(g, "homogeneity")[0, 0], 0.80833333)
    np.testing.assert_almost_equal(T.graycoprops(g, "energy")[0, 0], 0.38188131)
    np.testing.assert_almost_equal(T.graycoprops(g, "correlation")[0, 0], 0.71953255)
    np.testing.assert_almost_equal(T.graycoprops(g, "ASM")[0, 0], 0.38188131 ** 2, decimal=6)
    with pytest.raises(ValueError):
        T.graycoprops(g, "ABC")


def test_uniform_and_empty_matrices():
    im = np.ones((4, 4), dtype=np.uint8)
    g = T.graycomatrix(im, [1, 2], [0, np.pi / 2], 4, normed=True, symmetric=True)
    np.testing.assert_array_equal(T.graycoprops(g, "correlation"), 1.0)      # std == 0 -> 1
    np.testing.assert_array_equal(T.graycoprops(g, "contrast"), 0.0)
    # distance larger than the image: no pairs, all sums 0, correlation 1
    g = T.graycomatrix(im[:1, :2], [2], [0, np.pi / 2], 4, normed=True, symmetric=True)
    assert g.sum() == 0
    np.testing.assert_array_equal(T.graycoprops(g, "ASM"), 0.0)
    np.testing.assert_array_equal(T.graycoprops(g, "correlation"), 1.0)


def _pair_sum_features(q):
    """The closed forms csrc/texture.cu evaluates (integer pair sums), on the CPU."""
    h, w = q.shape
    q = q.astype(np.int64)
    out = np.zeros(6)
    for dr, dc in ((0, 2), (1, 1), (2, 0), (1, -1)):
        r0, r1, c0, c1 = 0, h - dr, max(0, -dc), min(w, w - dc)
        if r1 <= r0 or c1 <= c0:
            out[5] += 1.0
            continue
        i = q[r0:r1, c0:c1].ravel()
        j = q[r0 + dr:r1 + dr, c0 + dc:c1 + dc].ravel()
        N = i.size
        d = np.abs(i - j)
        lo, hi = np.minimum(i, j), np.maximum(i, j)
        _, u = np.unique(hi * 256 + lo, return_counts=True)
        diag = np.unique((hi * 256 + lo)[d == 0], return_counts=True)[1]
        sdiag, soff = int((diag ** 2).sum()), int((u ** 2).sum()) - int((diag ** 2).sum())
        asm = (soff + 2 * sdiag) / (2.0 * N * N)
        sa, sb, sc = int((i + j).sum()), int((i * i + j * j).sum()), int((i * j).sum())
        vnum = 2 * N * sb - sa * sa
        out += [(d ** 2).sum() / N, d.sum() / N, (1.0 / (1.0 + d ** 2)).sum() / N, asm, np.sqrt(asm),
                1.0 if vnum == 0 else (4 * N * sc - sa * sa) / vnum]
    return out / 4.0


@pytest.mark.parametrize("shape,seed", [((9, 13), 0), ((1, 7), 1), ((2, 2), 2), ((30, 5), 3), ((3, 1), 4)])
def test_pair_sum_algebra_matches_matrix_definition(shape, seed):
    rs = np.random.RandomState(seed)
    crop = (rs.rand(*shape) * (rs.rand(*shape) < 0.8)).astype(np.float32)      # zeros = outside pixels
    masked = np.where(crop > 0, crop, np.nan).astype(np.float32)[None]
    want = T.calculate_textural_stats(masked, [0])
    want = np.array([want[f"b0_{n}"] for n in T.TEXTURE_NAMES])
    valid = ~np.isnan(masked[0])
    if not valid.any():
        assert np.isnan(want).all()
        return
    clean = masked[0].copy()
    clean[~valid] = 0
    got = _pair_sum_features(T.quantise_band(clean))
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-15)


def test_textural_stats_per_label_and_nan_band():
    rs = np.random.RandomState(5)
    H, W = 24, 30
    labels = (np.arange(H)[:, None] // 8) * 3 + (np.arange(W)[None, :] // 10)
    labels = labels.astype(np.int32)
    labels[0:3, 0:4] = 7                              # makes label 0 non-rectangular
    raw = rs.rand(H, W, 2).astype(np.float32)
    raw[labels == 4, 1] = np.nan                       # a segment without a valid sample in band 1
    ids = np.unique(labels)
    out = T.textural_stats(labels, raw, [0, 1], ids)
    assert out.shape == (len(ids), 2, 6)
    assert np.isnan(out[list(ids).index(4), 1]).all() and not np.isnan(out[list(ids).index(4), 0]).any()
    assert ((out[:, 0, 3] > 0) & (out[:, 0, 3] <= 1)).all()                  # ASM in (0, 1]
    np.testing.assert_allclose(out[:, 0, 4] <= np.sqrt(out[:, 0, 3]) + 1e-12, True)   # mean sqrt <= sqrt mean


@pytest.mark.parametrize("name", ["texture_ms8_96", "texture_rgb_64"])
def test_texture_golden_fixtures(name):
    """Regression pins generated by tests/golden/make_golden.py (oracle output, see its header)."""
    import os
    gold = os.path.join(os.path.dirname(__file__), "golden")
    z = np.load(os.path.join(gold, name + ".npz"))
    src = np.load(os.path.join(gold, str(z["source"]) + ".npz"))
    got = T.textural_stats(src["labels"], src["raw"].astype(np.float32), z["bands"].tolist(), z["ids"],
                           compute_dtype=np.float64 if bool(z["quantise_f64"]) else np.float32)
    np.testing.assert_array_equal(got, z["texture"])

"""Batched tiled driver (BatchedTiledSegmenter: one launch per stage over all windows of a pass / tile-row)
against the per-tile driver (TiledSegmenter: one pipeline call per tile): same label raster, pixel for pixel.
Reference behaviour: /root/reference/obia/utils/tiling.py:62-291 (restated in oracle/tiling_oracle.py, against
which the per-tile driver is tested in tests/test_tiling.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _raster(H, W, C, seed):
    from gpu_helpers_cpu import synth_raster_cpu
    return synth_raster_cpu(H, W, C, seed=seed)


def _mask(H, W):
    yy, xx = np.mgrid[:H, :W]
    return (np.sin(yy / 45.0) + np.cos(xx / 35.0)) > -1.1


def _run(raw, mask, batched, **kw):
    from obia_b200.utils.tiling import create_tiled_segments
    labels, n, cols = create_tiled_segments(raw, None, mask, return_labels=True, polygons=False, batched=batched, **kw)
    torch.cuda.synchronize()
    return labels.cpu().numpy(), n


@pytest.mark.parametrize("shape,tile,buffer,kw", [
    ((620, 830, 4), 200, 30, dict(crown_radius=5, compactness=0.2)),                  # reference defaults, masked
    ((500, 700, 3), 160, 20, dict(crown_radius=7, compactness=5.0, max_num_iter=4)),  # Lab path
    ((450, 450, 5), 150, 40, dict(crown_radius=4, compactness=0.5, start_label=0)),
])
def test_batched_equals_per_tile_masked(shape, tile, buffer, kw):
    H, W, C = shape
    raw = _raster(H, W, C, seed=11)
    mask = _mask(H, W)
    a, na = _run(raw, mask, True, tile_size=tile, buffer=buffer, **kw)
    b, nb = _run(raw, mask, False, tile_size=tile, buffer=buffer, **kw)
    assert na == nb and na > 50
    assert np.array_equal(a, b)


@pytest.mark.parametrize("n_segments", [60, 400])
def test_batched_equals_per_tile_fixed_n(n_segments):
    raw = _raster(520, 760, 4, seed=5)
    a, na = _run(raw, None, True, tile_size=180, buffer=24, n_segments=n_segments, compactness=0.3)
    b, nb = _run(raw, None, False, tile_size=180, buffer=24, n_segments=n_segments, compactness=0.3)
    assert na == nb and na > 20
    assert np.array_equal(a, b)


def test_batched_empty_and_degenerate_windows():
    """Tiles whose mask is empty (ValueError in the reference -> skipped) and tiles with a constant band."""
    H, W = 600, 600
    raw = _raster(H, W, 3, seed=3)
    raw[:200, :200, 1] = 0.25                      # constant band inside one black tile
    mask = _mask(H, W)
    mask[200:400, 200:400] = False                 # an empty black tile
    mask[390:460, :130] &= (np.arange(130)[None, :] % 7 == 0)   # thin mask slivers
    a, na = _run(raw, mask, True, tile_size=200, buffer=30, crown_radius=5, compactness=0.2, convert2lab=False)
    b, nb = _run(raw, mask, False, tile_size=200, buffer=30, crown_radius=5, compactness=0.2, convert2lab=False)
    assert na == nb
    assert np.array_equal(a, b)


@pytest.mark.parametrize("with_mask", [False, True], ids=["no-mask", "user-mask"])
def test_batched_white_windows_without_earlier_segments(with_mask):
    """Every black tile fails (constant band inside the black tiles only -> the reference's swallowed ValueError), so
    the white windows of the first tile-row meet no earlier segment: the reference leaves their mask untouched
    (tiling.py:259-262: `None` without an input mask -> plain SLIC, no corner squares).  Later rows see the first
    row's segments in the corner overlaps and take the usual path."""
    H, W, T = 560, 600, 200
    raw = _raster(H, W, 3, seed=9)
    for j in range(0, H, T):
        for i in range(0, W, T):
            if (i // T + j // T) % 2 == 0:
                raw[j:j + T, i:i + T, 2] = 0.5
    mask = _mask(H, W) if with_mask else None
    kw = dict(tile_size=T, buffer=30, compactness=0.3, convert2lab=False)
    kw.update(dict(crown_radius=6) if with_mask else dict(n_segments=80))
    a, na = _run(raw, mask, True, **kw)
    b, nb = _run(raw, mask, False, **kw)
    assert na == nb and na > 20
    assert np.array_equal(a, b)
    # the black tile stays empty outside the reach of its white neighbours' buffers; the white tile is segmented
    assert (a[:T - 40, :T - 40] < 0).all() and (a[:T, T + 40:2 * T - 40] > 0).any()

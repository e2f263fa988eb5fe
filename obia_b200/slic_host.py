"""Host-side set-up of a SLIC run: grid geometry, centre initialisation and the
Gaussian taps.  O(n_segments) work that scikit-image also does in Python
before entering its Cython loop (reached from
/root/reference/obia/segmentation/segment_boundaries.py:51); the per-pixel work
is all in the CUDA library.

Restated from scikit-image (`skimage/util/_regular_grid.py::regular_grid`,
`skimage/segmentation/slic_superpixels.py::_get_grid_centroids`,
`::_get_mask_centroids`), see SURVEY.md section 3.4 step 4.
"""
from __future__ import annotations

import math

import numpy as np


def regular_grid_steps(shape_zyx, n_points):
    """(starts, steps) of skimage's `regular_grid`; step None -> every index.

    skimage sorts the dimensions, spreads `n_points` over the volume with a
    cubic-root step, then clamps dimensions that are thinner than the step
    (always the depth-1 axis for obia's 2-D rasters).
    """
    shape = np.asarray(shape_zyx)
    ndim = len(shape)
    order = np.argsort(shape)
    unsort = np.argsort(order)
    dims = shape[order]
    space = float(np.prod(shape))
    if space <= n_points:
        return [0] * ndim, [None] * ndim
    steps = np.full(ndim, (space / n_points) ** (1.0 / ndim), dtype=np.float64)
    if (dims < steps).any():
        for d in range(ndim):
            steps[d] = dims[d]
            space = float(np.prod(dims[d + 1:]))
            steps[d + 1:] = (space / n_points) ** (1.0 / (ndim - d - 1))
            if (dims >= steps).all():
                break
    starts = (steps // 2).astype(int)
    isteps = np.round(steps).astype(int)  # numpy rounds half to even, like skimage
    return [int(starts[i]) for i in unsort], [int(isteps[i]) for i in unsort]


def window_steps(H, W, n_centres):
    """Integer (step_y, step_x) that `_slic_cython` recomputes for its +-2*step windows."""
    _, steps = regular_grid_steps((1, H, W), n_centres)
    sy = 1 if steps[1] is None else steps[1]
    sx = 1 if steps[2] is None else steps[2]
    return int(sy), int(sx)


def grid_centroids(H, W, n_segments):
    """Unmasked initial centres: all grid points, y-major.  Returns (yx float64 (n,2), steps (3,))."""
    starts, steps = regular_grid_steps((1, H, W), n_segments)
    ys = np.arange(H)[slice(starts[1], None, steps[1])]
    xs = np.arange(W)[slice(starts[2], None, steps[2])]
    gy, gx = np.meshgrid(ys, xs, indexing="ij")
    yx = np.stack([gy.ravel(), gx.ravel()], axis=-1).astype(np.float64)
    fsteps = np.asarray([1.0 if s is None else float(s) for s in steps])
    return yx, fsteps


def mask_centroids(mask_hw, n_segments):
    """maskSLIC initial centres (RandomState(123) sampling + 5 k-means sweeps).

    Same scipy routines scikit-image calls (`kmeans2`, `pdist`).
    Returns (yx float64 (n,2), steps (3,) = mean |centroid - nearest centroid| per axis).
    """
    from scipy.cluster.vq import kmeans2
    from scipy.spatial import cKDTree

    mask3 = np.ascontiguousarray(mask_hw, dtype=bool)[np.newaxis]
    coord = np.array(np.nonzero(mask3), dtype=float).T
    if len(coord) == 0:
        # scikit-image fails on `image[mask].min()` of an empty selection
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    rng = np.random.RandomState(123)
    idx_full = np.arange(len(coord), dtype=int)
    idx = np.sort(rng.choice(idx_full, min(n_segments, len(coord)), replace=False))
    n_dense = int((10 ** 2) * n_segments)
    idx_dense = np.sort(rng.choice(idx_full, min(n_dense, len(coord)), replace=False))
    centroids, _ = kmeans2(coord[idx_dense], coord[idx], iter=5)
    # nearest other centroid: scikit-image builds the full O(n^2) pdist matrix and
    # takes argmin; a k-d tree gives the same neighbour for distinct distances
    if len(centroids) > 1:
        if len(centroids) <= 2048:
            from scipy.spatial.distance import pdist, squareform
            dist = squareform(pdist(centroids))
            np.fill_diagonal(dist, np.inf)
            closest = dist.argmin(-1)
        else:
            _, nn = cKDTree(centroids).query(centroids, k=2)
            closest = nn[:, 1]
        steps = np.abs(centroids - centroids[closest, :]).mean(0)
    else:
        # pdist of one point is empty; argmin of the 1x1 [[inf]] matrix is 0
        steps = np.abs(centroids - centroids[[0], :]).mean(0)
    return centroids[:, 1:3].copy(), steps


_CHOICE_CACHE = {}      # (n_coord, n_segments) -> (idx, idx_dense) or a Future of it
_CHOICE_POOL = None


def _draw_mask_samples(n_coord, n_segments):
    """The two RandomState(123) draws, bit for bit numpy's legacy algorithm, computed by the C++
    restatement in the library (`obia_b200_mask_sample_indices`): ctypes releases the GIL, so
    several draws run in parallel host threads."""
    import ctypes

    from . import _lib
    lib = _lib.load()
    idx = np.empty(min(n_segments, n_coord), dtype=np.int64)
    # 100 * n_segments >= n_coord: the second draw is every pixel (returned as None, nothing to compute)
    idx_dense = None if 100 * n_segments >= n_coord else np.empty(100 * n_segments, dtype=np.int64)
    _lib.check(lib.obia_b200_mask_sample_indices(int(n_coord), int(n_segments),
                                                 idx.ctypes.data_as(ctypes.c_void_p),
                                                 None if idx_dense is None else idx_dense.ctypes.data_as(ctypes.c_void_p)),
               "mask_sample_indices")
    return idx, idx_dense


def prefetch_mask_samples(keys):
    """Start the draws for `keys` = iterable of (n_coord, n_segments) on background threads; a later
    `mask_sample_indices` call for the same key waits for that result instead of recomputing it.
    The legacy permutation is O(n_coord) and sequential (about 0.1 s for a 2000 x 2000 tile), which
    would otherwise serialise with the GPU work of every tile whose mask count is new."""
    global _CHOICE_POOL
    import os
    from concurrent.futures import ThreadPoolExecutor
    if _CHOICE_POOL is None:
        # one process per GPU under torchrun: share the host cores between the local ranks
        local = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1))
        workers = max(1, min(32, ((os.cpu_count() or 2) - 1) // local))
        _CHOICE_POOL = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="obia_b200_rng")
    for n_coord, n_segments in keys:
        key = (int(n_coord), int(n_segments))
        if key[0] <= 0 or key[1] <= 0 or key in _CHOICE_CACHE:
            continue
        if len(_CHOICE_CACHE) > 4096:
            _CHOICE_CACHE.clear()
        _CHOICE_CACHE[key] = _CHOICE_POOL.submit(_draw_mask_samples, *key)


def submit_mask_samples(keys):
    """Futures of the draws for `keys` (iterable of (n_coord, n_segments)), one per distinct key, owned by the
    caller (the window batches of the tiled driver: tens of thousands of distinct keys per run, more than the
    shared cache keeps)."""
    global _CHOICE_POOL
    import os
    from concurrent.futures import ThreadPoolExecutor
    if _CHOICE_POOL is None:
        local = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1))
        workers = max(1, min(32, ((os.cpu_count() or 2) - 1) // local))
        _CHOICE_POOL = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="obia_b200_rng")
    out = {}
    for n_coord, n_segments in keys:
        key = (int(n_coord), int(n_segments))
        if key[0] <= 0 or key[1] <= 0 or key in out:
            continue
        hit = _CHOICE_CACHE.get(key)
        out[key] = hit if hit is not None else _CHOICE_POOL.submit(_draw_mask_samples, *key)
    return out


def mask_sample_indices(n_coord, n_segments):
    """(idx, idx_dense) of `_get_mask_centroids`: two draws from RandomState(123); idx_dense is None when it
    is every pixel (100 * n_segments >= n_coord).

    They depend only on the number of mask pixels and n_segments, so tiles with equal counts
    share them (cached; the legacy permutation is O(n_coord) on the host).
    """
    key = (int(n_coord), int(n_segments))
    hit = _CHOICE_CACHE.get(key)
    if hit is None:
        if len(_CHOICE_CACHE) > 4096:
            _CHOICE_CACHE.clear()
        hit = _CHOICE_CACHE[key] = _draw_mask_samples(*key)
    elif not isinstance(hit, tuple):
        hit = _CHOICE_CACHE[key] = hit.result()
    return hit


def steps_from_centroids(centroids_yx, closest=None):
    """`steps` of `_get_mask_centroids`: mean |centroid - nearest other centroid| per axis (z, y, x).

    `closest`: index of the nearest other centroid when the caller already has it (device search).
    """
    c = np.concatenate([np.zeros((len(centroids_yx), 1)), np.asarray(centroids_yx, dtype=np.float64)], axis=1)
    if closest is not None and len(c) > 1:
        return np.abs(c - c[np.asarray(closest, dtype=np.int64), :]).mean(0)
    if len(c) > 1:
        if len(c) <= 2048:
            from scipy.spatial.distance import pdist, squareform
            dist = squareform(pdist(c))
            np.fill_diagonal(dist, np.inf)
            closest = dist.argmin(-1)
        else:
            from scipy.spatial import cKDTree
            _, nn = cKDTree(c).query(c, k=2)
            closest = nn[:, 1]
        return np.abs(c - c[closest, :]).mean(0)
    return np.abs(c - c[[0], :]).mean(0)


def gaussian_taps(sigma):
    """scipy.ndimage `_gaussian_kernel1d(sigma, 0, int(4*sigma+0.5))` (symmetric, float64)."""
    sd = float(sigma)
    radius = int(4.0 * sd + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sd * sd) * x ** 2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi, dtype=np.float64), radius


def fixed_point_scale(max_abs_value, H, W, step_y, step_x):
    """Power-of-two scale so that per-centre colour sums fit 62 bits."""
    reach = min(H * W, (4 * step_y + 1) * (4 * step_x + 1))
    bits_px = max(1, math.ceil(math.log2(reach + 1)))
    bits_val = math.ceil(math.log2(max(max_abs_value, 1e-30))) + 1
    shift = 62 - bits_px - bits_val
    return float(2.0 ** shift)


def parse_spacing(spacing):
    """(sy, sx) float32 of skimage.segmentation.slic's `spacing` argument for a 2-D image: None -> (1, 1); an
    iterable of two values, or of three with the leading (z) one ignored as slic still accepts with a
    FutureWarning; anything else raises like slic does."""
    if spacing is None:
        return np.float32(1.0), np.float32(1.0)
    if isinstance(spacing, (str, bytes)) or not hasattr(spacing, "__iter__"):
        raise TypeError("spacing must be None or iterable.")
    sp = np.asarray(list(spacing), dtype=np.float32).ravel()
    if sp.size not in (2, 3):
        raise ValueError(f"Input image is 2D, but spacing has {sp.size} elements (expected 2).")
    sp_y, sp_x = sp[-2], sp[-1]
    if not (np.isfinite(sp_y) and np.isfinite(sp_x) and sp_y > 0 and sp_x > 0):
        raise ValueError("spacing must be positive and finite")
    return sp_y, sp_x

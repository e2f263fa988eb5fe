"""CPU oracle for `method="quickshift"` -- TEST INFRASTRUCTURE ONLY (parity unpinned).

The reference calls `skimage.segmentation.quickshift(img_to_segment, **kwargs)`
(/root/reference/obia/segmentation/segment_boundaries.py:48-49) on the band-selected, per-band
normalised float32 raster (:31-43).  scikit-image (>= 0.23.2, pyproject.toml:23) is a third-party
dependency that is neither vendored under /root/reference nor installable here, so its published
algorithm is restated (skimage/segmentation/_quickshift.py::quickshift and
_quickshift_cython.pyx::_quickshift_cython, as recalled; the reference holds no fixture for it):

  1. img_as_float; RGB -> CIELAB when `convert2lab` (3 bands required); reflecting Gaussian with
     `sigma` on both spatial axes; multiply by `ratio`.
  2. density of every pixel = sum over the (2w+1)^2 window, w = ceil(3 * kernel_size), clipped to the
     raster, visited in row-major order, of exp(-dist / (2 kernel_size^2)) with dist = squared feature
     distance + squared row offset + squared column offset; float32 `dist`, libm double `exp`, the sum
     kept in float32 (the Cython memoryview element type for float32 input).
  3. densities += default_rng(rng).normal(scale=1e-5, size=(H, W))  (tie breaking).
  4. parent of a pixel = the window pixel of strictly higher density at the smallest dist (first one
     in row-major order among equals); no such pixel, or sqrt(dist) > max_dist -> its own root.
  5. labels = rank of every pixel's root among the sorted root indices (np.unique(..., return_inverse)).

Only tests/ may import this module.
"""
from __future__ import annotations

import math

import numpy as np


def quickshift(image, ratio=1.0, kernel_size=5, max_dist=10, return_tree=False, sigma=0, convert2lab=True, rng=42):
    import slic_oracle as so
    from scipy.ndimage import gaussian_filter
    image = np.atleast_3d(np.asarray(image))
    if image.dtype.kind in "ui":
        image = image.astype(np.float64) / np.iinfo(image.dtype).max
    dt = np.float32 if image.dtype in (np.float16, np.float32) else np.float64
    image = image.astype(dt, copy=False)
    if convert2lab:
        if image.shape[-1] != 3:
            raise ValueError("Only RGB images can be converted to Lab space.")
        image = so.rgb2lab(image)
    if kernel_size < 1:
        raise ValueError("`kernel_size` should be >= 1.")
    if sigma > 0:
        image = np.stack([gaussian_filter(image[..., c], sigma=sigma, mode="reflect") for c in range(image.shape[-1])], -1)
    image = np.ascontiguousarray(image * dt(ratio)).astype(dt)
    if return_tree:
        raise NotImplementedError
    H, W, C = image.shape
    ks = dt(kernel_size)
    inv = dt(-0.5) / (ks * ks)
    kw = int(math.ceil(3 * kernel_size))
    dens = np.zeros((H, W), dtype=dt)

    def window_dist(dr, dc):
        """(slices of the centre pixels, dist array) for the neighbour offset (dr, dc)."""
        r0, r1 = max(0, -dr), min(H, H - dr)
        c0, c1 = max(0, -dc), min(W, W - dc)
        if r1 <= r0 or c1 <= c0:
            return None
        a = image[r0:r1, c0:c1]
        b = image[r0 + dr:r1 + dr, c0 + dc:c1 + dc]
        dist = np.zeros(a.shape[:2], dtype=dt)
        for ch in range(C):
            t = a[..., ch] - b[..., ch]
            dist = dist + t * t
        dist = dist + dt(dr * dr)
        dist = dist + dt(dc * dc)
        return (slice(r0, r1), slice(c0, c1)), dist

    for dr in range(-kw, kw + 1):
        for dc in range(-kw, kw + 1):
            wd = window_dist(dr, dc)
            if wd is None:
                continue
            sl, dist = wd
            e = np.exp((dist * inv).astype(np.float64))
            dens[sl] = (dens[sl].astype(np.float64) + e).astype(dt)
    noise = np.random.default_rng(rng).normal(scale=0.00001, size=(H, W))
    dens = (dens.astype(np.float64) + noise).astype(dt)

    parent = np.arange(H * W, dtype=np.int64).reshape(H, W)
    closest = np.full((H, W), np.inf, dtype=dt)
    for dr in range(-kw, kw + 1):
        for dc in range(-kw, kw + 1):
            wd = window_dist(dr, dc)
            if wd is None:
                continue
            (rs, cs), dist = wd
            nb_d = dens[rs.start + dr:rs.stop + dr, cs.start + dc:cs.stop + dc]
            better = (nb_d > dens[rs, cs]) & (dist < closest[rs, cs])
            idx = (np.arange(rs.start, rs.stop)[:, None] + dr) * W + (np.arange(cs.start, cs.stop)[None, :] + dc)
            parent[rs, cs] = np.where(better, idx, parent[rs, cs])
            closest[rs, cs] = np.where(better, dist, closest[rs, cs])
    flat = parent.ravel().copy()
    too_far = np.sqrt(closest).ravel() > dt(max_dist)
    flat[too_far] = np.arange(H * W)[too_far]
    old = np.zeros_like(flat)
    while (old != flat).any():
        old = flat
        flat = flat[flat]
    return np.unique(flat, return_inverse=True)[1].reshape(H, W)


def create_segments_labels(raw_hwc, bands=None, **kwargs):
    """obia's wrapper (segment_boundaries.py:31-49): normalise every band, select, quickshift."""
    import slic_oracle as so
    img = np.asarray(raw_hwc, dtype=np.float32).copy()
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(img.shape[2]):
            img[:, :, i] = so.normalize_band(img[:, :, i])
    if bands is None:
        bands = list(range(img.shape[2]))
    return quickshift(img[:, :, bands], **kwargs)

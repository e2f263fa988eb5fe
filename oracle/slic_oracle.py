"""CPU oracle for the SLIC half of obia's hot path -- TEST INFRASTRUCTURE ONLY.

Restates, on the CPU with numpy/scipy + the C loops in slic_core.c, what
`obia.segmentation.segment_boundaries.create_segments`
(/root/reference/obia/segmentation/segment_boundaries.py:18-57) computes up
to the label raster:

    per-band min-max normalise (in place)        segment_boundaries.py:11-16, 31-33
    band selection / img_as_float                segment_boundaries.py:35-43
    skimage.segmentation.slic(img, **kwargs)     segment_boundaries.py:51
    segments[mask == 0] = -1                     segment_boundaries.py:55-57

The `slic` arithmetic lives in scikit-image (>=0.23.2, pyproject.toml:23; not
vendored, not installable here).  `slic()` below restates the published
`skimage/segmentation/slic_superpixels.py::slic`, `_get_grid_centroids`,
`_get_mask_centroids`, `skimage/util/_regular_grid.py::regular_grid` and
`skimage/color/colorconv.py::rgb2lab` following SURVEY.md section 3.4; the
pieces scikit-image itself delegates to scipy (`gaussian_filter`, `kmeans2`,
`pdist`) call the real scipy, which IS installed.

PARITY UNPINNED: the reference has no tests / golden vectors for this path
and scikit-image cannot be imported in this image.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module.
"""
from __future__ import annotations

import ctypes
import math
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        import build as _oracle_build  # oracle/build.py
        path = _oracle_build.build()
        lib = ctypes.CDLL(path)
        i64, f32 = ctypes.c_int64, ctypes.c_float
        vp = ctypes.c_void_p
        lib.obia_oracle_slic_core.restype = i64
        lib.obia_oracle_slic_core.argtypes = [vp, vp, vp, i64, i64, i64, i64, f32,
                                              i64, vp, ctypes.c_int, i64,
                                              ctypes.c_int, i64, i64, vp]
        lib.obia_oracle_slic_core_f64.restype = i64
        lib.obia_oracle_slic_core_f64.argtypes = lib.obia_oracle_slic_core.argtypes
        lib.obia_oracle_slic_core_fma.restype = i64
        lib.obia_oracle_slic_core_fma.argtypes = lib.obia_oracle_slic_core.argtypes
        lib.obia_oracle_slic_assign_once_fma.restype = None
        lib.obia_oracle_slic_assign_once.restype = None
        lib.obia_oracle_slic_assign_once.argtypes = [vp, vp, vp, i64, i64, i64, i64,
                                                     f32, i64, ctypes.c_int, i64,
                                                     i64, vp, vp]
        lib.obia_oracle_slic_assign_once_fma.argtypes = lib.obia_oracle_slic_assign_once.argtypes
        lib.obia_oracle_enforce_connectivity.restype = ctypes.c_int
        lib.obia_oracle_enforce_connectivity.argtypes = [vp, i64, i64, i64, i64, i64, vp]
        _LIB = lib
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------
# skimage.util.regular_grid  [UPSTREAM restatement, SURVEY.md 3.4 step 4]
# --------------------------------------------------------------------------
def regular_grid(ar_shape, n_points):
    ar_shape = np.asanyarray(ar_shape)
    ndim = len(ar_shape)
    unsort_dim_idxs = np.argsort(np.argsort(ar_shape))
    sorted_dims = np.sort(ar_shape)
    space_size = float(np.prod(ar_shape))
    if space_size <= n_points:
        return (slice(None),) * ndim
    stepsizes = np.full(ndim, (space_size / n_points) ** (1.0 / ndim), dtype="float64")
    if (sorted_dims < stepsizes).any():
        for dim in range(ndim):
            stepsizes[dim] = sorted_dims[dim]
            space_size = float(np.prod(sorted_dims[dim + 1:]))
            stepsizes[dim + 1:] = (space_size / n_points) ** (1.0 / (ndim - dim - 1))
            if (sorted_dims >= stepsizes).all():
                break
    starts = (stepsizes // 2).astype(int)
    stepsizes = np.round(stepsizes).astype(int)
    slices = [slice(start, None, step) for start, step in zip(starts, stepsizes)]
    slices = tuple(slices[i] for i in unsort_dim_idxs)
    return slices


def grid_steps(shape_zyx, n_points):
    """Integer (step_z, step_y, step_x) as `_slic_cython` recomputes them."""
    return tuple(int(s.step if s.step is not None else 1)
                 for s in regular_grid(shape_zyx, n_points))


def _get_grid_centroids(shape_zyx, n_centroids):
    d, h, w = shape_zyx
    slices = regular_grid(shape_zyx, n_centroids)
    zs = np.arange(d)[slices[0]]
    ys = np.arange(h)[slices[1]]
    xs = np.arange(w)[slices[2]]
    gz, gy, gx = np.meshgrid(zs, ys, xs, indexing="ij")
    centroids = np.stack([gz.ravel(), gy.ravel(), gx.ravel()], axis=-1)
    steps = np.asarray([float(s.step) if s.step is not None else 1.0 for s in slices])
    return centroids, steps


def _get_mask_centroids(mask_zyx, n_centroids, multichannel=True):
    """maskSLIC initialisation (masked branch)."""
    from scipy.cluster.vq import kmeans2
    from scipy.spatial.distance import pdist, squareform

    coord = np.array(np.nonzero(mask_zyx), dtype=float).T
    rng = np.random.RandomState(123)
    idx_full = np.arange(len(coord), dtype=int)
    idx = np.sort(rng.choice(idx_full, min(n_centroids, len(coord)), replace=False))
    dense_factor = 10
    ndim_spatial = mask_zyx.ndim - 1 if multichannel else mask_zyx.ndim
    n_dense = int((dense_factor ** ndim_spatial) * n_centroids)
    idx_dense = np.sort(rng.choice(idx_full, min(n_dense, len(coord)), replace=False))
    centroids, _ = kmeans2(coord[idx_dense], coord[idx], iter=5)
    dist = squareform(pdist(centroids))
    np.fill_diagonal(dist, np.inf)
    closest_pts = dist.argmin(-1)
    steps = abs(centroids - centroids[closest_pts, :]).mean(0)
    return centroids, steps


# --------------------------------------------------------------------------
# skimage.color.rgb2lab (float32 in -> float32 out)  [UPSTREAM restatement]
# --------------------------------------------------------------------------
_XYZ_FROM_RGB = np.array([[0.412453, 0.357580, 0.180423],
                          [0.212671, 0.715160, 0.072169],
                          [0.019334, 0.119193, 0.950227]])
_D65_2 = (0.95047, 1.0, 1.08883)


def rgb2lab(rgb):
    arr = np.array(rgb, copy=True)
    dt = arr.dtype
    m = arr > 0.04045
    arr[m] = np.power((arr[m] + 0.055) / 1.055, 2.4)
    arr[~m] /= 12.92
    xyz = arr @ _XYZ_FROM_RGB.T.astype(dt)
    xyz = xyz / np.asarray(_D65_2, dtype=dt)
    m = xyz > 0.008856
    xyz[m] = np.cbrt(xyz[m])
    xyz[~m] = 7.787 * xyz[~m] + 16.0 / 116.0
    x, y, z = xyz[..., 0], xyz[..., 1], xyz[..., 2]
    L = (116.0 * y) - 16.0
    a = 500.0 * (x - y)
    b = 200.0 * (y - z)
    return np.concatenate([v[..., np.newaxis] for v in (L, a, b)], axis=-1).astype(dt, copy=False)


# --------------------------------------------------------------------------
# the two Cython loops (C core)
# --------------------------------------------------------------------------
# Which float32 build of the C core the module uses: False = separate multiply/add in the
# colour term (x86-64 scikit-image wheels), True = fused multiply-add (arm64 wheels and the
# CUDA kernel).  Tests flip it to check the GPU bit-for-bit (fma) and statistically (both).
USE_FMA = False


def slic_core(image_hwc, mask_hw, segments, step, max_num_iter, spacing,
              slic_zero, start_label, ignore_color, fma=None):
    """`_slic_cython` for depth 1.  `segments` (n, 3+C) float32 is updated in place."""
    fma = USE_FMA if fma is None else fma
    H, W, C = image_hwc.shape
    n = segments.shape[0]
    _, step_y, step_x = grid_steps((1, H, W), n)
    nearest = np.full((H, W), start_label - 1, dtype=np.int64)
    dt = segments.dtype
    assert dt in (np.float32, np.float64) and segments.flags.c_contiguous
    image_hwc = np.ascontiguousarray(image_hwc, dtype=dt)
    spacing = np.ascontiguousarray(spacing, dtype=dt)
    mask_u8 = None if mask_hw is None else np.ascontiguousarray(mask_hw, dtype=np.uint8)
    if dt == np.float32:
        fn = _lib().obia_oracle_slic_core_fma if fma else _lib().obia_oracle_slic_core
    else:
        fn = _lib().obia_oracle_slic_core_f64
    rc = fn(_ptr(image_hwc), _ptr(mask_u8), _ptr(segments),
                                      H, W, C, n, float(step), int(max_num_iter),
                                      _ptr(spacing), int(bool(slic_zero)),
                                      int(start_label), int(bool(ignore_color)),
                                      step_y, step_x, _ptr(nearest))
    if rc < 0:
        raise MemoryError("oracle slic_core allocation failed")
    return nearest


def slic_assign_once(image_hwc, mask_hw, segments, step, start_label, ignore_color, fma=None):
    """One assignment sweep (no update) -> (labels int64, distance float32)."""
    fma = USE_FMA if fma is None else fma
    H, W, C = image_hwc.shape
    n = segments.shape[0]
    _, step_y, step_x = grid_steps((1, H, W), n)
    nearest = np.full((H, W), start_label - 1, dtype=np.int64)
    distance = np.empty((H, W), dtype=np.float32)
    image_hwc = np.ascontiguousarray(image_hwc, dtype=np.float32)
    segments = np.ascontiguousarray(segments, dtype=np.float32)
    mask_u8 = None if mask_hw is None else np.ascontiguousarray(mask_hw, dtype=np.uint8)
    fn = _lib().obia_oracle_slic_assign_once_fma if fma else _lib().obia_oracle_slic_assign_once
    fn(_ptr(image_hwc), _ptr(mask_u8), _ptr(segments),
                                        H, W, C, n, float(step), int(start_label),
                                        int(bool(ignore_color)), step_y, step_x,
                                        _ptr(nearest), _ptr(distance))
    return nearest, distance


def enforce_connectivity(labels_hw, min_size, max_size, start_label=1):
    """`_enforce_label_connectivity_cython` for depth 1."""
    labels = np.ascontiguousarray(labels_hw, dtype=np.int64)
    H, W = labels.shape
    out = np.empty_like(labels)
    rc = _lib().obia_oracle_enforce_connectivity(_ptr(labels), H, W, int(min_size),
                                                 int(max_size), int(start_label), _ptr(out))
    if rc != 0:
        raise ValueError("max_size must be >= 1")
    return out


# --------------------------------------------------------------------------
# skimage.segmentation.slic  [UPSTREAM restatement, SURVEY.md 3.4 steps 1-10]
# --------------------------------------------------------------------------
def slic(image, n_segments=100, compactness=10.0, max_num_iter=10, sigma=0,
         spacing=None, convert2lab=None, enforce_connectivity_=True,
         min_size_factor=0.5, max_size_factor=3, slic_zero=False,
         start_label=1, mask=None, *, channel_axis=-1, return_state=False,
         **kw):
    if "enforce_connectivity" in kw:
        enforce_connectivity_ = kw.pop("enforce_connectivity")
    if kw:
        raise TypeError(f"unexpected kwargs {sorted(kw)}")
    image = np.asarray(image)
    if image.ndim == 2 and channel_axis is not None:
        raise ValueError("channel_axis=-1 indicates multichannel, which is not "
                         "supported for a two-dimensional image")
    if image.dtype.kind in "ui":
        image = image.astype(np.float64) / np.iinfo(image.dtype).max  # img_as_float
    float_dtype = np.float32 if image.dtype in (np.float16, np.float32) else np.float64
    image = image.astype(float_dtype, copy=True)
    multichannel = channel_axis is not None
    if multichannel and channel_axis not in (-1, image.ndim - 1):
        raise NotImplementedError("oracle supports channel_axis in (-1, None)")
    if image.ndim not in (2, 3) or (image.ndim == 3 and not multichannel):
        raise NotImplementedError("oracle supports 2-D rasters only (obia passes H,W,C)")

    use_mask = mask is not None
    if use_mask:
        mask = np.ascontiguousarray(mask, dtype=bool)
        if multichannel:
            mask_ = np.broadcast_to(np.expand_dims(mask, axis=-1), image.shape)
        else:
            mask_ = mask
        image_values = image[mask_]
    else:
        image_values = image

    imin = image_values.min()
    imax = image_values.max()
    if np.isnan(imin):
        raise ValueError("unmasked NaN values in image are not supported")
    if np.isinf(imin) or np.isinf(imax):
        raise ValueError("unmasked infinite values in image are not supported")
    image -= imin
    if imax != imin:
        image /= (imax - imin)

    dtype = image.dtype
    if image.ndim == 2:
        image = image[np.newaxis, ..., np.newaxis]
    else:
        image = image[np.newaxis, ...]

    if multichannel and (convert2lab or convert2lab is None):
        if image.shape[-1] != 3 and convert2lab:
            raise ValueError("Lab colorspace conversion requires a RGB image.")
        elif image.shape[-1] == 3:
            image = rgb2lab(image)

    if start_label not in [0, 1]:
        raise ValueError("start_label should be 0 or 1.")

    update_centroids = False
    if use_mask:
        mask3 = np.ascontiguousarray(mask[np.newaxis, ...]).view("uint8")
        if mask3.shape != image.shape[:3]:
            raise ValueError("image and mask should have the same shape.")
        centroids, steps = _get_mask_centroids(mask3, n_segments, multichannel)
        update_centroids = True
    else:
        mask3 = None
        centroids, steps = _get_grid_centroids(image.shape[:3], n_segments)

    if spacing is None:
        spacing = np.ones(3, dtype=dtype)
    else:
        spacing = np.asarray(spacing, dtype=dtype)
        if spacing.shape == (2,):
            spacing = np.insert(spacing, 0, 1)
    if not isinstance(sigma, (list, tuple, np.ndarray)):
        sigma = np.array([sigma, sigma, sigma], dtype=dtype)
        sigma /= spacing
    else:
        sigma = np.asarray(sigma, dtype=dtype)
        if sigma.shape == (2,):
            sigma = np.insert(sigma, 0, 0)
        sigma = sigma / spacing
    if (sigma > 0).any():
        from scipy import ndimage as ndi
        # skimage.filters.gaussian(image, sigma=[sz,sy,sx,0], mode='reflect')
        image = ndi.gaussian_filter(image, list(sigma) + [0], mode="reflect",
                                    truncate=4.0).astype(dtype, copy=False)

    n_centroids = centroids.shape[0]
    segments = np.ascontiguousarray(
        np.concatenate([centroids, np.zeros((n_centroids, image.shape[3]))], axis=-1),
        dtype=dtype)
    step = max(steps)
    ratio = 1.0 / compactness
    image = np.ascontiguousarray(image * ratio, dtype=dtype)

    mask_hw = None if mask3 is None else mask3[0]
    img_hwc = image[0]
    if update_centroids:
        slic_core(img_hwc, mask_hw, segments, step, max_num_iter, spacing, slic_zero,
                  start_label, ignore_color=True)
    labels = slic_core(img_hwc, mask_hw, segments, step, max_num_iter, spacing,
                       slic_zero, start_label, ignore_color=False)
    pre_cc = labels
    if enforce_connectivity_:
        if use_mask:
            segment_size = mask3.sum() / n_centroids
        else:
            segment_size = math.prod(image.shape[:3]) / n_centroids
        min_size = int(min_size_factor * segment_size)
        max_size = int(max_size_factor * segment_size)
        labels = enforce_connectivity(labels, min_size, max_size, start_label=start_label)
    if return_state:
        return labels, dict(features=img_hwc, centres=segments, step=float(step),
                            pre_connectivity=pre_cc, n_centroids=n_centroids,
                            init_centroids=centroids, steps=steps)
    return labels


# --------------------------------------------------------------------------
# obia wrapper semantics (label-raster part of create_segments)
# --------------------------------------------------------------------------
def normalize_band(band):
    """segment_boundaries.py:11-16."""
    return (band - np.min(band)) / (np.max(band) - np.min(band))


def create_segments_labels(img_data, segmentation_bands=None, **kwargs):
    """segment_boundaries.py:31-57 up to the label raster.

    Mutates `img_data` in place exactly like the reference (every band
    normalised).  Returns the int64 label raster with masked pixels = -1.
    """
    num_bands = img_data.shape[2]
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(num_bands):
            img_data[:, :, i] = normalize_band(img_data[:, :, i])
    if segmentation_bands is None:
        segmentation_bands = list(range(num_bands))
    for band in segmentation_bands:
        if band >= num_bands or band < 0:
            raise IndexError(f"Band index {band} out of range. Available bands indices: 0 to {num_bands - 1}.")
    img_to_segment = img_data[:, :, segmentation_bands]
    segments = slic(img_to_segment, **kwargs)
    mask = kwargs.get("mask", None)
    if mask is not None:
        segments[np.asarray(mask) == 0] = -1
    return segments

"""Device pipeline: raw raster (H, W, C) float32 in HBM -> SLIC label raster ->
per-segment zonal statistics, all through the C ABI of libobia_b200.so.

PyTorch is used for device memory, streams and (elsewhere) torch.distributed
only; every per-pixel operation is a hand-written sm_100a kernel.  There is no
CPU path here: a missing library or a non-CUDA tensor raises.

Reference path being replaced (file:line under /root/reference):
  obia/segmentation/segment_boundaries.py:31-57  (normalise, select, slic, mask)
  obia/segmentation/segment_statistics.py:113-176, 475-508 (spectral statistics)
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib, slic_host

STAT_FIELDS = ("count", "mean", "variance", "min", "max", "skewness", "kurtosis", "sum")
TEXTURE_FIELDS = ("contrast", "dissimilarity", "homogeneity", "ASM", "energy", "correlation")


def _stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _require_cuda(t, name, dtype=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (obia_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def _aligned(t):
    """The kernels use 128-bit loads: a view that starts off a 16-byte boundary (e.g. a row slice of
    a raster whose row size is not a multiple of 16 bytes) is copied to a fresh allocation."""
    return t if t.data_ptr() % 16 == 0 else t.clone()


# Scratch memory kept by this module instead of being handed back to torch's caching allocator after every call.
# The connectivity workspace is one 41 B/pixel block (4.4 GB on c2); once it is back in the allocator's pool the next
# 0.4 GB request is carved out of it and the following step pays a fresh cudaMalloc of several GB (6 - 100 ms,
# measured on the strip path) while the reserved pool keeps growing.  One buffer per (device, tag), grow-only.
_SCRATCH = {}
_SCRATCH_MAX_FRACTION = 0.2      # blocks beyond this share of the device memory are not kept


def scratch(device, nbytes, tag):
    """uint8 device buffer of `nbytes` for the duration of one call (one stream per process, as everywhere in this
    package: two concurrent calls with the same tag would share it).  `release_scratch()` frees them."""
    device = torch.device(device)
    nbytes = int(nbytes)
    if nbytes > _SCRATCH_MAX_FRACTION * torch.cuda.get_device_properties(device).total_memory:
        return torch.empty((nbytes,), dtype=torch.uint8, device=device)
    key = (device.index if device.index is not None else torch.cuda.current_device(), tag)
    buf = _SCRATCH.get(key)
    if buf is None or buf.numel() < nbytes:
        _SCRATCH.pop(key, None)
        buf = None                      # the smaller block goes back before the larger one is requested
        buf = torch.empty((nbytes,), dtype=torch.uint8, device=device)
        _SCRATCH[key] = buf
    return buf[:nbytes]


def release_scratch():
    _SCRATCH.clear()


def _i32_array(values):
    arr = (ctypes.c_int32 * len(values))(*[int(v) for v in values])
    return arr


# ------------------------------------------------------------------ K1 ------
def band_minmax(raw, mask=None):
    """Per-band (min, max, masked min, masked max) and non-finite flags, on the host.

    One pass over the raster; the tiny result is copied back (one stream sync)
    because the host needs it to validate the input the way numpy/skimage do.
    """
    lib = _lib.load()
    raw = _aligned(_require_cuda(raw, "raw", torch.float32))
    H, W, C = raw.shape
    out = torch.empty((C, 4), dtype=torch.float32, device=raw.device)
    flags = torch.empty((C,), dtype=torch.int32, device=raw.device)
    if mask is not None:
        _require_cuda(mask, "mask", torch.uint8)
    _lib.check(lib.obia_b200_band_minmax(_p(raw), H * W, C, _p(mask), _p(out), _p(flags), _stream_ptr()),
               "band_minmax")
    return out, flags


def normalize_to_host(raw, host, minmax_dev, max_ctas=16):
    """Normalised copy of `raw` written by the kernel straight into the page-locked host tensor
    `host` (same shape, float32, contiguous); see obia_b200_normalize_to.  Asynchronous.
    16 CTAs keep PCIe busy and cost the concurrently running SLIC kernels the least (measured on
    c2: 8 / 16 / 64 / 148 CTAs -> 152 / 149 / 160 / 190 ms end to end)."""
    lib = _lib.load()
    _require_cuda(raw, "raw", torch.float32)
    H, W, C = raw.shape
    if not (host.is_pinned() and host.is_contiguous() and host.dtype == torch.float32
            and tuple(host.shape) == tuple(raw.shape)):
        raise ValueError("normalize_to_host needs a pinned, contiguous float32 host tensor of raw's shape")
    work = _aligned(raw)
    _lib.check(lib.obia_b200_normalize_to(_p(work), ctypes.c_void_p(host.data_ptr()), H * W, C, _p(minmax_dev),
                                          int(max_ctas), _stream_ptr()), "normalize_to")


def normalize_to(raw, out, minmax_dev):
    """Normalised copy of `raw` into the CUDA tensor `out` (same shape, float32, contiguous, 16-byte
    aligned).  `raw` is left untouched."""
    lib = _lib.load()
    _require_cuda(raw, "raw", torch.float32)
    _require_cuda(out, "out", torch.float32)
    H, W, C = raw.shape
    if tuple(out.shape) != tuple(raw.shape) or out.data_ptr() % 16:
        raise ValueError("normalize_to needs an aligned output of raw's shape")
    work = _aligned(raw)
    _lib.check(lib.obia_b200_normalize_to(_p(work), _p(out), H * W, C, _p(minmax_dev), 0, _stream_ptr()),
               "normalize_to")


def normalize_inplace(raw, minmax_dev):
    """`img_data[:, :, i] = normalize_band(...)` for every band (segment_boundaries.py:31-33)."""
    lib = _lib.load()
    _require_cuda(raw, "raw", torch.float32)
    H, W, C = raw.shape
    work = _aligned(raw)
    _lib.check(lib.obia_b200_normalize_inplace(_p(work), H * W, C, _p(minmax_dev), _stream_ptr()),
               "normalize_inplace")
    if work is not raw:
        raw.copy_(work)


def mask_centroids_device(mask_dev, n_segments):
    """maskSLIC initial centres (skimage `_get_mask_centroids`) with the k-means sweeps on the GPU.

    Host part: the two RandomState(123) draws (cached per mask size) and the mean of the
    nearest-centroid offsets; device part: `obia_b200_mask_kmeans` (bit-identical to scipy's
    kmeans2 on pixel coordinates) and, beyond 64 centroids, `obia_b200_nearest_centroid`.
    Returns (yx float64 (n, 2) numpy, steps (3,), number of mask pixels).
    """
    lib = _lib.load()
    coord = torch.nonzero(mask_dev)                 # row-major, like np.nonzero
    n_coord = int(coord.shape[0])
    if n_coord == 0:
        # scikit-image fails on `image[mask].min()` of an empty selection
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    if n_segments <= 0:
        raise ValueError("n_segments must be positive")
    idx, idx_dense = slic_host.mask_sample_indices(n_coord, n_segments)
    dev = mask_dev.device
    if idx_dense is None:      # coord[idx_dense] is every mask pixel
        pts = coord.to(torch.int32).contiguous()
    else:
        pts = coord.index_select(0, torch.from_numpy(idx_dense).to(dev)).to(torch.int32).contiguous()
    cent = coord.index_select(0, torch.from_numpy(idx).to(dev)).to(torch.float64).contiguous()
    n = int(cent.shape[0])
    H, W = (int(v) for v in mask_dev.shape)
    ws = torch.empty((lib.obia_b200_mask_kmeans_workspace_bytes(n, H, W),), dtype=torch.uint8, device=dev)
    _lib.check(lib.obia_b200_mask_kmeans(_p(pts), int(pts.shape[0]), _p(cent), n, 5, H, W, _p(ws), _stream_ptr()),
               "mask_kmeans")
    if n <= 64:
        yx = cent.cpu().numpy()
        return yx, slic_host.steps_from_centroids(yx), n_coord
    # the nearest-other-centroid search runs on the device as well (the host version needs scipy's
    # n x n pdist matrix -- 0.4 ms for a 200-pixel tile -- or a k-d tree that breaks ties differently)
    closest = torch.empty((n,), dtype=torch.int32, device=dev)
    _lib.check(lib.obia_b200_nearest_centroid(_p(cent), n, H, W, _p(closest), _p(ws), _stream_ptr()),
               "nearest_centroid")
    # one read-back for both results
    packed = torch.cat([cent.reshape(-1), closest.to(torch.float64)]).cpu().numpy()
    yx = packed[:2 * n].reshape(n, 2).copy()
    return yx, slic_host.steps_from_centroids(yx, closest=packed[2 * n:].astype(np.int64)), n_coord


@dataclass
class SlicResult:
    labels: torch.Tensor            # (H, W) int32; masked pixels = -1 when a mask was given
    n_labels: int                   # number of kept segments after connectivity (or centres)
    start_label: int
    n_centres: int
    step: float
    step_yx: tuple
    minmax: torch.Tensor            # (C, 4) device
    minmax_host: np.ndarray
    centres: torch.Tensor           # (n, 2 + Cf) device, final means
    pre_connectivity: torch.Tensor | None = None
    features: torch.Tensor | None = None
    timings: dict = field(default_factory=dict)


def slic_labels(raw, segmentation_bands=None, *, n_segments=100, compactness=10.0, max_num_iter=10,
                sigma=0, spacing=None, convert2lab=None, enforce_connectivity=True,
                min_size_factor=0.5, max_size_factor=3, slic_zero=False, start_label=1, mask=None,
                channel_axis=-1, keep_intermediates=False, init_centroids=None, minmax=None, exact=False):
    """SLIC label raster of `raw[:, :, segmentation_bands]` with obia's wrapper semantics.

    Equivalent of segment_boundaries.py:31-57 up to the label raster: every band
    is min-max normalised (on the fly; `raw` itself is NOT modified here), the
    selected bands go through skimage's slic() semantics, masked pixels get -1.

    kwargs are skimage.segmentation.slic's (obia forwards **kwargs verbatim).

    `spacing=(sy, sx)` other than (1, 1) runs the exact kernel whatever `exact` says.
    `exact=False` (default) runs the tolerance-mode assignment kernel (one FMA per channel, see
    csrc/slic_fast.cu: north_star's ">= 99.5 % label agreement" bar); `exact=True` the kernel that
    reproduces `_slic_cython`'s float32 operation order bit for bit given the centres.
    """
    lib = _lib.load()
    _require_cuda(raw, "raw", torch.float32)
    if raw.dim() != 3:
        raise ValueError("raw must be (H, W, C)")
    raw = _aligned(raw)
    if channel_axis not in (-1, 2):
        raise NotImplementedError("only channel_axis=-1 (obia always passes H, W, C)")
    # skimage's `spacing` (voxel size per spatial axis): scales the spatial term of the distance and divides
    # sigma; grid, windows and connectivity stay in pixel units.  float32 like the image (slic: `dtype`).
    sp_y, sp_x = slic_host.parse_spacing(spacing)
    anisotropic = bool(sp_y != 1.0 or sp_x != 1.0)
    if start_label not in (0, 1):
        raise ValueError("start_label should be 0 or 1.")
    H, W, C = (int(s) for s in raw.shape)
    dev = raw.device
    if segmentation_bands is None:
        segmentation_bands = list(range(C))
    segmentation_bands = [int(b) for b in segmentation_bands]
    for band in segmentation_bands:
        if band >= C or band < 0:
            raise IndexError(f"Band index {band} out of range. Available bands indices: 0 to {C - 1}.")
    Cs = len(segmentation_bands)

    mask_dev = None
    if mask is not None:
        if isinstance(mask, torch.Tensor):
            mask_dev = (mask != 0).to(device=dev, dtype=torch.uint8).contiguous()
        else:
            mask_dev = torch.from_numpy(np.ascontiguousarray(np.asarray(mask) != 0).astype(np.uint8)).to(dev)
        if tuple(mask_dev.shape) != (H, W):
            raise ValueError("image and mask should have the same shape.")

    # ---- K1a: band ranges (needed on the host for validation, like numpy does) ----
    if minmax is not None and mask_dev is None:
        minmax, flags = minmax           # already computed by the caller (same raster, no mask)
    else:
        minmax, flags = band_minmax(raw, mask_dev)
    mm = minmax.cpu().numpy()
    fl = flags.cpu().numpy()
    f32 = np.float32
    with np.errstate(invalid="ignore", divide="ignore"):
        nmin, nmax = [], []
        for b in segmentation_bands:
            mn, mx, mmn, mmx = (f32(v) for v in mm[b])
            if fl[b] or not np.isfinite(mn) or not np.isfinite(mx) or mx == mn:
                # NaN/inf in the band, or a constant band (0/0 after obia's normalise)
                raise ValueError("unmasked NaN values in image are not supported")
            if mask_dev is not None and np.isnan(mmn):
                raise ValueError("zero-size array to reduction operation minimum which has no identity")
            d = f32(mx - mn)
            nmin.append(f32(f32(mmn - mn) / d))
            nmax.append(f32(f32(mmx - mn) / d))
        imin, imax = f32(min(nmin)), f32(max(nmax))

    multichannel = True
    to_lab = False
    if multichannel and (convert2lab or convert2lab is None):
        if Cs != 3 and convert2lab:
            raise ValueError("Lab colorspace conversion requires a RGB image.")
        to_lab = Cs == 3
    Cf = 3 if to_lab else Cs

    # ---- centres --------------------------------------------------------------
    n_mask = None      # number of mask pixels, when already known
    if init_centroids is not None or mask_dev is not None:
        if init_centroids is not None:
            yx, steps = init_centroids
        else:
            yx, steps, n_mask = mask_centroids_device(mask_dev, int(n_segments))
        n = int(yx.shape[0])
        centres_np = np.zeros((n, 2 + Cf), dtype=np.float32)
        centres_np[:, :2] = yx.astype(np.float32)
        centres = torch.from_numpy(centres_np).to(dev)
    else:
        # regular grid (skimage regular_grid, y-major): laid out on the device, the host only
        # derives the per-axis start / step (integers < 2^24: exact in float32)
        starts, isteps = slic_host.regular_grid_steps((1, H, W), n_segments)
        ys = torch.arange(starts[1], H, isteps[1] or 1, device=dev, dtype=torch.float32)
        xs = torch.arange(starts[2], W, isteps[2] or 1, device=dev, dtype=torch.float32)
        n = int(ys.numel() * xs.numel())
        centres = torch.zeros((n, 2 + Cf), dtype=torch.float32, device=dev)
        centres[:, 0] = ys.repeat_interleave(xs.numel())
        centres[:, 1] = xs.repeat(ys.numel())
        steps = [1.0 if s is None else float(s) for s in isteps]
    step = float(max(steps))
    step_y, step_x = slic_host.window_steps(H, W, n)
    if not step > 0:
        raise ValueError("degenerate SLIC initialisation (step == 0)")

    # ---- K1b: features ----------------------------------------------------------
    ratio = f32(1.0 / compactness)
    sig = f32(sigma) if np.ndim(sigma) == 0 else None
    if sig is None:
        sig_y, sig_x = (f32(s) for s in np.ravel(sigma)[-2:])
    else:
        sig_y = sig_x = sig
    sig_y, sig_x = f32(sig_y / sp_y), f32(sig_x / sp_x)     # slic: `sigma /= spacing`
    smooth = bool(sig_y > 0 or sig_x > 0)
    pitch = (W + 31) // 32 * 32
    feats = torch.empty((Cf, H, pitch), dtype=torch.float32, device=dev)
    bands_arr = _i32_array(segmentation_bands)
    bmin = np.ascontiguousarray(mm[:, 0], dtype=np.float32)
    bmax = np.ascontiguousarray(mm[:, 1], dtype=np.float32)
    _lib.check(lib.obia_b200_slic_features(
        _p(raw), H, W, C, bands_arr, Cs, bmin.ctypes.data_as(ctypes.c_void_p),
        bmax.ctypes.data_as(ctypes.c_void_p), float(imin), float(imax), int(to_lab),
        float(1.0 if smooth else ratio), _p(feats), pitch, _stream_ptr()), "slic_features")
    if smooth:
        wy, ry = slic_host.gaussian_taps(sig_y) if sig_y > 0 else (np.ones(1), 0)
        wx, rx = slic_host.gaussian_taps(sig_x) if sig_x > 0 else (np.ones(1), 0)
        smoothed = torch.empty_like(feats)
        if max(ry, rx) <= 63:
            # fused tiled kernel: in -> out, tmp unused (any distinct buffer satisfies the argument check)
            unused = torch.empty((4,), dtype=torch.float32, device=dev)
            _lib.check(lib.obia_b200_gaussian_planar(
                _p(feats), _p(unused), _p(smoothed), H, W, pitch, Cf,
                wy.ctypes.data_as(ctypes.c_void_p), ry, wx.ctypes.data_as(ctypes.c_void_p), rx, float(ratio),
                _stream_ptr()), "gaussian_planar")
            feats = smoothed
        else:
            _lib.check(lib.obia_b200_gaussian_planar(
                _p(feats), _p(smoothed), _p(feats), H, W, pitch, Cf, wy.ctypes.data_as(ctypes.c_void_p), ry,
                wx.ctypes.data_as(ctypes.c_void_p), rx, float(ratio), _stream_ptr()), "gaussian_planar")
            del smoothed

    # ---- K2: iterations -----------------------------------------------------------
    max_abs = float(ratio) * (256.0 if to_lab else 4.0)
    fix_scale = slic_host.fixed_point_scale(max_abs, H, W, step_y, step_x)
    ws_bytes = lib.obia_b200_slic_workspace_bytes(H, W, Cf, n, step_y, step_x)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    labels = torch.empty((H, W), dtype=torch.int32, device=dev)
    status = torch.zeros((4,), dtype=torch.int32, device=dev)

    def run(ignore_color):
        if anisotropic:
            # the spacing factors sit inside the squared spatial term: exact kernel only (see the header)
            _lib.check(lib.obia_b200_slic_iterate_spacing(
                _p(feats), _p(mask_dev), _p(centres), _p(labels), _p(ws), H, W, pitch, Cf, n, step,
                step_y, step_x, int(max_num_iter), int(start_label), int(ignore_color), int(bool(slic_zero)),
                fix_scale, float(sp_y), float(sp_x), _p(status), _stream_ptr()), "slic_iterate_spacing")
        else:
            iterate = lib.obia_b200_slic_iterate if exact else lib.obia_b200_slic_iterate_fast
            _lib.check(iterate(
                _p(feats), _p(mask_dev), _p(centres), _p(labels), _p(ws), H, W, pitch, Cf, n, step,
                step_y, step_x, int(max_num_iter), int(start_label), int(ignore_color), int(bool(slic_zero)),
                fix_scale, _p(status), _stream_ptr()), "slic_iterate")
        if int(status[0].item()) != 0:
            raise _lib.CandidateOverflowError(
                "slic_iterate: a 32 x 64 pixel tile collected more than 1024 candidate centres -- the centre "
                "density is beyond what the kernels stage per tile (n_segments above about one centre per 2 "
                "pixels, or a mask that packs all centres into a small part of the raster)")

    if mask_dev is not None:
        run(True)   # maskSLIC step 2: spatial-only k-means moves the centres first
    run(False)
    del ws

    pre_cc = labels.clone() if keep_intermediates else None
    n_labels = n
    if enforce_connectivity:
        if mask_dev is not None:
            segment_size = float(n_mask if n_mask is not None else mask_dev.sum().item()) / n
        else:
            segment_size = float(H * W) / n
        min_size = int(min_size_factor * segment_size)
        max_size = int(max_size_factor * segment_size)
        if max_size < 1:
            raise ValueError("max_size_factor too small: connectivity needs max_size >= 1")
        cc_ws = scratch(dev, lib.obia_b200_connectivity_workspace_bytes(H, W), "connectivity")
        out = torch.empty_like(labels)
        nl = ctypes.c_int64(0)
        _lib.check(lib.obia_b200_enforce_connectivity(
            _p(labels), _p(out), _p(cc_ws), H, W, min_size, max_size, int(start_label),
            ctypes.byref(nl), _stream_ptr()), "enforce_connectivity")
        labels = out
        n_labels = int(nl.value)
        del cc_ws

    if mask_dev is not None:
        labels.masked_fill_(mask_dev == 0, -1)   # segment_boundaries.py:55-57

    return SlicResult(labels=labels, n_labels=n_labels, start_label=int(start_label), n_centres=n,
                      step=step, step_yx=(step_y, step_x), minmax=minmax, minmax_host=mm, centres=centres,
                      pre_connectivity=pre_cc, features=feats if keep_intermediates else None)


def quickshift_labels(raw, segmentation_bands=None, *, ratio=1.0, kernel_size=5, max_dist=10, return_tree=False,
                      sigma=0, convert2lab=True, rng=42, random_seed=None, channel_axis=-1):
    """`skimage.segmentation.quickshift` on `raw[:, :, segmentation_bands]` with obia's wrapper semantics
    (segment_boundaries.py:31-49: every band min-max normalised first).  kwargs are scikit-image's
    (`random_seed` is the pre-0.21 name of `rng`).  Returns (labels (H, W) int32 on the device, 0..n-1; n)."""
    lib = _lib.load()
    _require_cuda(raw, "raw", torch.float32)
    raw = _aligned(raw)
    if channel_axis not in (-1, 2):
        raise NotImplementedError("only channel_axis=-1 (obia always passes H, W, C)")
    if return_tree:
        raise NotImplementedError("return_tree=True returns a tuple obia's create_segments cannot use")
    if random_seed is not None:
        rng = random_seed
    H, W, C = (int(v) for v in raw.shape)
    dev = raw.device
    bands = list(range(C)) if segmentation_bands is None else [int(b) for b in segmentation_bands]
    for band in bands:
        if band >= C or band < 0:
            raise IndexError(f"Band index {band} out of range. Available bands indices: 0 to {C - 1}.")
    if convert2lab and len(bands) != 3:
        raise ValueError("Only RGB images can be converted to Lab space.")
    if kernel_size < 1:
        raise ValueError("`kernel_size` should be >= 1.")
    minmax, flags = band_minmax(raw)
    mm = minmax.cpu().numpy()
    Cf = len(bands)
    pitch = (W + 31) // 32 * 32
    feats = torch.empty((Cf, H, pitch), dtype=torch.float32, device=dev)
    bmin = np.ascontiguousarray(mm[:, 0], dtype=np.float32)
    bmax = np.ascontiguousarray(mm[:, 1], dtype=np.float32)
    smooth = float(sigma) > 0
    f32 = np.float32
    _lib.check(lib.obia_b200_slic_features(
        _p(raw), H, W, C, _i32_array(bands), Cf, bmin.ctypes.data_as(ctypes.c_void_p),
        bmax.ctypes.data_as(ctypes.c_void_p), 0.0, 1.0, int(bool(convert2lab)), 1.0 if smooth else float(f32(ratio)),
        _p(feats), pitch, _stream_ptr()), "slic_features")
    if smooth:
        w, r = slic_host.gaussian_taps(f32(sigma))
        out = torch.empty_like(feats)
        tmp = torch.empty_like(feats) if r > 63 else torch.empty((4,), dtype=torch.float32, device=dev)
        _lib.check(lib.obia_b200_gaussian_planar(
            _p(feats), _p(tmp), _p(out), H, W, pitch, Cf, w.ctypes.data_as(ctypes.c_void_p), r,
            w.ctypes.data_as(ctypes.c_void_p), r, float(f32(ratio)), _stream_ptr()), "gaussian_planar")
        feats = out
    # tie-breaking noise of `_quickshift_cython`: numpy's Generator stream, drawn on the host
    noise = torch.from_numpy(np.random.default_rng(rng).normal(scale=0.00001, size=(H, W))).to(dev)
    labels = torch.empty((H, W), dtype=torch.int32, device=dev)
    ws = torch.empty((lib.obia_b200_quickshift_workspace_bytes(H, W),), dtype=torch.uint8, device=dev)
    n = ctypes.c_int64(0)
    _lib.check(lib.obia_b200_quickshift(_p(feats), _p(noise), _p(labels), _p(ws), H, W, pitch, Cf, float(kernel_size),
                                        float(max_dist), ctypes.byref(n), _stream_ptr()), "quickshift")
    return labels, int(n.value)


def enforce_connectivity(labels, min_size, max_size, start_label=1):
    """K3 on its own: (labels_out int32 (H, W), number of kept segments)."""
    lib = _lib.load()
    labels = _aligned(_require_cuda(labels, "labels", torch.int32))
    H, W = (int(s) for s in labels.shape)
    ws = scratch(labels.device, lib.obia_b200_connectivity_workspace_bytes(H, W), "connectivity")
    out = torch.empty_like(labels)
    nl = ctypes.c_int64(0)
    _lib.check(lib.obia_b200_enforce_connectivity(_p(labels), _p(out), _p(ws), H, W, int(min_size),
                                                  int(max_size), int(start_label), ctypes.byref(nl),
                                                  _stream_ptr()), "enforce_connectivity")
    return out, int(nl.value)


# ------------------------------------------------------------------ K4 ------
def zonal_stats(labels, raw, bands=None, max_label=None, resolution=1e-6):
    """Per-label, per-band statistics table, float64 (max_label+1, Cz, 8) on the device.

    Fields: STAT_FIELDS.  Labels < 0 (obia's -1 = masked) are ignored.
    """
    lib = _lib.load()
    labels = _aligned(_require_cuda(labels, "labels", torch.int32))
    raw = _aligned(_require_cuda(raw, "raw", torch.float32))
    H, W, C = (int(s) for s in raw.shape)
    if tuple(labels.shape) != (H, W):
        raise ValueError("labels and raster shapes differ")
    if bands is None:
        bands = list(range(C))
    bands = [int(b) for b in bands]
    if max_label is None:
        max_label = int(labels.max().item())
    max_label = max(int(max_label), 0)
    Cz = len(bands)
    stats = torch.empty((max_label + 1, Cz, 8), dtype=torch.float64, device=raw.device)
    ws = torch.empty((lib.obia_b200_zonal_workspace_bytes(max_label, 8),), dtype=torch.uint8,
                     device=raw.device)
    for c0 in range(0, Cz, 64):   # kernel-parameter table holds 64 bands per call
        sub = bands[c0:c0 + 64]
        if Cz <= 64:
            view = stats
        else:
            view = torch.empty((max_label + 1, len(sub), 8), dtype=torch.float64, device=raw.device)
        _lib.check(lib.obia_b200_zonal_stats(_p(labels), _p(raw), H, W, C, _i32_array(sub), len(sub),
                                             max_label, float(resolution), _p(view), _p(ws),
                                             _stream_ptr()), "zonal_stats")
        if Cz > 64:
            stats[:, c0:c0 + len(sub)] = view
    return stats


def texture_stats(labels, raw, bands=None, max_label=None, quantise_f64=False):
    """Per-label, per-band GLCM texture features, float64 (max_label+1, n_bands, 6) on the device.

    Fields: TEXTURE_FIELDS (mean over the four angles at distance 2, segment_statistics.py:261-296).
    `quantise_f64`: the reference scales the crop to uint8 in the dtype of its masked crop --
    float32 for float32 rasters (False), float64 for integer rasters (True).
    """
    lib = _lib.load()
    labels = _aligned(_require_cuda(labels, "labels", torch.int32))
    raw = _aligned(_require_cuda(raw, "raw", torch.float32))
    H, W, C = (int(s) for s in raw.shape)
    if tuple(labels.shape) != (H, W):
        raise ValueError("labels and raster shapes differ")
    if bands is None:
        bands = list(range(C))
    bands = [int(b) for b in bands]
    if max_label is None:
        max_label = int(labels.max().item())
    max_label = max(int(max_label), 0)
    nb = len(bands)
    feats = torch.empty((max_label + 1, nb, 6), dtype=torch.float64, device=raw.device)
    ws = torch.empty((lib.obia_b200_texture_workspace_bytes(max_label),), dtype=torch.uint8, device=raw.device)
    for c0 in range(0, nb, 64):   # kernel-parameter table holds 64 bands per call
        sub = bands[c0:c0 + 64]
        view = feats if nb <= 64 else torch.empty((max_label + 1, len(sub), 6), dtype=torch.float64,
                                                  device=raw.device)
        _lib.check(lib.obia_b200_texture_stats(_p(labels), _p(raw), H, W, C, _i32_array(sub), len(sub),
                                               max_label, int(bool(quantise_f64)), _p(view), _p(ws),
                                               _stream_ptr()), "texture_stats")
        if nb > 64:
            feats[:, c0:c0 + len(sub)] = view
    return feats

"""Raster container of the hot path: mirror of obia/handlers/geotif.py.

`Image` keeps the reference's constructor and attributes
(/root/reference/obia/handlers/geotif.py:8-44): `img_data` is an (H, W, C)
float32 array (numpy, or a CUDA torch tensor when the raster is already
resident on the GPU).  File I/O (`open_geotiff`, geotif.py:78-106) stays on the
host and needs rasterio, which is optional.
"""
from __future__ import annotations

import numpy as np


class Image:
    img_data = None
    crs = None
    transform = None
    affine_transformation = None
    rasterio_obj = None

    def __init__(self, img_data, crs, affine_transformation, transform, rasterio_obj):
        self.img_data = img_data
        self.crs = crs
        self.affine_transformation = affine_transformation
        self.transform = transform
        self.rasterio_obj = rasterio_obj

    # -- B200 additions (not part of the reference contract) ------------------
    def stats_in_float64(self):
        """True when the reference would compute the per-segment features in float64.

        `create_objects` re-reads the FILE in its native dtype (obia/utils/utils.py:45-48) and
        `np.where(mask, crop, nan)` (:64) promotes integer rasters to float64, float32 rasters
        stay float32.  An in-memory `img_data` of integer dtype counts as an integer raster.
        """
        obj = self.rasterio_obj
        dtypes = getattr(obj, "dtypes", None)
        if dtypes:
            return np.dtype(dtypes[0]) != np.float32
        dt = getattr(self.img_data, "dtype", None)
        if dt is None:
            return False
        if not isinstance(dt, np.dtype):   # torch dtype
            import torch
            return dt not in (torch.float32, torch.float16, torch.bfloat16)
        return dt != np.float32

    def device_raw(self, device=None, refresh=False):
        """Raw (un-normalised) raster as a CUDA float32 (H, W, C) tensor, uploaded once.

        The reference reads raw values back from the file for the statistics
        (obia/utils/utils.py:45-48) after `create_segments` has normalised
        `img_data` in place; here the raw values are kept in HBM instead.
        """
        import torch

        cached = getattr(self, "_obia_b200_raw", None)
        key = self._raw_key()
        if cached is not None and not refresh and getattr(self, "_obia_b200_raw_key", None) == key:
            return cached
        data = self.img_data
        if isinstance(data, torch.Tensor):
            if not data.is_cuda:
                data = data.to(device or "cuda")
            raw = data.to(dtype=torch.float32).contiguous()
            if raw.data_ptr() == self.img_data.data_ptr():
                raw = raw.clone()   # img_data may be normalised in place later
        else:
            arr = np.ascontiguousarray(np.asarray(data), dtype=np.float32)
            raw = torch.from_numpy(arr).to(device or "cuda", non_blocking=True)
        if raw.dim() != 3:
            raise ValueError("img_data must be (H, W, C)")
        self._obia_b200_raw = raw
        self._obia_b200_raw_key = key
        return raw

    def _raw_key(self):
        """Identity of `img_data` the cached raw raster belongs to: object id, shape, dtype and, for
        torch tensors, the in-place version counter.  Assigning a new array to `img_data` (a crop,
        band maths, another tile) or editing a tensor in place invalidates the cache; in-place edits
        of a numpy array cannot be detected (pass `refresh=True`).  "Raw" = the values at first use:
        the reference normalises `img_data` in place and re-reads the file for the statistics."""
        data = self.img_data
        ver = getattr(data, "_version", None)
        return (id(data), tuple(getattr(data, "shape", ())), str(getattr(data, "dtype", None)), ver)

    def _note_mutation(self):
        """Called after obia_b200 itself has normalised `img_data` in place (the reference's side
        effect): the cached raw values stay the ones to use."""
        if getattr(self, "_obia_b200_raw", None) is not None:
            self._obia_b200_raw_key = self._raw_key()


def open_geotiff(image_path, bands=None):
    """geotif.py:78-106.  Host-side file read; needs rasterio."""
    try:
        import rasterio
    except ImportError as e:  # pragma: no cover - rasterio is not in this image
        raise ImportError("open_geotiff needs rasterio (host-side file I/O is outside the GPU path)") from e
    rasterio_obj = rasterio.open(image_path)
    crs = rasterio_obj.crs
    transform = rasterio_obj.transform
    affine_transformation = [transform.a, transform.b, transform.d, transform.e, transform.c, transform.f]
    if bands is None:
        bands = list(range(1, rasterio_obj.count + 1))
    data = np.empty((rasterio_obj.height, rasterio_obj.width, len(bands)), dtype=np.float32)
    for i, b in enumerate(bands):
        data[:, :, i] = rasterio_obj.read(b)
    return Image(data, crs, affine_transformation, transform, rasterio_obj)

"""CPU oracle for the zonal-statistics half of the hot path -- TEST INFRASTRUCTURE ONLY.

Restates `calculate_spectral_stats`
(/root/reference/obia/segmentation/segment_statistics.py:113-176) applied to
each segment the way `create_objects` does (:475-491): the pixels of the
segment (crop_image_to_bbox + mask_image_with_polygon, obia/utils/utils.py:37-67,
i.e. exactly the segment's pixels, everything else NaN) reduced per band with
`np.mean`, `np.var`, `np.min`, `np.max`, `scipy.stats.skew`,
`scipy.stats.kurtosis` -- the very numpy/scipy calls the reference makes (both
are installed here, so this half of the oracle is the reference's own
arithmetic, not a restatement of it).

The reference computes in the dtype `np.where(mask, crop, nan)` yields:
float32 for float32 rasters, float64 for integer rasters.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage
from scipy.stats import kurtosis, skew

STAT_NAMES = ("mean", "variance", "min", "max", "skewness", "kurtosis")


def calculate_spectral_stats(image_chw, statistics_bands):
    """segment_statistics.py:113-176 verbatim in behaviour (all calc_* flags on)."""
    stats = {}
    for b in statistics_bands:
        band = image_chw[b, :, :]
        flat = band[~np.isnan(band)]
        p = f"b{b}"
        if flat.size == 0:
            for n in STAT_NAMES:
                stats[f"{p}_{n}"] = np.nan
        else:
            stats[f"{p}_mean"] = np.mean(flat)
            stats[f"{p}_variance"] = np.var(flat)
            stats[f"{p}_min"] = np.min(flat)
            stats[f"{p}_max"] = np.max(flat)
            stats[f"{p}_skewness"] = skew(flat)
            stats[f"{p}_kurtosis"] = kurtosis(flat)
    return stats


def zonal_stats(labels_hw, raw_hwc, bands, label_values, compute_dtype=None):
    """Reference statistics for every label in `label_values`.

    Returns float64 array (len(label_values), len(bands), 6) in STAT_NAMES order
    plus the pixel counts.  `compute_dtype`: dtype of the masked crop the
    reference would reduce (float32 for float32 rasters, float64 for integer ones).
    """
    labels_hw = np.asarray(labels_hw)
    raw_hwc = np.asarray(raw_hwc)
    if compute_dtype is None:
        compute_dtype = np.float32 if raw_hwc.dtype == np.float32 else np.float64
    out = np.full((len(label_values), len(bands), 6), np.nan, dtype=np.float64)
    counts = np.zeros(len(label_values), dtype=np.int64)
    slices = ndimage.find_objects(np.where(labels_hw >= 0, labels_hw + 1, 0).astype(np.int64))
    for i, lv in enumerate(label_values):
        sl = slices[lv] if lv < len(slices) else None
        if sl is None:
            continue
        crop = np.moveaxis(raw_hwc[sl], -1, 0)                 # (C, h, w) like rasterio.read
        m = labels_hw[sl] == lv
        counts[i] = int(m.sum())
        masked = np.where(m[np.newaxis], crop, np.nan).astype(compute_dtype, copy=False)
        with np.errstate(all="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                st = calculate_spectral_stats(masked, bands)
        for j, b in enumerate(bands):
            for k, n in enumerate(STAT_NAMES):
                out[i, j, k] = st[f"b{b}_{n}"]
    return out, counts

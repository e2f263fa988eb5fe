"""CPU oracle for the GLCM texture columns -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.

Restates `calculate_textural_stats`
(/root/reference/obia/segmentation/segment_statistics.py:179-298) as
`create_objects` applies it to every segment (:496-508): the masked bounding-box
crop of the segment (obia/utils/utils.py:37-67: pixels outside the segment are
NaN) -> NaN replaced by 0 (:246-247) -> min/max scaling of the WHOLE crop
(zeros included) to uint8 (:251-258) -> `graycomatrix(distances=[2],
angles=[0, pi/4, pi/2, 3pi/4], levels=256, symmetric=True, normed=True)`
(:261-268) -> mean over the four angles of `graycoprops` contrast /
dissimilarity / homogeneity / ASM / energy / correlation (:285-296).

scikit-image is not installable here (no network, not in the wheelhouse), so
`graycomatrix` / `graycoprops` are restated from the published algorithm
[UPSTREAM skimage/feature/texture.py, skimage/feature/_texture.pyx::_glcm_loop,
scikit-image >= 0.23.2 per the reference's pyproject.toml:23, unpinned]:
offset = (round(sin(a) * d), round(cos(a) * d)); every pixel whose offset
partner lies inside the image contributes one count at [i, j]; symmetric adds
the transpose; normed divides each (distance, angle) slice by its sum (0 -> 1);
graycoprops normalises again, then takes weighted sums; correlation is 1 where
either standard deviation is < 1e-15.  The restatement reproduces the worked
example of the graycomatrix docstring and the property values of skimage's own
test-suite as recalled (tests/test_texture_oracle.py) -- recalled, not
re-verified against a live scikit-image: hence "parity unpinned".

ONE documented deviation from the reference, SURVEY.md 8a a11 / defect 6: the
reference indexes the band-FIRST crop (C, h, w) with `image[:, :, band_index]`
(:214), i.e. it takes column `band_index` of every band and raises IndexError
whenever the crop is narrower than the band index.  This oracle (and the CUDA
path) index the band: `image[band_index, :, :]`.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module.
"""
from __future__ import annotations

import numpy as np

TEXTURE_NAMES = ("contrast", "dissimilarity", "homogeneity", "ASM", "energy", "correlation")
DISTANCES = (2,)
ANGLES = (0.0, np.pi / 4, np.pi / 2, 3 * np.pi / 4)


def _c_round(v):
    """C `round`: half away from zero (the Cython loop uses libc round)."""
    return int(np.floor(abs(v) + 0.5) * (1 if v >= 0 else -1))


def graycomatrix(image, distances, angles, levels=256, symmetric=False, normed=False):
    """[UPSTREAM skimage.feature.graycomatrix] -> (levels, levels, n_dist, n_angle)."""
    image = np.asarray(image)
    if image.ndim != 2:
        raise ValueError("The parameter `image` must be a 2-dimensional array")
    rows, cols = image.shape
    out = np.zeros((levels, levels, len(distances), len(angles)), dtype=np.uint32)
    img = image.astype(np.int64)
    for a_idx, angle in enumerate(angles):
        for d_idx, distance in enumerate(distances):
            dr = _c_round(np.sin(angle) * distance)
            dc = _c_round(np.cos(angle) * distance)
            r0, r1 = max(0, -dr), min(rows, rows - dr)
            c0, c1 = max(0, -dc), min(cols, cols - dc)
            if r1 <= r0 or c1 <= c0:
                continue
            i = img[r0:r1, c0:c1].ravel()
            j = img[r0 + dr:r1 + dr, c0 + dc:c1 + dc].ravel()
            ok = (i >= 0) & (i < levels) & (j >= 0) & (j < levels)
            np.add.at(out[:, :, d_idx, a_idx], (i[ok], j[ok]), 1)
    if symmetric:
        out = out + np.transpose(out, (1, 0, 2, 3))
    if normed:
        out = out.astype(np.float64)
        sums = np.sum(out, axis=(0, 1), keepdims=True)
        sums[sums == 0] = 1
        out /= sums
    return out


def graycoprops(P, prop="contrast"):
    """[UPSTREAM skimage.feature.graycoprops] -> (n_dist, n_angle) float64."""
    num_level, num_level2, num_dist, num_angle = P.shape
    P = P.astype(np.float64)
    sums = np.sum(P, axis=(0, 1), keepdims=True)
    sums[sums == 0] = 1
    P = P / sums
    I, J = np.ogrid[0:num_level, 0:num_level]
    if prop == "contrast":
        weights = (I - J) ** 2
    elif prop == "dissimilarity":
        weights = np.abs(I - J)
    elif prop == "homogeneity":
        weights = 1.0 / (1.0 + (I - J) ** 2)
    elif prop in ("ASM", "energy", "correlation"):
        weights = None
    else:
        raise ValueError(f"{prop} is an invalid property")
    if prop == "energy":
        return np.sqrt(np.sum(P ** 2, axis=(0, 1)))
    if prop == "ASM":
        return np.sum(P ** 2, axis=(0, 1))
    if prop == "correlation":
        results = np.zeros((num_dist, num_angle), dtype=np.float64)
        Ic = np.arange(num_level).reshape((num_level, 1, 1, 1))
        Jc = np.arange(num_level).reshape((1, num_level, 1, 1))
        diff_i = Ic - np.sum(Ic * P, axis=(0, 1))
        diff_j = Jc - np.sum(Jc * P, axis=(0, 1))
        std_i = np.sqrt(np.sum(P * diff_i ** 2, axis=(0, 1)))
        std_j = np.sqrt(np.sum(P * diff_j ** 2, axis=(0, 1)))
        cov = np.sum(P * (diff_i * diff_j), axis=(0, 1))
        mask_0 = std_i < 1e-15
        mask_0[std_j < 1e-15] = True
        results[mask_0] = 1
        mask_1 = ~mask_0
        results[mask_1] = cov[mask_1] / (std_i[mask_1] * std_j[mask_1])
        return results
    weights = weights.reshape((num_level, num_level, 1, 1))
    return np.sum(P * weights, axis=(0, 1))


def quantise_band(band_clean):
    """segment_statistics.py:249-258 for the float crops `mask_image_with_polygon` yields."""
    if np.issubdtype(band_clean.dtype, np.integer):
        return band_clean.astype(np.uint8)
    band_min, band_max = np.min(band_clean), np.max(band_clean)
    if band_max == band_min:
        return np.zeros(band_clean.shape, dtype=np.uint8)
    return ((band_clean - band_min) / (band_max - band_min) * 255).astype(np.uint8)


def calculate_textural_stats(image_chw, textural_bands):
    """segment_statistics.py:179-298 with the band axis fixed (see module docstring)."""
    stats = {}
    for b in textural_bands:
        band_data = image_chw[b, :, :]
        valid = ~np.isnan(band_data)
        p = f"b{b}"
        if not np.any(valid):
            for n in TEXTURE_NAMES:
                stats[f"{p}_{n}"] = np.nan
            continue
        band_clean = band_data.copy()
        band_clean[~valid] = 0
        glcm = graycomatrix(quantise_band(band_clean), distances=list(DISTANCES), angles=list(ANGLES),
                            levels=256, symmetric=True, normed=True)
        for n in TEXTURE_NAMES:
            stats[f"{p}_{n}"] = np.mean(graycoprops(glcm, n))
    return stats


def textural_stats(labels_hw, raw_hwc, bands, label_values, compute_dtype=None):
    """Texture features of every label in `label_values`: float64 (len(label_values), len(bands), 6)
    in TEXTURE_NAMES order.  `compute_dtype` as in stats_oracle.zonal_stats: the dtype of the masked
    crop (float32 for float32 rasters, float64 for integer rasters)."""
    labels_hw = np.asarray(labels_hw)
    raw_hwc = np.asarray(raw_hwc)
    if compute_dtype is None:
        compute_dtype = np.float32 if raw_hwc.dtype == np.float32 else np.float64
    out = np.full((len(label_values), len(bands), 6), np.nan, dtype=np.float64)
    for n, lab in enumerate(label_values):
        ys, xs = np.nonzero(labels_hw == lab)
        if ys.size == 0:
            continue
        y0, y1, x0, x1 = ys.min(), ys.max() + 1, xs.min(), xs.max() + 1
        crop = np.moveaxis(raw_hwc[y0:y1, x0:x1, :], 2, 0).astype(compute_dtype)      # (C, h, w)
        inside = (labels_hw[y0:y1, x0:x1] == lab)[None, :, :]
        masked = np.where(inside, crop, np.nan).astype(compute_dtype)
        st = calculate_textural_stats(masked, bands)
        for k, b in enumerate(bands):
            out[n, k] = [st[f"b{b}_{name}"] for name in TEXTURE_NAMES]
    return out

"""Drop-in for obia/segmentation/segment_statistics.py (`create_objects`).

Same signature, guards and column contract as the reference
(/root/reference/obia/segmentation/segment_statistics.py:392-511, columns from
`_create_empty_stats_columns` :12-110): `segment_id`, then per spectral band
`b{i}_mean, _variance, _min, _max, _skewness, _kurtosis`, then per textural band
`b{i}_contrast, _dissimilarity, _homogeneity, _ASM, _energy, _correlation`, then
`pai, fhd, ch, mean_intensity, variance_intensity`, then `geometry`.

The per-segment loop (crop -> polygon mask -> numpy/scipy statistics, :475-508)
is replaced by ONE fused pass of the CUDA zonal-statistics kernel over the label
raster.  Statistics are taken over the RAW raster values, like the reference
(which re-reads the file, obia/utils/utils.py:45-48).

The GLCM texture columns (`calculate_textural_stats`, :179-298) come from the CUDA
texture kernel (one CTA per segment, see csrc/texture.cu).  ONE documented
deviation: the reference indexes its band-first crop with `[:, :, band]` (:214) --
an axis slip that takes column `band` of every band and raises IndexError for
crops narrower than the band index; here the band is indexed.
The point-cloud columns are NaN in the reference too.
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def _create_empty_stats_columns(spectral_bands, textural_bands, calc_mean, calc_variance, calc_min, calc_max,
                                calc_skewness, calc_kurtosis,
                                calc_contrast, calc_dissimilarity, calc_homogeneity, calc_ASM, calc_energy,
                                calc_correlation,
                                calc_pai, calc_fhd, calc_ch, calc_mean_intensity, calc_variance_intensity):
    """Column list in the reference's order (segment_statistics.py:64-110)."""
    columns = ['segment_id']
    spectral = [("mean", calc_mean), ("variance", calc_variance), ("min", calc_min), ("max", calc_max),
                ("skewness", calc_skewness), ("kurtosis", calc_kurtosis)]
    textural = [("contrast", calc_contrast), ("dissimilarity", calc_dissimilarity),
                ("homogeneity", calc_homogeneity), ("ASM", calc_ASM), ("energy", calc_energy),
                ("correlation", calc_correlation)]
    for b in spectral_bands:
        columns += [f"b{b}_{name}" for name, on in spectral if on]
    for b in textural_bands:
        columns += [f"b{b}_{name}" for name, on in textural if on]
    cloud = [("pai", calc_pai), ("fhd", calc_fhd), ("ch", calc_ch), ("mean_intensity", calc_mean_intensity),
             ("variance_intensity", calc_variance_intensity)]
    columns += [name for name, on in cloud if on]
    columns.append('geometry')
    return columns


def _labels_of(segments, image, raw):
    """(label raster CUDA int32, label value per row) of a segments table.

    Tables made by obia_b200 `create_segments` carry their label raster; the label of a row follows its
    `segment_id` (1..N), so filtered / reordered rows keep their own segment.  Any other table with
    `geometry` + `segment_id` columns (a GeoDataFrame read from a GeoPackage, polygons edited by the
    user: what the reference accepts, segment_statistics.py:475-484) is burnt into a label raster with
    the pixel-centre rule of `mask_image_with_polygon` (utils.py:53-67) by the CUDA scanline kernel."""
    raster = getattr(segments, "label_raster", None)
    all_labels = getattr(segments, "segment_labels", None)
    if raster is not None and all_labels is not None:
        all_labels = np.asarray(all_labels, dtype=np.int64)
        ids = np.asarray(segments["segment_id"], dtype=np.int64)
        if len(ids) == len(all_labels) and np.array_equal(ids, np.arange(1, len(ids) + 1)):
            return raster, all_labels
        if len(ids) and (ids.min() < 1 or ids.max() > len(all_labels)):
            raise ValueError("segment_id values do not belong to the label raster this table carries")
        return raster, all_labels[ids - 1]
    if "geometry" not in segments or "segment_id" not in segments:
        raise TypeError("create_objects needs a table with 'geometry' and 'segment_id' columns")
    from ..utils.rasterize import rasterize_polygons
    geoms = list(segments["geometry"])
    if any(g is None for g in geoms):
        raise ValueError("create_objects needs polygon geometries (or the table returned by create_segments)")
    H, W = int(raw.shape[0]), int(raw.shape[1])
    raster = rasterize_polygons(geoms, H, W, getattr(image, "affine_transformation", None), device=raw.device)
    return raster, np.arange(1, len(geoms) + 1, dtype=np.int64)


def create_objects(
        segments, image, ept=None, ept_srs=None, spectral_bands=None, textural_bands=None, voxel_resolution=None,
        calculate_spectral=True, calculate_textural=True, calculate_structural=False, calculate_radiometric=False,
        calc_mean=True, calc_variance=True, calc_min=True, calc_max=True, calc_skewness=True, calc_kurtosis=True,
        calc_contrast=True, calc_dissimilarity=True, calc_homogeneity=True, calc_ASM=True, calc_energy=True,
        calc_correlation=True,
        calc_pai=True, calc_fhd=True, calc_ch=True, calc_mean_intensity=True, calc_variance_intensity=True
):
    """Per-segment feature table (see module docstring for the column contract)."""
    from .. import pipeline

    if not (calculate_spectral or calculate_textural or calculate_structural or calculate_radiometric):
        raise ValueError(
            "At least one of 'calculate_spectral', 'calculate_textural', 'calculate_structural', or 'calculate_radiometric' must be True."
        )
    if ept is not None or calculate_structural or calculate_radiometric:
        raise NotImplementedError(
            "Point-cloud workflows are temporarily disabled. "
            "Use spectral/textural statistics only for now."
        )

    raw = image.device_raw()
    n_bands = int(raw.shape[2])
    if spectral_bands is None:
        spectral_bands = list(range(n_bands))
    if textural_bands is None:
        textural_bands = list(range(n_bands))

    columns = _create_empty_stats_columns(
        spectral_bands, textural_bands,
        calc_mean, calc_variance, calc_min, calc_max, calc_skewness, calc_kurtosis,
        calc_contrast, calc_dissimilarity, calc_homogeneity, calc_ASM, calc_energy, calc_correlation,
        calc_pai, calc_fhd, calc_ch, calc_mean_intensity, calc_variance_intensity
    )

    raster, row_labels = _labels_of(segments, image, raw)
    n_rows = len(row_labels)
    float_cols = columns[1:-1]
    # one (n_columns, n_rows) float64 block, column-major for pandas (zero-copy): the spectral
    # statistics, then the texture features, in the reference's order; point-cloud columns stay NaN
    import torch
    # (page-locked: the device-to-host copy of the table runs at PCIe speed instead of through the
    #  pageable staging path -- 80 MB for c2: 2 ms instead of 20; torch caches the pinned allocation)
    block_t = torch.empty((len(float_cols), n_rows), dtype=torch.float64, pin_memory=raw.is_cuda)
    block = block_t.numpy()
    # the reference computes in the dtype of its masked crop: float32 for float32 rasters,
    # float64 for integer rasters (np.where(mask, crop, nan), utils.py:64)
    in_f64 = bool(getattr(image, "stats_in_float64", lambda: False)())
    want = [("mean", calc_mean), ("variance", calc_variance), ("min", calc_min), ("max", calc_max),
            ("skewness", calc_skewness), ("kurtosis", calc_kurtosis)]
    tex_want = [("contrast", calc_contrast), ("dissimilarity", calc_dissimilarity),
                ("homogeneity", calc_homogeneity), ("ASM", calc_ASM), ("energy", calc_energy),
                ("correlation", calc_correlation)]
    fields = [pipeline.STAT_FIELDS.index(name) for name, on in want if on]
    tex_fields = [pipeline.TEXTURE_FIELDS.index(name) for name, on in tex_want if on]
    n_spec = len(spectral_bands) * len(fields)          # spectral columns come first (:64-89) ...
    n_tex = len(textural_bands) * len(tex_fields)       # ... then the texture columns (:90-101)
    for b in list(spectral_bands) + (list(textural_bands) if calculate_textural else []):
        if b < 0 or b >= n_bands:
            # the reference indexes the (C, h, w) crop with the band number (:144)
            raise IndexError(f"index {b} is out of bounds for axis 0 with size {n_bands}")
    filled = np.zeros(len(float_cols), dtype=bool)
    if n_rows > 0:
        max_label = int(row_labels.max())
        rows = None
        # the reference computes the spectral statistics whatever `calculate_spectral` says (:485-491)
        if n_spec > 0:
            # scipy's "nearly constant" NaN rule uses the compute dtype's resolution
            stats = pipeline.zonal_stats(raster, raw, spectral_bands, max_label=max_label,
                                         resolution=1e-15 if in_f64 else 1e-6)
            rows = torch.from_numpy(row_labels).to(stats.device)
            sel = stats.index_select(0, rows)[:, :, fields]                  # (rows, Cz, nstat)
            sel = sel.permute(1, 2, 0).reshape(n_spec, n_rows).contiguous()
            assert float_cols[:n_spec] == [f"b{b}_{name}" for b in spectral_bands for name, on in want if on]
            block_t[:n_spec].copy_(sel, non_blocking=True)
            filled[:n_spec] = True
        if calculate_textural and n_tex > 0:                                 # (:494-506)
            feats = pipeline.texture_stats(raster, raw, textural_bands, max_label=max_label,
                                           quantise_f64=in_f64)
            if rows is None:
                rows = torch.from_numpy(row_labels).to(feats.device)
            sel = feats.index_select(0, rows)[:, :, tex_fields]
            sel = sel.permute(1, 2, 0).reshape(n_tex, n_rows).contiguous()
            assert float_cols[n_spec:n_spec + n_tex] == [f"b{b}_{name}" for b in textural_bands
                                                         for name, on in tex_want if on]
            block_t[n_spec:n_spec + n_tex].copy_(sel, non_blocking=True)
            filled[n_spec:n_spec + n_tex] = True
    if raw.is_cuda:
        torch.cuda.current_stream(raw.device).synchronize()     # the asynchronous table copies above
    block[~filled] = np.nan
    out = pd.DataFrame(block.T, columns=float_cols, copy=False)
    out.insert(0, "segment_id", np.asarray(segments["segment_id"]))
    out["geometry"] = segments["geometry"].to_numpy() if len(segments) else None
    from .segment_boundaries import SegmentsFrame
    out = SegmentsFrame(out, copy=False)
    out.label_raster = raster
    out.segment_labels = row_labels
    out.crs = getattr(segments, "crs", None)
    out.transform = getattr(segments, "transform", None)
    aff = getattr(segments, "affine_transformation", None)
    out.affine_transformation = aff if aff is not None else getattr(image, "affine_transformation", None)
    return out

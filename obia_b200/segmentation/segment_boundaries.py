"""Drop-in for obia/segmentation/segment_boundaries.py (`create_segments`).

Same call: `create_segments(image, segmentation_bands=None, method="slic", **kwargs)`
with kwargs = skimage.segmentation.slic's (`n_segments`, `compactness`,
`max_num_iter`, `sigma`, `convert2lab`, `enforce_connectivity`,
`min_size_factor`, `max_size_factor`, `start_label`, `mask`, ...), same
IndexError / ValueError / Exception behaviour
(/root/reference/obia/segmentation/segment_boundaries.py:18-78).

What differs, on purpose (DESIGN.md "boundary"):
  * the work is done on the GPU in label-raster space; the returned table has
    the reference's columns (`geometry`, `segment_id`) and row order (ascending
    label, one row per 4-connected region, `segment_id = 1..N`) and carries the
    label raster as `.label_raster`.  Polygon geometries are materialised on
    the host only on request (`polygonize=True`): one vectorised outline
    trace of the whole label raster (utils/polygonize.py; shapely Polygons
    when shapely is installed, `SimplePolygon` otherwise) instead of the
    reference's per-label full-raster `rasterio.features.shapes` loop
    (:62-70), which is O(n_segments * H * W).
  * the reference prints the shape of the band stack (:46); this does not.
"""
from __future__ import annotations

import numpy as np
import pandas as pd


class SegmentsFrame(pd.DataFrame):
    """pandas DataFrame [`geometry`, `segment_id`] + the label raster it describes.

    Extra attributes (kept through pandas operations where possible):
      label_raster    (H, W) int32 CUDA tensor; -1 = masked out
      segment_labels  int64 numpy array: label value of each row (row i <-> segment_id i+1)
      crs, transform  copied from the Image
    """
    _metadata = ["label_raster", "segment_labels", "crs", "transform", "slic_result", "affine_transformation",
                 "_pending_mutation"]

    @property
    def _constructor(self):
        return SegmentsFrame

    def materialize_geometry(self, affine_transformation=None):
        """Fill the `geometry` column from the label raster (host step, utils/polygonize.py)."""
        from ..utils.polygonize import polygons_from_labels
        self["geometry"] = polygons_from_labels(self.label_raster.cpu().numpy(), self.segment_labels,
                                                affine_transformation)
        return self

    def to_file(self, file_path, driver=None):
        """`GeoDataFrame.to_file` of the reference (segment.py:55-60, tiling.py:289-291).

        With geopandas + shapely installed the table is handed to geopandas (GeoPackage etc.).
        Without them only GeoJSON can be written (pure Python): `.geojson` / `.json` paths, or
        `driver="GeoJSON"`; other formats raise NotImplementedError naming the missing library.
        """
        import json
        import os

        geoms = list(self["geometry"])
        if any(g is None for g in geoms):
            raise ValueError("to_file needs polygon geometries: call create_segments(..., polygonize=True) "
                             "or .materialize_geometry() first")
        ext = os.path.splitext(str(file_path))[1].lower()
        if driver is None:
            driver = "GeoJSON" if ext in (".geojson", ".json") else None
        if driver != "GeoJSON":
            try:
                import geopandas as gpd
            except ImportError as e:
                raise NotImplementedError("writing this format needs geopandas (GDAL); without it only "
                                          "GeoJSON is supported") from e
            gpd.GeoDataFrame(pd.DataFrame(self), geometry="geometry", crs=self.crs).to_file(file_path, driver=driver)
            return
        cols = [c for c in self.columns if c != "geometry"]
        feats = []
        for i, g in enumerate(geoms):
            props = {}
            for c in cols:
                v = self[c].iloc[i]
                v = v.item() if hasattr(v, "item") else v
                props[c] = None if (isinstance(v, float) and v != v) else v
            feats.append({"type": "Feature", "properties": props, "geometry": g.__geo_interface__})
        doc = {"type": "FeatureCollection", "features": feats}
        if self.crs:
            doc["crs"] = {"type": "name", "properties": {"name": str(self.crs)}}
        with open(file_path, "w") as f:
            json.dump(doc, f)


def normalize_band(band):
    """segment_boundaries.py:11-16 (host helper kept for API parity)."""
    return (band - np.min(band)) / (np.max(band) - np.min(band))


def _epsg_string(crs):
    try:
        import pyproj
        return f"EPSG:{pyproj.CRS(crs).to_epsg()}"
    except Exception:
        return crs


def _apply_image_mutation(image, raw, minmax_dev, stream=None, after=None):
    """Side effect of segment_boundaries.py:31-33: every band of `img_data` normalised in place.

    With `stream` the normalise + device-to-host copy is enqueued on that side stream once the
    work already queued on `after` (the main stream) is done, so it overlaps the SLIC kernels;
    the caller synchronises `stream` before returning to the user.
    """
    import contextlib

    import torch

    from .. import pipeline

    data = image.img_data
    if isinstance(data, torch.Tensor) and data.is_cuda and data.dtype == torch.float32 and data.is_contiguous():
        # always derived from the cached RAW values (never img_data in place: a second call would
        # rescale the already normalised raster with the raw range)
        if data.data_ptr() % 16 == 0 and tuple(data.shape) == tuple(raw.shape):
            pipeline.normalize_to(raw, data, minmax_dev)
        else:
            tmp = raw.clone()
            pipeline.normalize_inplace(tmp, minmax_dev)
            data.copy_(tmp)
        if hasattr(image, "_note_mutation"):
            image._note_mutation()
        return
    with (torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()):
        if stream is not None and after is not None:
            stream.wait_stream(after)
        host = None
        if isinstance(data, torch.Tensor):
            host = data
        else:
            arr = np.asarray(data)
            if arr.dtype == np.float32 and arr.flags.c_contiguous and arr.flags.writeable:
                host = torch.from_numpy(arr)
        if (host is not None and not host.is_cuda and host.dtype == torch.float32 and host.is_contiguous()
                and host.is_pinned() and host.data_ptr() % 16 == 0):
            # page-locked img_data: the kernel streams the normalised raster straight into it over
            # PCIe (no staging copy; the copy engine stays free for the SLIC path's small read-backs)
            pipeline.normalize_to_host(raw, host, minmax_dev)
            if hasattr(image, "_note_mutation"):
                image._note_mutation()
            return
        tmp = raw.clone()
        pipeline.normalize_inplace(tmp, minmax_dev)
        if isinstance(data, torch.Tensor):
            data.copy_(tmp.to(data.device, dtype=data.dtype))
        elif host is not None:
            host.copy_(tmp)
        else:
            arr[...] = tmp.cpu().numpy()
    if hasattr(image, "_note_mutation"):
        image._note_mutation()


def frame_from_labels(labels, start_label, n_labels, connected, image=None, polygonize=False):
    """Build the `[geometry, segment_id]` table from a label raster.

    Row order = ascending label value, then raster order of the region's first
    pixel (what `np.unique` + `rasterio.features.shapes` give at :59-70).
    """
    import torch

    if connected:
        # consecutive labels, one 4-connected region each: no search needed.  Label 0
        # can exist besides start_label=1 (SURVEY.md defect 7): found by one reduction.
        lo = int(start_label)
        has_zero = False
        if start_label == 1:
            has_zero = bool((labels == 0).any().item())
        seg_labels = np.arange(lo, lo + n_labels, dtype=np.int64)
        if has_zero:
            # label 0 regions may be several: number them by component
            seg_labels = None
    else:
        seg_labels = None
    if seg_labels is None:
        # general case: one row per 4-connected region of equal label
        from .. import pipeline
        lab_cc = torch.where(labels < 0, torch.full_like(labels, -1), labels + 1).contiguous()
        comp, _ = pipeline.enforce_connectivity(lab_cc, 0, 2 ** 31 - 2, start_label=0)
        # comp numbers regions by first pixel; order rows by (label, first pixel)
        valid = labels >= 0
        comp_v = comp[valid].to(torch.int64)
        lab_v = labels[valid].to(torch.int64)
        ncomp = int(comp_v.max().item()) + 1 if comp_v.numel() else 0
        comp_label = torch.zeros((ncomp,), dtype=torch.int64, device=labels.device)
        comp_label[comp_v] = lab_v
        order = torch.argsort(comp_label, stable=True)   # comp ids are already in first-pixel order
        seg_labels = comp_label[order].cpu().numpy()
        # rows index regions, not labels: remap the raster to region ids (1..N)
        rank = torch.empty_like(order)
        rank[order] = torch.arange(ncomp, device=labels.device)
        region = torch.full_like(labels, -1)
        region[valid] = (rank[comp_v] + 1).to(torch.int32)
        raster, ids = region, np.arange(1, ncomp + 1, dtype=np.int64)
    else:
        raster, ids = labels, np.arange(1, len(seg_labels) + 1, dtype=np.int64)

    geometry = None
    if polygonize:
        from ..utils.polygonize import polygons_from_labels
        geometry = polygons_from_labels(raster.cpu().numpy(), seg_labels if raster is labels else ids,
                                        None if image is None else image.affine_transformation)
    gdf = SegmentsFrame({"geometry": geometry, "segment_id": ids}, index=pd.RangeIndex(len(ids)))
    gdf.label_raster = raster
    gdf.segment_labels = seg_labels if raster is labels else ids
    gdf.crs = None if image is None else _epsg_string(image.crs)
    gdf.transform = None if image is None else image.transform
    gdf.affine_transformation = None if image is None else image.affine_transformation
    return gdf


def _create_segments_quickshift(image, segmentation_bands, mutate_image, polygonize, kwargs):
    """`method="quickshift"` (segment_boundaries.py:48-49): same normalisation side effect and band
    validation as the slic branch, then `skimage.segmentation.quickshift(img, **kwargs)` on the GPU
    (csrc/quickshift.cu).  Labels start at 0; one table row per 4-connected region of a label, in
    ascending label order, like `np.unique` + `rasterio.features.shapes` (:59-70)."""
    from .. import pipeline
    raw = image.device_raw()
    num_bands = int(raw.shape[2])
    if segmentation_bands is None:
        segmentation_bands = list(range(num_bands))
    if mutate_image:
        minmax, _ = pipeline.band_minmax(raw)
        _apply_image_mutation(image, raw, minmax)
    for band in segmentation_bands:
        if band >= num_bands or band < 0:
            raise IndexError(f"Band index {band} out of range. Available bands indices: 0 to {num_bands - 1}.")
    labels, n = pipeline.quickshift_labels(raw, segmentation_bands, **kwargs)
    gdf = frame_from_labels(labels, 0, n, connected=False, image=image, polygonize=polygonize)
    gdf.slic_result = None
    gdf._pending_mutation = None
    return gdf


def create_segments(image, segmentation_bands=None, method="slic", *, mutate_image=True,
                    polygonize=False, _defer_mutation_sync=False, **kwargs):
    """
    :param image: Image (obia_b200.handlers.geotif.Image) -- `img_data` (H, W, C) float32.
    :param segmentation_bands: band indices used for segmentation (None = all).
    :param method: 'slic' (the GPU hot path) or 'quickshift'.
    :param mutate_image: reproduce the reference's in-place normalisation of `image.img_data`.
    :param polygonize: also build the polygon geometries on the host (optional; utils/polygonize.py).
    :param kwargs: skimage.segmentation.slic keyword arguments.
    :return: SegmentsFrame with columns `geometry`, `segment_id`.
    """
    from .. import pipeline

    if method == "quickshift":
        return _create_segments_quickshift(image, segmentation_bands, mutate_image, polygonize, kwargs)
    if method != "slic":
        raise Exception('An unknown segmentation method was requested.')

    raw = image.device_raw()
    num_bands = int(raw.shape[2])
    if segmentation_bands is None:
        segmentation_bands = list(range(num_bands))
    # same validation and message as :38-40; note the reference normalises (mutates)
    # every band BEFORE validating the indices (:31-33)
    bad = [b for b in segmentation_bands if b >= num_bands or b < 0]

    if bad and mutate_image:
        minmax, _ = pipeline.band_minmax(raw)
        _apply_image_mutation(image, raw, minmax)
    for band in bad:
        raise IndexError(f"Band index {band} out of range. Available bands indices: 0 to {num_bands - 1}.")

    import torch
    side = None
    try:
        pre = None
        if mutate_image and kwargs.get("mask", None) is None:
            # the band ranges are known after one pass: start writing the normalised raster back
            # on a side stream while SLIC runs (the reference mutates img_data first, :31-33)
            pre = pipeline.band_minmax(raw)
            if not (isinstance(image.img_data, torch.Tensor) and image.img_data.is_cuda):
                side = torch.cuda.Stream(device=raw.device)
                _apply_image_mutation(image, raw, pre[0], stream=side,
                                      after=torch.cuda.current_stream(raw.device))
        res = pipeline.slic_labels(raw, segmentation_bands, minmax=pre, **kwargs)
    except Exception:
        # the reference has already normalised img_data when slic() raises (:31-33 run first)
        if mutate_image and side is None:
            minmax, _ = pipeline.band_minmax(raw)
            _apply_image_mutation(image, raw, minmax)
        if side is not None:
            side.synchronize()
        raise
    pending = None
    if side is not None:
        if _defer_mutation_sync:
            pending = side      # segment() waits for the write-back after the statistics are queued
        else:
            side.synchronize()
    elif mutate_image:
        _apply_image_mutation(image, raw, res.minmax)

    connected = bool(kwargs.get("enforce_connectivity", True))
    gdf = frame_from_labels(res.labels, res.start_label, res.n_labels, connected, image=image,
                            polygonize=polygonize)
    gdf.slic_result = res
    gdf._pending_mutation = pending
    return gdf

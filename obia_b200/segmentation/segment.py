"""Drop-in for obia/segmentation/segment.py (`segment`, `Segments`).

Same signature and return object as the reference
(/root/reference/obia/segmentation/segment.py:10-93): `segment()` =
`create_segments` (SLIC on the GPU) followed by `create_objects` (zonal
statistics on the GPU); `calc_min` / `calc_max` are not forwarded, exactly like
the reference (:87-91), so they are always on.
"""
from __future__ import annotations

import numpy as np

from .segment_boundaries import create_segments
from .segment_statistics import create_objects


class Segments:
    _segments = None
    segments = None
    method = None
    params = {}   # class-level shared dict, as in the reference (:33)

    def __init__(self, _segments, segments, method, **kwargs):
        self._segments = _segments
        self.segments = segments
        self.method = method
        self.params.update(kwargs)

    def to_segmented_image(self, image):
        """Overlay segment boundaries (yellow) on a PIL image (segment.py:41-53)."""
        from PIL.Image import Image as PILImage
        from PIL.Image import fromarray
        if not isinstance(image, PILImage):
            raise TypeError('Input must be a PIL Image')
        img = np.array(image).astype(np.float64) / 255.0
        if img.ndim == 2:
            img = np.stack([img] * 3, axis=-1)
        lab = self._segments.label_raster.cpu().numpy()
        edge = np.zeros(lab.shape, dtype=bool)
        edge[:, :-1] |= lab[:, :-1] != lab[:, 1:]
        edge[:-1, :] |= lab[:-1, :] != lab[1:, :]
        img[edge, :3] = (1.0, 1.0, 0.0)
        return fromarray((img * 255).astype(np.uint8))

    def write_segments(self, file_path):
        """segment.py:55-60.  Geometries are traced from the label raster on first use."""
        if any(g is None for g in self.segments["geometry"]):
            image_affine = getattr(self._segments, "affine_transformation", None)
            self._segments.materialize_geometry(image_affine)
            self.segments["geometry"] = self._segments["geometry"].to_numpy()
        self.segments.to_file(file_path)


def segment(
    image,
    segmentation_bands=None,
    statistics_bands=None,
    method="slic",
    calc_mean=True,
    calc_variance=True,
    calc_skewness=True,
    calc_kurtosis=True,
    calc_contrast=True,
    calc_dissimilarity=True,
    calc_homogeneity=True,
    calc_ASM=True,
    calc_energy=True,
    calc_correlation=True,
    **kwargs,
):
    """`create_segments` then `create_objects` (reference segment.py:63-93: same positional order,
    names and defaults; `calc_min` / `calc_max` cannot be switched off here, as in the reference)."""
    column_flags = dict(calc_mean=calc_mean, calc_variance=calc_variance, calc_skewness=calc_skewness,
                        calc_kurtosis=calc_kurtosis, calc_contrast=calc_contrast,
                        calc_dissimilarity=calc_dissimilarity, calc_homogeneity=calc_homogeneity,
                        calc_ASM=calc_ASM, calc_energy=calc_energy, calc_correlation=calc_correlation)
    # the in-place normalisation of image.img_data (a side effect of create_segments) is written back
    # on a side stream; inside segment() it only has to be complete when segment() returns
    boundaries = create_segments(image, segmentation_bands=segmentation_bands, method=method,
                                 _defer_mutation_sync=True, **kwargs)
    pending = getattr(boundaries, "_pending_mutation", None)
    try:
        features = create_objects(boundaries, image, spectral_bands=statistics_bands, **column_flags)
    finally:
        if pending is not None:
            pending.synchronize()
            boundaries._pending_mutation = None
    slic_kwargs = {k: v for k, v in kwargs.items() if k not in ("mutate_image", "polygonize")}
    return Segments(boundaries, features, method, **slic_kwargs)

"""Drop-in for the hot-path consumers in obia/utils/utils.py.

`label_segments` (/root/reference/obia/utils/utils.py:12-34) joins labelled points to the segments
that contain them (`gpd.sjoin(..., predicate='intersects')`), sets `feature_class` where all points of
a segment agree and reports the `segment_id` of mixed segments.  It is the step between the feature
table of `create_objects` and `classify` (classification/classify.py:83, :125), i.e. the consumer of
the column contract.  Here the join is a pixel look-up in the label raster the table carries (a point
intersects the segment whose pixel it falls in); tables without a raster (arbitrary polygons) are
burnt into one first (utils/rasterize.py).
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def _point_xy(g):
    gi = g if isinstance(g, dict) else getattr(g, "__geo_interface__", None)
    if gi is not None and gi.get("type") == "Point":
        return float(gi["coordinates"][0]), float(gi["coordinates"][1])
    if hasattr(g, "x") and hasattr(g, "y"):
        return float(g.x), float(g.y)
    x, y = g[:2]
    return float(x), float(y)


def label_segments(segments, labelled_points):
    """
    :param segments: segments / feature table (`geometry`, `segment_id`, ...).
    :param labelled_points: table with a `geometry` column of points (or `x`, `y` columns) and a `class` column.
    :return: (table of the labelled segments with a `feature_class` column, list of mixed segment ids)
    """
    from .rasterize import rasterize_polygons, world_to_pixel

    raster = getattr(segments, "label_raster", None)
    seg_ids = np.asarray(segments["segment_id"])
    if raster is not None and getattr(segments, "segment_labels", None) is not None:
        all_labels = np.asarray(segments.segment_labels, dtype=np.int64)
        # rows may have been filtered: the label of a row follows its segment_id (1..N)
        row_label = all_labels[seg_ids - 1] if len(all_labels) != len(seg_ids) or seg_ids.max(initial=0) > len(seg_ids) \
            else all_labels
        lab = raster.cpu().numpy()
        aff = getattr(segments, "affine_transformation", None)
    else:
        aff = getattr(segments, "affine_transformation", None)
        geoms = list(segments["geometry"])
        b = np.array([g.bounds for g in geoms])
        inv = world_to_pixel(aff)
        corners = inv(np.array([[b[:, 0].min(), b[:, 1].min()], [b[:, 2].max(), b[:, 3].max()],
                                [b[:, 0].min(), b[:, 3].max()], [b[:, 2].max(), b[:, 1].min()]]))
        H, W = int(np.ceil(corners[:, 1].max())), int(np.ceil(corners[:, 0].max()))
        lab = rasterize_polygons(geoms, H, W, aff).cpu().numpy()
        row_label = np.arange(1, len(geoms) + 1)
    if "geometry" in labelled_points:
        pts = np.array([_point_xy(g) for g in labelled_points["geometry"]], dtype=np.float64).reshape(-1, 2)
    else:
        pts = np.stack([np.asarray(labelled_points["x"], float), np.asarray(labelled_points["y"], float)], 1)
    px = world_to_pixel(aff)(pts)
    col, row = np.floor(px[:, 0]).astype(np.int64), np.floor(px[:, 1]).astype(np.int64)
    ok = (row >= 0) & (row < lab.shape[0]) & (col >= 0) & (col < lab.shape[1])
    hit = np.full(len(pts), -1, dtype=np.int64)
    hit[ok] = lab[row[ok], col[ok]]
    classes = np.asarray(labelled_points["class"])
    label_to_row = {int(l): i for i, l in enumerate(row_label)}
    per_row = {}
    for h, c in zip(hit, classes):
        i = label_to_row.get(int(h))
        if i is not None:
            per_row.setdefault(i, set()).add(c)
    labelled = segments.copy()
    feature_class = np.full(len(labelled), None, dtype=object)
    mixed = []
    for i, cs in sorted(per_row.items()):
        if len(cs) == 1:
            feature_class[i] = next(iter(cs))
        else:
            mixed.append(seg_ids[i].item() if hasattr(seg_ids[i], "item") else seg_ids[i])
    labelled["feature_class"] = feature_class
    labelled = labelled[pd.notna(labelled["feature_class"])]
    return labelled, mixed

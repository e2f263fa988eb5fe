"""ctypes binding of libobia_b200.so (the C ABI declared in include/obia_b200.h).

The library is built in-tree by `obia_b200/csrc/build.sh` (or
`__graft_entry__.build()`); there is NO CPU fallback: if the shared object is
missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OBIA_B200_LIB") or os.path.join(_HERE, "_lib", "libobia_b200.so")   # (override: kernel experiments)

_i32, _i64 = ctypes.c_int32, ctypes.c_int64
_f32, _f64 = ctypes.c_float, ctypes.c_double
_vp = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/obia_b200.h one to one
SIGNATURES = {
    "obia_b200_last_error": (ctypes.c_char_p, []),
    "obia_b200_version": (ctypes.c_int, []),
    "obia_b200_launch_count": (_i64, []),
    "obia_b200_profile_enable": (ctypes.c_int, [ctypes.c_int]),
    "obia_b200_profile_read": (ctypes.c_int, [_vp, _vp]),
    "obia_b200_band_minmax": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "obia_b200_normalize_inplace": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "obia_b200_normalize_to": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp, _i32, _vp]),
    "obia_b200_slic_features": (ctypes.c_int, [_vp, _i64, _i64, _i32, _vp, _i32, _vp, _vp, _f32, _f32,
                                               _i32, _f32, _vp, _i64, _vp]),
    "obia_b200_gaussian_planar": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i32, _vp, _i32, _vp,
                                                 _i32, _f32, _vp]),
    "obia_b200_mask_kmeans_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "obia_b200_mask_kmeans": (ctypes.c_int, [_vp, _i64, _vp, _i64, _i32, _i64, _i64, _vp, _vp]),
    "obia_b200_mask_sample_indices": (ctypes.c_int, [_i64, _i64, _vp, _vp]),
    "obia_b200_nearest_centroid": (ctypes.c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "obia_b200_slic_workspace_bytes": (_i64, [_i64, _i64, _i32, _i64, _i32, _i32]),
    "obia_b200_slic_iterate": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _i64, _f32,
                                              _i32, _i32, _i32, _i32, _i32, _i32, _f64, _vp, _vp]),
    "obia_b200_slic_iterate_spacing": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _i64, _f32,
                                                      _i32, _i32, _i32, _i32, _i32, _i32, _f64, _f32, _f32, _vp, _vp]),
    "obia_b200_slic_iterate_fast": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _i64, _f32,
                                                   _i32, _i32, _i32, _i32, _i32, _i32, _f64, _vp, _vp]),
    "obia_b200_slic_sweep_fast": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _i64, _f32, _i32,
                                                 _i32, _i32, _i32, _i32, _f64, _i64, _i64, _vp, _vp]),
    "obia_b200_slic_fast_variant": (ctypes.c_int, [_i32]),
    "obia_b200_slic_begin": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, _i32, _i64, _i32, _i32, _i32, _vp, _vp]),
    "obia_b200_slic_sweep": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _i64, _f32, _i32,
                                            _i32, _i32, _i32, _i32, _f64, _i64, _i64, _vp, _vp]),
    "obia_b200_slic_band_check": (ctypes.c_int, [_vp, _i64, _i32, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i32,
                                                 _vp, _vp]),
    "obia_b200_slic_update_max_color": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i32,
                                                       _i64, _i32, _i32, _i32, _vp]),
    "obia_b200_slic_finish_sweep": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i32, _i64, _i32, _i32, _f64, _vp]),
    "obia_b200_quickshift_workspace_bytes": (_i64, [_i64, _i64]),
    "obia_b200_quickshift": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _f32, _f32, _vp, _vp]),
    "obia_b200_connectivity_workspace_bytes": (_i64, [_i64, _i64]),
    "obia_b200_enforce_connectivity": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i32, _vp,
                                                      _vp]),
    "obia_b200_connectivity_strip_begin": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i32, _i32, _i64, _i64,
                                                          _i32, _vp, _vp]),
    "obia_b200_connectivity_strip_finish": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i32,
                                                           _i64, _vp, _vp]),
    "obia_b200_zonal_workspace_bytes": (_i64, [_i64, _i32]),
    "obia_b200_zonal_stats": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i32, _vp, _i32, _i64, _f64, _vp, _vp,
                                             _vp]),
    "obia_b200_zonal_stats_range": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i32, _vp, _i32, _i64, _i64, _i32, _f64,
                                                   _vp, _vp, _vp]),
    "obia_b200_rasterize_polygons": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _vp]),
    "obia_b200_window_stats": (ctypes.c_int, [_vp, _i64, _i32, _vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _vp]),
    "obia_b200_window_mask_copy": (ctypes.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _vp, _i32, _vp]),
    "obia_b200_window_features": (ctypes.c_int, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _i64, _i32, _i32, _i32,
                                                 _f32, _vp, _i64, _i64, _vp]),
    "obia_b200_mask_kmeans_batch_workspace_bytes": (_i64, [_i64, _i64]),
    "obia_b200_mask_kmeans_batch": (ctypes.c_int, [_vp, _i64, _vp, _i32, _i32, _vp, _vp, _vp, _i64, _i64, _i64, _i32,
                                                   _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "obia_b200_slic_batch_workspace_bytes": (_i64, [_i64, _i64, _i32]),
    "obia_b200_slic_batch_prepare": (ctypes.c_int, [_vp, _i64, _i32]),
    "obia_b200_slic_iterate_batch": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _i32,
                                                    _i64, _i32, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "obia_b200_enforce_connectivity_windows": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _i64, _i32, _vp, _vp]),
    "obia_b200_tiled_paint": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp,
                                             _i64, _vp, _vp]),
    "obia_b200_tiled_seam_import": (ctypes.c_int, [_vp, _i64, _i32, _i32, _vp, _i32, _i32, _vp, _i64, _i32, _vp, _vp, _vp,
                                                   _vp, _vp, _vp, _vp]),
    "obia_b200_tiled_white_prepare": (ctypes.c_int, [_vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _i64, _vp, _vp,
                                                     _vp, _vp, _i64, _vp, _i32, _vp]),
    "obia_b200_texture_workspace_bytes": (_i64, [_i64]),
    "obia_b200_texture_stats": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i32, _vp, _i32, _i64, _i32, _vp, _vp,
                                               _vp]),
}

_LIB = None


class ObiaB200Error(RuntimeError):
    pass


class CandidateOverflowError(ObiaB200Error, ValueError):
    """A SLIC tile exceeded the 1024 candidate centres the kernels stage per tile.  A ValueError like
    the errors scikit-image raises for unusable inputs, so `create_tiled_segments` skips the tile
    (obia/utils/tiling.py:149-150) instead of aborting."""


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(`python -c 'import __graft_entry__ as g; g.build()'` or obia_b200/csrc/build.sh). "
            "obia_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the build disagree
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().obia_b200_last_error()
        raise ObiaB200Error(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

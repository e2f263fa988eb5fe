"""Thin test-side wrappers that call single C-ABI entry points with numpy data."""
import ctypes

import numpy as np
import torch

from obia_b200 import _lib, pipeline, slic_host


def synth_raster(H, W, C, seed=0, noise=0.05, dtype=np.float32, quantize=False):
    """Piecewise-smooth multiband raster (sum of low-frequency sinusoids + noise), SURVEY.md 8d."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[:H, :W].astype(np.float64)
    bands = []
    for c in range(C):
        fy, fx, ph = rng.uniform(0.01, 0.08, 3)
        a = rng.uniform(0.5, 1.0)
        b = a * (np.sin(yy * fy + ph) + np.cos(xx * fx - ph) + np.sin((yy + xx) * (fy + fx) / 3))
        bands.append(b)
    img = np.stack(bands, -1)
    img = (img - img.min()) / (img.max() - img.min())
    img = img + rng.normal(0, noise, img.shape)
    if quantize:
        img = np.clip(np.round(img * 255), 0, 255)
    return np.ascontiguousarray(img.astype(dtype))


def to_planar(features_hwc):
    """(H, W, C) float32 numpy -> CUDA planar [C][H][pitch] tensor + pitch."""
    H, W, C = features_hwc.shape
    pitch = (W + 31) // 32 * 32
    t = torch.zeros((C, H, pitch), dtype=torch.float32, device="cuda")
    t[:, :, :W] = torch.from_numpy(np.ascontiguousarray(np.moveaxis(features_hwc, -1, 0))).cuda()
    return t, pitch


def run_slic_iterate(features_hwc, mask, centres_yxc, step, iters, start_label=1, ignore_color=False, slic_zero=False,
                     fix_scale=None, fast=False):
    """Run obia_b200_slic_iterate on oracle-prepared features/centres.

    centres_yxc: (n, 2 + C) float32 rows (cy, cx, colour...).  Returns (labels, centres_out)."""
    lib = _lib.load()
    H, W, C = features_hwc.shape
    n = centres_yxc.shape[0]
    step_y, step_x = slic_host.window_steps(H, W, n)
    feats, pitch = to_planar(features_hwc)
    centres = torch.from_numpy(np.ascontiguousarray(centres_yxc, dtype=np.float32)).cuda()
    labels = torch.empty((H, W), dtype=torch.int32, device="cuda")
    status = torch.zeros((4,), dtype=torch.int32, device="cuda")
    mask_t = None if mask is None else torch.from_numpy(np.ascontiguousarray(mask, dtype=np.uint8)).cuda()
    if fix_scale is None:
        fix_scale = slic_host.fixed_point_scale(float(np.abs(features_hwc).max()) * 2 + 1e-6, H, W, step_y, step_x)
    ws = torch.empty((lib.obia_b200_slic_workspace_bytes(H, W, C, n, step_y, step_x),), dtype=torch.uint8,
                     device="cuda")
    p = pipeline._p
    iterate = lib.obia_b200_slic_iterate_fast if fast else lib.obia_b200_slic_iterate
    _lib.check(iterate(p(feats), p(mask_t), p(centres), p(labels), p(ws), H, W, pitch, C, n,
                                          float(step), step_y, step_x, int(iters), int(start_label),
                                          int(ignore_color), int(slic_zero), float(fix_scale), p(status),
                                          pipeline._stream_ptr()), "slic_iterate")
    torch.cuda.synchronize()
    assert int(status[0].item()) == 0
    return labels.cpu().numpy(), centres.cpu().numpy()


def oracle_features(raw_hwc, bands, compactness, convert2lab=None, sigma=0, mask=None):
    """Features exactly as the oracle's slic() prepares them (after obia's normalise)."""
    import slic_oracle as so
    img = raw_hwc.copy()
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(img.shape[2]):
            img[:, :, i] = so.normalize_band(img[:, :, i])
    sel = img[:, :, bands]
    _, st = so.slic(sel, n_segments=4, compactness=compactness, max_num_iter=0, convert2lab=convert2lab,
                    sigma=sigma, enforce_connectivity=False, mask=mask, return_state=True)
    return st["features"]

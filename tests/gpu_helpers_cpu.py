"""Synthetic rasters for the CPU-only tests (no torch.cuda / obia_b200.pipeline import)."""
import numpy as np


def synth_raster_cpu(H, W, C, seed=0, noise=0.05):
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[:H, :W].astype(np.float64)
    bands = []
    for c in range(C):
        fy, fx, ph = rng.uniform(0.01, 0.08, 3)
        a = rng.uniform(0.5, 1.0)
        bands.append(a * (np.sin(yy * fy + ph) + np.cos(xx * fx - ph) + np.sin((yy + xx) * (fy + fx) / 3)))
    img = np.stack(bands, -1)
    img = (img - img.min()) / (img.max() - img.min()) + rng.normal(0, noise, img.shape)
    return np.ascontiguousarray(img.astype(np.float32))

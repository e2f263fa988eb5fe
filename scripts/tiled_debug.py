import sys, os, cProfile, pstats, io
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_helpers_cpu import synth_raster_cpu
from obia_b200.utils.tiling import create_tiled_segments

def mask_of(H, W):
    yy, xx = np.mgrid[:H, :W]
    return (np.sin(yy / 45.0) + np.cos(xx / 35.0)) > -1.1

def run(raw, mask, batched, **kw):
    l, n, _ = create_tiled_segments(raw, None, mask, return_labels=True, polygons=False, batched=batched, **kw)
    return l.cpu().numpy(), n

cases = [
    ("s2 sl0", (450, 450, 5), dict(tile_size=150, buffer=40, crown_radius=4, compactness=0.5, start_label=0)),
    ("s2 sl1", (450, 450, 5), dict(tile_size=150, buffer=40, crown_radius=4, compactness=0.5)),
    ("s0 sl0", (620, 830, 4), dict(tile_size=200, buffer=30, crown_radius=5, compactness=0.2, start_label=0)),
]
for name, (H, W, C), kw in cases:
    raw = synth_raster_cpu(H, W, C, seed=11); mask = mask_of(H, W)
    a, na = run(raw, mask, True, **kw); b, nb = run(raw, mask, False, **kw)
    d = a != b
    print(name, "n", na, nb, "diff px", int(d.sum()))
    if d.any():
        ys, xs = np.nonzero(d)
        print("  bbox", ys.min(), ys.max(), xs.min(), xs.max())
        # same partition up to renumbering?
        pairs = np.unique(np.stack([a[d], b[d]], 1), axis=0)
        print("  distinct (a,b) pairs", len(pairs), pairs[:10].tolist())
        same_partition = len(np.unique(np.stack([a.ravel(), b.ravel()], 1), axis=0)) == len(np.unique(a))
        print("  same partition:", same_partition)

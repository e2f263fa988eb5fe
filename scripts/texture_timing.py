"""Time the texture kernel (K5) on the bench workload c2 and on c1-like data.  GPU box only."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from obia_b200 import pipeline

def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out

for (H, W, C, n, comp) in ((10000, 10000, 8, 200000, 0.1), (10000, 10000, 8, 200000, 10.0), (2048, 2048, 3, 3000, 10.0)):
    raw = bench.synth_raster_cuda(H, W, C, 2, torch.device("cuda"))
    res = pipeline.slic_labels(raw, None, n_segments=n, compactness=comp)
    labels = res.labels
    mx = int(labels.max())
    t_z, _ = timed(lambda: pipeline.zonal_stats(labels, raw, None, max_label=mx))
    t_t, f = timed(lambda: pipeline.texture_stats(labels, raw, None, max_label=mx))
    print(f"{H}x{W}x{C} n={n} c={comp}: segments={mx} zonal {t_z:.2f} ms  texture {t_t:.2f} ms  "
          f"({H*W*C/t_t/1e3:.1f} M band-pixels/s)  nan={int(torch.isnan(f).any())}", flush=True)
    del raw, res, labels

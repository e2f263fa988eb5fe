"""Wall-clock breakdown of one resident step (c2) by host stage, to find host-side gaps.  GPU box only."""
import os, sys, time, cProfile, pstats
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from obia_b200 import pipeline
raw = bench.synth_raster_cuda(10000, 10000, 8, 2, torch.device("cuda"))
kw = dict(n_segments=200000, compactness=0.1, max_num_iter=10)
for _ in range(2):
    res = pipeline.slic_labels(raw, None, **kw); st = pipeline.zonal_stats(res.labels, raw, None, max_label=res.n_labels)
torch.cuda.synchronize()
pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable()
for _ in range(3):
    res = pipeline.slic_labels(raw, None, **kw); st = pipeline.zonal_stats(res.labels, raw, None, max_label=res.n_labels)
torch.cuda.synchronize(); pr.disable()
print("ms per step", (time.perf_counter() - t0) / 3 * 1e3)
pstats.Stats(pr).sort_stats("tottime").print_stats(22)

"""Small driver for ncu: two SLIC sweeps of the tolerance-mode kernel on c2.  GPU box only."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from obia_b200 import pipeline
comp = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
raw = bench.synth_raster_cuda(10000, 10000, 8, 2, torch.device("cuda"))
res = pipeline.slic_labels(raw, None, n_segments=200000, compactness=comp, max_num_iter=3, enforce_connectivity=False)
torch.cuda.synchronize()
print("ok", int(res.labels.max()))

import os, sys, torch
sys.path.insert(0, "/root/repo")
import bench
from obia_b200 import pipeline
H = W = 4000
raw = bench.synth_raster_cuda(H, W, 8, 2, torch.device("cuda"))
res = pipeline.slic_labels(raw, None, n_segments=32000, compactness=10.0)
mx = int(res.labels.max())
f = pipeline.texture_stats(res.labels, raw, None, max_label=mx)
torch.cuda.synchronize()
print("ok", mx)

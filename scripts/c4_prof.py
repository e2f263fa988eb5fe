"""c4-shaped SLIC (64 bands) at reduced size for profiling the many-band assign kernel.  GPU box only."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from obia_b200 import _lib, pipeline
if len(sys.argv) > 2:   # experimental build: obia_b200/_lib/libobia_exp_<variant>.so
    _lib.LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "obia_b200", "_lib",
                                 f"libobia_exp_{sys.argv[2]}.so")
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
raw = bench.synth_raster_cuda(H, W, 64, 4, torch.device("cuda"))
n = int(50000 * (H / 8192) ** 2)
for it in range(2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = pipeline.slic_labels(raw, None, n_segments=n, compactness=0.3, max_num_iter=3, enforce_connectivity=False)
    e1.record(); torch.cuda.synchronize()
print(sys.argv[2] if len(sys.argv) > 2 else "default", "c4-like", H, "ms for 3 sweeps", e0.elapsed_time(e1), "segments", res.n_labels)

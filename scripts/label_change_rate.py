"""How many pixels change their centre from one SLIC sweep to the next (c2 workload, tolerance mode)?
Decides whether an incremental centre update (records only for changed pixels) would pay."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from obia_b200 import sharded

dev = torch.device("cuda", 0)
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
n_seg = int(round(200000 * (H / 10000) ** 2))
for comp in (0.1, 10.0):
    raw = bench.synth_raster_cuda(H, W, 8, 1, dev)
    s = sharded.ShardedSlic(raw, 0, H, None, n_segments=n_seg, compactness=comp, max_num_iter=10)
    mm, fl = s.local_minmax()
    s.prepare(mm, fl)
    s.begin()
    prev = None
    rates = []
    for it in range(10):
        s.sweep()
        s.finish_sweep()
        cur = s.labels.clone()
        if prev is not None:
            rates.append(float((cur != prev).float().mean().item()))
        prev = cur
    print(f"compactness {comp}: fraction of pixels whose centre changed in sweeps 2..10:", [round(r, 4) for r in rates])
    del s, raw

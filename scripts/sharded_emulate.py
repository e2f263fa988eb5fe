"""Emulate W ranks of the sharded path on ONE GPU (LocalComm) and time the connectivity calls per strip.
Usage: python scripts/sharded_emulate.py [world] [rows_per_strip] [width].  GPU box only."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from obia_b200 import sharded

world = int(sys.argv[1]) if len(sys.argv) > 1 else 4
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 2560
W = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
H = rows * world
n = int(round(200000 * H * W / 1e8))
kw = dict(n_segments=n, compactness=0.1, max_num_iter=10)
dev = torch.device("cuda")
split = sharded.split_rows(H, world)
raws = [bench.synth_strip_cuda(r0, h, W, 8, 2 + i, dev) for i, (r0, h) in enumerate(split)]
for it in range(2):
    strips = [sharded.ShardedSlic(raw, r0, H, None, **kw) for raw, (r0, h) in zip(raws, split)]
    # instrument strip_begin / strip_finish
    tb, tf = [], []
    ob, of = sharded.ShardedSlic.strip_begin, sharded.ShardedSlic.strip_finish
    def sb(self, a, b):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = ob(self, a, b); torch.cuda.synchronize()
        tb.append((self.row0, round((time.perf_counter() - t0) * 1e3, 2), self.k_before, self.k_core, self.cc_rounds)); return r
    def sf(self, x):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = of(self, x); torch.cuda.synchronize()
        tf.append((self.row0, round((time.perf_counter() - t0) * 1e3, 2), r)); return r
    sharded.ShardedSlic.strip_begin, sharded.ShardedSlic.strip_finish = sb, sf
    try:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = sharded.run_sharded(strips, sharded.LocalComm(world), None, timings=True)
        torch.cuda.synchronize(); tot = (time.perf_counter() - t0) * 1e3
    finally:
        sharded.ShardedSlic.strip_begin, sharded.ShardedSlic.strip_finish = ob, of
    print(f"iter {it}: total {tot:.1f} ms (sum over {world} strips), mode {res.mode}, n_labels {res.n_labels}")
    print("  stage ms:", {k: round(v, 1) for k, v in res.timings.items()})
    print("  strip_begin (row0, ms, k_before, k_core, rounds):", tb)
    print("  strip_finish (row0, ms, incomplete):", tf)

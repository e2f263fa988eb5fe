import sys, os, cProfile, pstats, io, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from tiled_bench import synth
from obia_b200.utils.tiling import create_tiled_segments
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
dev = torch.device("cuda", 0)
raw, mask = synth(S, S, 4, dev)
kw = dict(tile_size=200, buffer=30, compactness=0.2, crown_radius=5)
create_tiled_segments(raw[:600, :600].contiguous(), None, mask[:600, :600], return_labels=True, polygons=False, **kw)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
create_tiled_segments(raw, None, mask, return_labels=True, polygons=False, **kw)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
# stage split (synchronised: slower than the real run, shares only)
from obia_b200 import batch
batch.TIMINGS = {}
t0 = time.perf_counter()
create_tiled_segments(raw, None, mask, return_labels=True, polygons=False, **kw)
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print("synchronised total %.3f s" % tot)
for k, v in sorted(batch.TIMINGS.items(), key=lambda kv: -kv[1]):
    print("  %-45s %.3f s" % (k, v))
print("  %-45s %.3f s" % ("outside WindowBatch.segment", tot - sum(batch.TIMINGS.values())))

"""Stage split of the batched tiled driver (synchronised per stage: shares, not absolute times).
    python scripts/tiled_profile.py SIZE [--cprofile]"""
import sys, os, cProfile, pstats, io, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from tiled_bench import synth
from obia_b200 import batch, slic_host
from obia_b200.utils import tiling
from obia_b200.utils.tiling import create_tiled_segments
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
dev = torch.device("cuda", 0)
raw, mask = synth(S, S, 4, dev)
kw = dict(tile_size=200, buffer=30, compactness=0.2, crown_radius=5)
create_tiled_segments(raw[:600, :600].contiguous(), None, mask[:600, :600], return_labels=True, polygons=False, **kw)
torch.cuda.synchronize()

passes = {}
def timed(name, fn):
    def wrap(self):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        before = dict(batch.TIMINGS) if batch.TIMINGS is not None else None
        fn(self)
        torch.cuda.synchronize()
        passes[name] = (time.perf_counter() - t0,
                        None if before is None else {k: v - before.get(k, 0.0) for k, v in batch.TIMINGS.items()})
    return wrap
tiling.BatchedTiledSegmenter.run_black = timed("black", tiling.BatchedTiledSegmenter.run_black)
tiling.BatchedTiledSegmenter.run_white = timed("white", tiling.BatchedTiledSegmenter.run_white)
_fin = tiling.BatchedTiledSegmenter.finalize
def fin(self):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = _fin(self); torch.cuda.synchronize()
    passes["finalize"] = (time.perf_counter() - t0, None); return r
tiling.BatchedTiledSegmenter.finalize = fin

for label, timings in (("asynchronous run (draws not cached)", None), ("synchronised per stage (draws cached)", {})):
    if timings is None:
        slic_host._CHOICE_CACHE.clear()
    batch.TIMINGS = timings
    passes.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if "--cprofile" in sys.argv and timings is None:
        pr = cProfile.Profile(); pr.enable()
    create_tiled_segments(raw, None, mask, return_labels=True, polygons=False, **kw)
    torch.cuda.synchronize()
    tot = time.perf_counter() - t0
    if "--cprofile" in sys.argv and timings is None:
        pr.disable(); s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25); print(s.getvalue()[:6000])
    print(f"{label}: total {tot:.3f} s ({S * S / 1e6 / tot:.0f} MP/s)")
    for name, (t, split) in passes.items():
        print(f"  {name:9s} {t:.3f} s")
        if split:
            for k, v in sorted(split.items(), key=lambda kv: -kv[1]):
                print(f"      {k:45s} {v:.3f} s")
            print(f"      {'outside WindowBatch':45s} {t - sum(split.values()):.3f} s")

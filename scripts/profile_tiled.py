import cProfile, pstats, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_helpers_cpu import synth_raster_cpu
from obia_b200.utils.tiling import create_tiled_segments
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 1000        # python scripts/profile_tiled.py [size] [tile]
TILE = int(sys.argv[2]) if len(sys.argv) > 2 else 200
raw = torch.from_numpy(synth_raster_cpu(H, W, 4, seed=7)).cuda()
yy, xx = np.mgrid[:H, :W]
mask = torch.from_numpy((np.sin(yy / 90.0) + np.cos(xx / 70.0)) > -1.2).cuda()
kw = dict(tile_size=TILE, buffer=30, crown_radius=8, compactness=0.2)
create_tiled_segments(raw, None, mask, distributed=False, return_labels=True, polygons=False, **kw)   # warm-up
torch.cuda.synchronize()
pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable()
labels, n, _ = create_tiled_segments(raw, None, mask, distributed=False, return_labels=True, polygons=False, **kw)
torch.cuda.synchronize(); pr.disable(); print("total s", time.perf_counter() - t0, "segments", n)
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)

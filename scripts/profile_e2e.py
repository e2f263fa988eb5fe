import cProfile, pstats, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from obia_b200.handlers.geotif import Image
from obia_b200.segmentation.segment import segment
dev = torch.device("cuda")
H = W = 10000; C = 8
raw = bench.synth_raster_cuda(H, W, C, 2, dev)
pristine = torch.empty((H, W, C), dtype=torch.float32, pin_memory=True); pristine.copy_(raw)
work = torch.empty((H, W, C), dtype=torch.float32, pin_memory=True)
del raw
kw = dict(n_segments=200000, compactness=0.1, max_num_iter=10)
TEX_OFF = dict(calc_contrast=False, calc_dissimilarity=False, calc_homogeneity=False, calc_ASM=False,
               calc_energy=False, calc_correlation=False)
MODE = sys.argv[1] if len(sys.argv) > 1 else "default"
if MODE == "nomutate":
    kw["mutate_image"] = False
for i in range(3):
    work.copy_(pristine); torch.cuda.synchronize()
    pr = cProfile.Profile()
    t0 = time.perf_counter(); pr.enable()
    img = Image(work.numpy(), "EPSG:32702", [1, 0, 0, -1, 0, 0], None, None)
    t1 = time.perf_counter()
    rawd = img.device_raw(); torch.cuda.synchronize(); t2 = time.perf_counter()
    seg = segment(img, None, None, "slic", **TEX_OFF, **kw)
    torch.cuda.synchronize(); pr.disable(); t3 = time.perf_counter()
    print(f"iter {i}: upload {1e3*(t2-t1):.1f} ms  segment {1e3*(t3-t2):.1f} ms total {1e3*(t3-t0):.1f} ms", flush=True)
    del img, seg, rawd
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)

import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from obia_b200 import pipeline
dev = torch.device("cuda")
for S in (256, 512):
    raw = bench.synth_raster_cuda(S, S, 3, seed=1, device=dev)
    raw = torch.clamp(torch.round(raw * 255), 0, 255)
    n = max(4, int(round(3000 * (S / 2048) ** 2)))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = pipeline.slic_labels(raw, None, n_segments=n, compactness=10, enforce_connectivity=False)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    seg = S * S / res.n_centres
    out, nl = pipeline.enforce_connectivity(res.labels, int(0.5 * seg), int(3 * seg), 1)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"S={S} centres={res.n_centres} slic {1e3*(t1-t0):.1f} ms  connectivity {1e3*(t2-t1):.1f} ms  kept={nl} zero_frac={(out==0).float().mean().item():.3f}", flush=True)

"""K2 on c2 (10000 x 10000 x 8, 207 025 centres): exact vs tolerance-mode kernel, both CTA shapes.
Per-launch time from the library's CUDA events; label agreement + ARI of the whole SLIC path
(fast vs exact) on the full raster.  GPU box only."""
import ctypes, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from obia_b200 import _lib, pipeline

lib = _lib.load()
size = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
raw = bench.synth_raster_cuda(size, size, 8, 2, torch.device("cuda"))
nseg = int(round(200000 * (size / 10000) ** 2))
out = {}


def timed(tag, **kw):
    for _ in range(2):
        res = pipeline.slic_labels(raw, None, **kw)
    torch.cuda.synchronize()
    lib.obia_b200_profile_enable(1)
    t0 = time.perf_counter()
    for _ in range(3):
        res = pipeline.slic_labels(raw, None, **kw)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 3 * 1e3
    lib.obia_b200_profile_enable(0)
    ms, n = ctypes.c_double(0), ctypes.c_int64(0)
    lib.obia_b200_profile_read(ctypes.byref(ms), ctypes.byref(n))
    out[tag] = {"k2_ms_per_launch": ms.value / max(1, n.value), "launches": n.value, "slic_ms": wall,
                "segments": int(res.n_labels)}
    print(tag, out[tag], flush=True)
    return res


def ari(a, b):
    # contingency-table ARI on the device (labels up to a few 1e5): pairs via sparse unique
    a = a.reshape(-1).to(torch.int64); b = b.reshape(-1).to(torch.int64)
    key = a * (int(b.max()) + 1) + b
    _, nij = torch.unique(key, return_counts=True)
    _, ai = torch.unique(a, return_counts=True)
    _, bj = torch.unique(b, return_counts=True)
    c2 = lambda v: (v.double() * (v.double() - 1) / 2).sum().item()
    n = a.numel()
    sij, sa, sb, tot = c2(nij), c2(ai), c2(bj), n * (n - 1) / 2
    exp = sa * sb / tot
    return (sij - exp) / (0.5 * (sa + sb) - exp)


for comp in (0.1, 10.0):
    kw = dict(n_segments=nseg, compactness=comp, max_num_iter=10)
    ex = timed(f"exact_c{comp}", exact=True, **kw)
    lib.obia_b200_slic_fast_variant(8)
    f8 = timed(f"fast8_c{comp}", exact=False, **kw)
    lib.obia_b200_slic_fast_variant(4)
    f4 = timed(f"fast4_c{comp}", exact=False, **kw)
    lib.obia_b200_slic_fast_variant(8)
    for tag, r in (("fast8", f8), ("fast4", f4)):
        agree = float((r.labels == ex.labels).float().mean().item())
        out[f"{tag}_c{comp}"]["agreement_vs_exact"] = agree
        out[f"{tag}_c{comp}"]["ari_vs_exact"] = ari(r.labels, ex.labels)
        print(tag, comp, "agreement", agree, "ARI", out[f"{tag}_c{comp}"]["ari_vs_exact"], flush=True)
    del ex, f8, f4
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "k2_modes.json"), "w"), indent=1)

"""Ablation of the tolerance-mode K2 kernel on c2: per-launch time with parts switched off
(obia_b200_slic_fast_variant bits 8..: 1 no update, 2 one candidate per chunk, 4 no seed).  GPU box only."""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from obia_b200 import _lib, pipeline
lib = _lib.load()
raw = bench.synth_raster_cuda(10000, 10000, 8, 2, torch.device("cuda"))
out = {}
for comp in (0.1, 10.0):
    for warps in (8, 4):
        for dbg, name in ((0, "full"), (1, "no_update"), (3, "no_update_one_candidate"), (4, "no_seed"), (7, "setup_only")):
            lib.obia_b200_slic_fast_variant(warps | (dbg << 8))
            kw = dict(n_segments=200000, compactness=comp, max_num_iter=4, enforce_connectivity=False)
            pipeline.slic_labels(raw, None, **kw)
            torch.cuda.synchronize()
            lib.obia_b200_profile_enable(1)
            pipeline.slic_labels(raw, None, **kw)
            torch.cuda.synchronize()
            lib.obia_b200_profile_enable(0)
            ms, n = ctypes.c_double(0), ctypes.c_int64(0)
            lib.obia_b200_profile_read(ctypes.byref(ms), ctypes.byref(n))
            out[f"c{comp}_w{warps}_{name}"] = ms.value / max(1, n.value)
            print(f"c{comp} warps={warps} {name}: {ms.value / max(1, n.value):.3f} ms/launch", flush=True)
lib.obia_b200_slic_fast_variant(8)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "k2_ablate.json"), "w"), indent=1)

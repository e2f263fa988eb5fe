"""Run the BASELINE.json configurations once each on one GPU and print MP/s (sanity at full size).

    python scripts/run_configs.py [c1 c2 c4 c5a c5b ...]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from obia_b200 import pipeline  # noqa: E402

CONFIGS = {
    # name: (H, W, C, bands, slic kwargs, quantize-to-uint8)
    "c1": (2048, 2048, 3, None, dict(n_segments=3000, compactness=10), True),
    "c2": (10000, 10000, 8, None, dict(n_segments=200000, compactness=0.1), False),
    "c2b": (10000, 10000, 8, None, dict(n_segments=200000, compactness=10), False),
    "c4": (8192, 8192, 64, None, dict(n_segments=50000, compactness=0.3), False),
    "c5a": (20000, 20000, 8, None, dict(n_segments=10000, compactness=1), False),
    "c5b": (20000, 20000, 8, None, dict(n_segments=100000, compactness=10), False),
    "c5c": (20000, 20000, 8, None, dict(n_segments=1000000, compactness=50, max_num_iter=20), False),
}


def c1_raster_cuda(H, W, C, dev, noise=8.0, seed=1):
    """sum of 3 low-frequency sinusoids per band quantised to 0..255 + uniform noise +-8."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    yy = torch.arange(H, device=dev, dtype=torch.float32)[:, None]
    xx = torch.arange(W, device=dev, dtype=torch.float32)[None, :]
    out = torch.empty((H, W, C), dtype=torch.float32, device=dev)
    for c in range(C):
        fr = [(0.004 * (c + 1), 0.003 * (c + 2)), (0.011, 0.007 * (c + 1)), (0.02 * (c + 1), 0.015)]
        v = sum(torch.sin(yy * a + c) + torch.cos(xx * b - c) for a, b in fr)
        v = (v - v.min()) / (v.max() - v.min()) * 255
        u = (torch.rand((H, W), generator=g, device=dev) * 2 - 1) * noise
        out[:, :, c] = torch.clamp(torch.round(v + u), 0, 255)
    return out


def main():
    names = sys.argv[1:] or ["c1", "c2", "c2b", "c4", "c5a", "c5b", "c5c"]
    dev = torch.device("cuda")
    for name in names:
        H, W, C, bands, kw, quant = CONFIGS[name]
        if quant:
            raw = c1_raster_cuda(H, W, C, dev)
        else:
            raw = bench.synth_raster_cuda(H, W, C, seed=1, device=dev)
        torch.cuda.synchronize()
        out = {}
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = pipeline.slic_labels(raw, bands, **kw)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            stats = pipeline.zonal_stats(res.labels, raw, None, max_label=res.n_labels)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            tex = pipeline.texture_stats(res.labels, raw, None, max_label=res.n_labels,
                                         quantise_f64=quant) if C <= 16 else None      # not part of mp_per_s
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            out = dict(config=name, shape=[H, W, C], kw=kw, centres=res.n_centres, segments=res.n_labels,
                       slic_ms=1e3 * (t1 - t0), stats_ms=1e3 * (t2 - t1), mp_per_s=H * W / 1e6 / (t2 - t0),
                       texture_ms=None if tex is None else 1e3 * (t3 - t2),
                       peak_mem_gb=torch.cuda.max_memory_allocated() / 1e9)
        cnt = stats[:, 0, 0].sum().item()
        out["pixels_counted"] = int(cnt)
        out["pixels_labelled"] = int((res.labels >= 0).sum().item())
        print(json.dumps(out), flush=True)
        del raw, res, stats, tex
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()


if __name__ == "__main__":
    main()

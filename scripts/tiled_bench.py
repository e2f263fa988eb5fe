"""Timing of create_tiled_segments (c3 of BASELINE.json: 4-band raster, tile 200, buffer 30) on one GPU.

    python scripts/tiled_bench.py --size 8000 [--tile 200 --buffer 30 --mode masked|fixed --per-tile]
Synthetic raster generated on the device; prints one JSON line per run."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synth(H, W, C, dev):
    g = torch.Generator(device=dev).manual_seed(7)
    yy = torch.arange(H, device=dev, dtype=torch.float32)[:, None]
    xx = torch.arange(W, device=dev, dtype=torch.float32)[None, :]
    raw = torch.empty((H, W, C), dtype=torch.float32, device=dev)
    for c in range(C):
        fy, fx, ph = 0.013 + 0.011 * c, 0.017 + 0.007 * c, 0.5 * c
        band = torch.sin(yy * fy + ph) + torch.cos(xx * fx - ph) + torch.sin((yy + xx) * (fy + fx) / 3)
        raw[:, :, c] = (band + 3) / 6 + 0.05 * torch.randn((H, W), generator=g, device=dev)
    mask = (torch.sin(yy / 90.0) + torch.cos(xx / 70.0)) > -1.2
    return raw, mask


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=8000)
    ap.add_argument("--tile", type=int, default=200)
    ap.add_argument("--buffer", type=int, default=30)
    ap.add_argument("--bands", type=int, default=4)
    ap.add_argument("--mode", default="masked")
    ap.add_argument("--per-tile", action="store_true")
    ap.add_argument("--repeat", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    from obia_b200.utils.tiling import create_tiled_segments
    H = W = args.size
    raw, mask = synth(H, W, args.bands, dev)
    kw = dict(tile_size=args.tile, buffer=args.buffer, compactness=0.2)
    if args.mode == "masked":
        kw["crown_radius"] = 5
        m = mask
    else:
        kw["n_segments"] = 100
        m = None
    # warm-up on a corner (library load, allocator)
    create_tiled_segments(raw[:3 * args.tile, :3 * args.tile].contiguous(), None, None if m is None else m[:3 * args.tile, :3 * args.tile],
                          return_labels=True, polygons=False, batched=not args.per_tile, **kw)
    for _ in range(args.repeat):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        labels, n, _ = create_tiled_segments(raw, None, m, return_labels=True, polygons=False, batched=not args.per_tile, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps({"workload": f"create_tiled_segments {H}x{W}x{args.bands} tile {args.tile} buffer {args.buffer} {args.mode}",
                          "driver": "per-tile" if args.per_tile else "batched", "seconds": dt, "MP_per_s": H * W / 1e6 / dt,
                          "segments": n, "peak_mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)


if __name__ == "__main__":
    main()

"""torchrun check: create_tiled_segments sharded over N GPUs (NCCL seam exchange) equals the 1-GPU result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/tiled_multigpu_check.py [--size 1200 --tile 200 --buffer 30]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1200)
    ap.add_argument("--tile", type=int, default=200)
    ap.add_argument("--buffer", type=int, default=30)
    ap.add_argument("--bands", type=int, default=4)
    ap.add_argument("--device-synth", action="store_true", help="generate the raster on the GPU (large sizes)")
    ap.add_argument("--per-tile", action="store_true", help="TiledSegmenter (one pipeline call per tile)")
    ap.add_argument("--no-single", action="store_true", help="skip the 1-GPU run of rank 0 (timing only)")
    ap.add_argument("--fixed-n", type=int, default=0, help="fixed n_segments, no mask")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    from gpu_helpers_cpu import synth_raster_cpu
    from obia_b200.utils.tiling import create_tiled_segments
    H = W = args.size
    if args.device_synth:
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        from tiled_bench import synth
        raw, mask = synth(H, W, args.bands, dev)
    else:
        raw = synth_raster_cpu(H, W, args.bands, seed=7)
        yy, xx = np.mgrid[:H, :W]
        mask = (np.sin(yy / 90.0) + np.cos(xx / 70.0)) > -1.2
    kw = dict(tile_size=args.tile, buffer=args.buffer, crown_radius=8 if not args.device_synth else 5, compactness=0.2,
              batched=False if args.per_tile else None)
    if args.fixed_n:
        kw["n_segments"] = args.fixed_n
        mask = None
    # warm up NCCL (communicator set-up, first point-to-point) outside the timed call
    w = torch.zeros(1, device=dev)
    dist.all_reduce(w)
    if world > 1:
        if rank % 2 == 0 and rank + 1 < world:
            dist.send(w, rank + 1)
        elif rank % 2 == 1:
            dist.recv(w, rank - 1)
    # warm-up on a corner of the raster (library load, allocator, kernels), then drop the cached sample draws so
    # that the timed runs pay for them
    from obia_b200 import slic_host
    c = 3 * args.tile
    create_tiled_segments(raw[:c, :c], None, None if mask is None else mask[:c, :c], distributed=False,
                          return_labels=True, polygons=False, **kw)
    slic_host._CHOICE_CACHE.clear()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    labels, n, (x0, x1) = create_tiled_segments(raw, None, mask, distributed=True, return_labels=True, polygons=False, **kw)
    torch.cuda.synchronize()
    t_multi = time.perf_counter() - t0
    # gather the column blocks on rank 0
    if not args.no_single:
        full = torch.full((H, W), -2, dtype=torch.int32, device=dev)
        full[:, x0:x1] = labels
        dist.all_reduce(full, op=dist.ReduceOp.MAX)
    ok = True
    if rank == 0 and args.no_single:
        print(f"tiled multi-GPU timing: world={world} size={H} tiles={args.tile} segments={n} "
              f"t_multi={t_multi:.2f}s ({H * W / 1e6 / t_multi:.1f} MP/s)", flush=True)
    elif rank == 0:
        slic_host._CHOICE_CACHE.clear()
        t0 = time.perf_counter()
        single, n1, _ = create_tiled_segments(raw, None, mask, distributed=False, return_labels=True, polygons=False, **kw)
        torch.cuda.synchronize()
        t_single = time.perf_counter() - t0
        same = bool((single == full).all().item())
        ok = same and n == n1
        print(f"tiled multi-GPU check: world={world} size={H} tiles={args.tile} segments multi={n} single={n1} "
              f"identical={same}  t_multi={t_multi:.2f}s ({H * W / 1e6 / t_multi:.1f} MP/s) "
              f"t_single={t_single:.2f}s ({H * W / 1e6 / t_single:.1f} MP/s)", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()

"""torchrun check + timing: global SLIC + zonal statistics on ONE raster sharded by row strips over N GPUs
(NCCL all-reduce of the int64 centre sums every sweep) equals the single-GPU result bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/sharded_check.py [--size 10000 --bands 8 --segments 200000 --compactness 0.1]
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=10000)
    ap.add_argument("--bands", type=int, default=8)
    ap.add_argument("--segments", type=int, default=200000)
    ap.add_argument("--compactness", type=float, default=0.1)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    import bench
    from obia_b200 import pipeline
    from obia_b200.sharded import slic_zonal_distributed, split_rows
    H = W = args.size
    full = bench.synth_raster_cuda(H, W, args.bands, seed=2, device=dev)      # same raster on every rank (seeded)
    r0, h = split_rows(H, world)[rank]
    strip = full[r0:r0 + h].contiguous()
    kw = dict(n_segments=args.segments, compactness=args.compactness, max_num_iter=10)
    if rank != 0:
        del full
        torch.cuda.empty_cache()
    times = []
    for rep in range(1 + args.reps):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        labels, n, stats = slic_zonal_distributed(strip, r0, H, None, None, **kw)
        torch.cuda.synchronize()
        dist.barrier()
        if rep:
            times.append(time.perf_counter() - t0)
    t_multi = min(times)
    gathered = torch.full((H, W), -9, dtype=torch.int32, device=dev)
    gathered[r0:r0 + h] = labels
    dist.all_reduce(gathered, op=dist.ReduceOp.MAX)
    ok = True
    if rank == 0:
        ts = []
        for rep in range(1 + args.reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ref = pipeline.slic_labels(full, None, **kw)
            ref_stats = pipeline.zonal_stats(ref.labels, full, None, max_label=ref.n_labels + 1)
            torch.cuda.synchronize()
            if rep:
                ts.append(time.perf_counter() - t0)
        same = bool(torch.equal(gathered, ref.labels))
        cnt_same = bool(torch.equal(stats[:, :, 0], ref_stats[:, :, 0]))
        mean_err = float((stats[:, :, 1] - ref_stats[:, :, 1]).abs().nan_to_num().max())
        ok = same and cnt_same and n == ref.n_labels
        mp = H * W / 1e6
        print(f"sharded global SLIC: world={world} raster={H}x{W}x{args.bands} segments={n} labels_identical={same} "
              f"counts_identical={cnt_same} max|mean diff|={mean_err:.2e}  "
              f"t_sharded={t_multi * 1e3:.1f} ms ({mp / t_multi:.0f} MP/s)  t_single={min(ts) * 1e3:.1f} ms "
              f"({mp / min(ts):.0f} MP/s)  speed-up {min(ts) / t_multi:.2f}x", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()

"""N ranks over NCCL (torchrun): the sharded path gives, strip by strip, exactly the labels and
statistics of the single-GPU pipeline on the same raster.  Every rank generates the whole raster
from the same seed, runs the single-GPU reference on its own GPU and compares its strip.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/sharded_check.py [H] [W]
"""
import json, os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from obia_b200 import pipeline, sharded

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
H = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
W = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
out = {}
for name, kw in (("c0.1", dict(compactness=0.1)), ("c10", dict(compactness=10.0)),
                 ("c1_exact_sl0", dict(compactness=1.0, exact=True, start_label=0))):
    kw = dict(n_segments=int(round(200000 * H * W / 1e8)), max_num_iter=10, **kw)
    raw = bench.synth_raster_cuda(H, W, 8, 7, dev)
    ref = pipeline.slic_labels(raw, None, **kw)
    ref_stats = pipeline.zonal_stats(ref.labels, raw, None, max_label=ref.n_labels + 1)
    row0, h = sharded.split_rows(H, world)[rank]
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    res = sharded.slic_zonal_distributed(raw[row0:row0 + h].contiguous(), row0, H, None, None, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    same = bool(torch.equal(res.labels[0], ref.labels[row0:row0 + h])) and res.n_labels == ref.n_labels
    stats_ok = True
    if res.mode["stats"] == "label-range":
        lo = res.label_lo[0]
        a, b = res.stats[0], ref_stats[lo:lo + res.stats[0].shape[0]]
        stats_ok = bool(torch.equal(a[..., 0], b[..., 0])) and bool(torch.equal(a[..., 3:5], b[..., 3:5])) and \
            bool(torch.allclose(a[..., 1:3], b[..., 1:3], rtol=1e-6, atol=1e-9, equal_nan=True))
    flag = torch.tensor([int(same), int(stats_ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out[name] = {"labels_identical_on_all_ranks": bool(flag[0].item()), "stats_match_on_all_ranks": bool(flag[1].item()),
                 "mode": res.mode, "segments": res.n_labels, "sharded_ms_incl_first_call": round(dt * 1e3, 1)}
if rank == 0:
    print(json.dumps({"world": world, "raster": [H, W, 8], "checks": out}))
dist.destroy_process_group()

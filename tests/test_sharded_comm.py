"""CPU (gloo, world_size 2 and 3): the exchanges of the sharded path -- DistComm over torch.distributed
gives what LocalComm (all strips in one process) gives; plus the host-side label-offset logic."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _messages(world):
    """Deterministic per-rank payloads: band sums (int64), label halos (int32), table rows (float64)."""
    out = []
    for r in range(world):
        g = torch.Generator().manual_seed(100 + r)
        out.append({
            "acc_up": None if r == 0 else torch.randint(-5, 5, (6, 4), generator=g, dtype=torch.int64),
            "acc_down": None if r == world - 1 else torch.randint(-5, 5, (6, 4), generator=g, dtype=torch.int64),
            "halo_up": None if r == 0 else torch.randint(0, 99, (3, 7), generator=g, dtype=torch.int32),
            "halo_down": None if r == world - 1 else torch.randint(0, 99, (3, 7), generator=g, dtype=torch.int32),
            "rows_up": None if r == 0 else torch.rand((r + 1, 2, 8), generator=g, dtype=torch.float64),
            "kcore": [10 * (r + 1)], "mm": torch.tensor([float(r), float(-r)]),
        })
    return out


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from obia_b200.sharded import DistComm
    comm = DistComm(device="cpu")
    m = _messages(world)[rank]
    like = lambda t: None if t is None else (tuple(t.shape), t.dtype)
    # band exchange: same shapes on both sides of a boundary
    ru, rd = comm.neighbour_exchange([m["acc_up"]], [m["acc_down"]], [like(m["acc_up"])], [like(m["acc_down"])])
    hu, hd = comm.neighbour_exchange([m["halo_up"]], [m["halo_down"]], [like(m["halo_up"])], [like(m["halo_down"])])
    # table rows go up only; the receiver learns the size from the all-gathered counts
    counts = comm.all_gather_host([[0 if m["rows_up"] is None else m["rows_up"].shape[0]]])
    nb = counts[rank + 1][0] if rank + 1 < world else 0
    _, rows = comm.neighbour_exchange([m["rows_up"]], [None], [None], [((nb, 2, 8), torch.float64) if nb else None])
    kc = comm.all_gather_host([m["kcore"]])
    mm = m["mm"].clone()
    comm.all_reduce([mm], "max")
    gathered = comm.all_gather([m["mm"]])[0]
    ret[rank] = dict(ru=ru[0], rd=rd[0], hu=hu[0], hd=hd[0], rows=rows[0], kc=kc, mm=mm, gathered=torch.stack(gathered))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distcomm_matches_localcomm_over_gloo(world):
    from obia_b200.sharded import LocalComm
    msgs = _messages(world)
    local = LocalComm(world)
    ru, rd = local.neighbour_exchange([m["acc_up"] for m in msgs], [m["acc_down"] for m in msgs], None, None)
    hu, hd = local.neighbour_exchange([m["halo_up"] for m in msgs], [m["halo_down"] for m in msgs], None, None)
    _, rows = local.neighbour_exchange([m["rows_up"] for m in msgs], [None] * world, None, None)
    mm = [m["mm"].clone() for m in msgs]
    local.all_reduce(mm, "max")
    with mp.Manager() as man:
        ret = man.dict()
        mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        got = dict(ret)
    def same(a, b):
        return (a is None and b is None) or (a is not None and b is not None and torch.equal(a, b))
    for r in range(world):
        g = got[r]
        assert same(g["ru"], ru[r]) and same(g["rd"], rd[r]), "band exchange"
        assert same(g["hu"], hu[r]) and same(g["hd"], hd[r]), "halo exchange"
        assert same(g["rows"], rows[r]), "table rows"
        assert g["kc"] == [m["kcore"] for m in msgs]
        assert torch.equal(g["mm"], mm[r])
        assert torch.equal(g["gathered"], torch.stack([m["mm"] for m in msgs]))
    # what a rank receives from above is what the rank above sent down, and vice versa
    for r in range(1, world):
        assert torch.equal(got[r]["ru"], msgs[r - 1]["acc_down"]) and torch.equal(got[r - 1]["rd"], msgs[r]["acc_up"])


def test_label_offsets_are_the_exclusive_scan_of_core_counts():
    """The raster-order numbering across strips: rank r numbers its core pieces from
    start_label + sum(kcore[:r]); pieces that start in its upper halo get the labels just below."""
    kcore, kbefore, start = [7, 5, 9], [0, 3, 2], 1
    prefix = np.concatenate([[0], np.cumsum(kcore)])
    ranges = [(start + prefix[r] - kbefore[r], start + prefix[r] + kcore[r]) for r in range(3)]
    assert ranges == [(1, 8), (5, 13), (11, 22)]
    # rows a rank sends up are the LAST kbefore[r] labels of the rank above
    for r in (1, 2):
        assert ranges[r][0] + kbefore[r] == ranges[r - 1][1]

"""Global SLIC + zonal statistics on ONE raster sharded by row strips across GPUs
("global SLIC on one huge raster").

Every rank keeps rows [row0, row0 + h) of the raw raster in HBM.  Per sweep each rank assigns its
own pixels against the full (replicated) centre table and accumulates its contribution to the
centre sums; the int64 fixed-point sums are all-reduced over NCCL (integer addition: exact and
order-independent), so every rank derives the same centres and the labels are BIT-IDENTICAL to the
single-GPU run for any number of ranks.  Connectivity needs the whole label raster: the int32
strips are all-gathered (4 B/pixel) and every rank runs the exact connectivity kernels on the full
raster, keeping its strip.  Zonal statistics are computed per strip and merged with the pairwise
moment-combination formulas (Chan et al.); counts / min / max stay exact.

`ShardedSlic` holds one strip's state and exposes the steps separately so the same code path is
driven either by `torch.distributed` (`slic_zonal_distributed`) or, in the tests, by several strips
living on one GPU with the reductions done by hand.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib, pipeline, slic_host
from .pipeline import _i32_array, _p, _require_cuda, _stream_ptr


class ShardedSlic:
    def __init__(self, raw_strip, row0, H_total, segmentation_bands=None, *, n_segments=100, compactness=10.0,
                 max_num_iter=10, sigma=0, convert2lab=None, enforce_connectivity=True, min_size_factor=0.5,
                 max_size_factor=3, slic_zero=False, start_label=1, mask=None, spacing=None, exact=False):
        self.exact = bool(exact)
        _require_cuda(raw_strip, "raw_strip", torch.float32)
        if mask is not None:
            raise NotImplementedError("sharded global SLIC supports unmasked rasters (use the tiled driver for masks)")
        if np.any(np.asarray(sigma) > 0):
            raise NotImplementedError("sigma > 0 needs a halo exchange of the features: not implemented for strips")
        if slic_zero or spacing is not None:
            raise NotImplementedError("slic_zero / spacing are not implemented on the B200 path")
        if start_label not in (0, 1):
            raise ValueError("start_label should be 0 or 1.")
        self.lib = _lib.load()
        self.raw = pipeline._aligned(raw_strip)
        self.h, self.W, self.C = (int(s) for s in raw_strip.shape)
        self.row0, self.H = int(row0), int(H_total)
        self.bands = list(range(self.C)) if segmentation_bands is None else [int(b) for b in segmentation_bands]
        for band in self.bands:
            if band >= self.C or band < 0:
                raise IndexError(f"Band index {band} out of range. Available bands indices: 0 to {self.C - 1}.")
        self.n_segments, self.compactness, self.max_num_iter = n_segments, compactness, int(max_num_iter)
        self.start_label = int(start_label)
        self.enforce = bool(enforce_connectivity)
        self.min_size_factor, self.max_size_factor = min_size_factor, max_size_factor
        Cs = len(self.bands)
        if convert2lab and Cs != 3:
            raise ValueError("Lab colorspace conversion requires a RGB image.")
        self.to_lab = Cs == 3 and (convert2lab or convert2lab is None)
        self.Cf = 3 if self.to_lab else Cs
        self.dev = raw_strip.device

    # -- step 1: band ranges of this strip, to be min/max-reduced over the strips ----------------
    def local_minmax(self):
        mm, fl = pipeline.band_minmax(self.raw)
        return mm, fl          # (C, 4) float32: min, max, ., . ; (C,) int32 flags

    # -- step 2: features + replicated centres -----------------------------------------------------
    def prepare(self, minmax_global, flags_global):
        mm = minmax_global.cpu().numpy()
        fl = flags_global.cpu().numpy()
        f32 = np.float32
        for b in self.bands:
            if fl[b] or not np.isfinite(mm[b, 0]) or not np.isfinite(mm[b, 1]) or mm[b, 1] == mm[b, 0]:
                raise ValueError("unmasked NaN values in image are not supported")
        yx, steps = slic_host.grid_centroids(self.H, self.W, self.n_segments)
        self.n = int(yx.shape[0])
        self.step = float(max(steps))
        self.step_y, self.step_x = slic_host.window_steps(self.H, self.W, self.n)
        ratio = f32(1.0 / self.compactness)
        self.pitch = (self.W + 31) // 32 * 32
        self.feats = torch.empty((self.Cf, self.h, self.pitch), dtype=torch.float32, device=self.dev)
        bmin = np.ascontiguousarray(mm[:, 0], dtype=np.float32)
        bmax = np.ascontiguousarray(mm[:, 1], dtype=np.float32)
        _lib.check(self.lib.obia_b200_slic_features(
            _p(self.raw), self.h, self.W, self.C, _i32_array(self.bands), len(self.bands),
            bmin.ctypes.data_as(ctypes.c_void_p), bmax.ctypes.data_as(ctypes.c_void_p), 0.0, 1.0,
            int(self.to_lab), float(ratio), _p(self.feats), self.pitch, _stream_ptr()), "slic_features")
        c = np.zeros((self.n, 2 + self.Cf), dtype=np.float32)
        c[:, :2] = yx.astype(np.float32)
        self.centres = torch.from_numpy(c).to(self.dev)
        self.fix_scale = slic_host.fixed_point_scale(float(ratio) * (256.0 if self.to_lab else 4.0), self.H, self.W,
                                                     self.step_y, self.step_x)
        nbytes = self.lib.obia_b200_slic_workspace_bytes(self.H, self.W, self.Cf, self.n, self.step_y, self.step_x)
        self.ws = torch.empty((nbytes,), dtype=torch.uint8, device=self.dev)
        self.labels = torch.empty((self.h, self.W), dtype=torch.int32, device=self.dev)
        self.status = torch.zeros((4,), dtype=torch.int32, device=self.dev)
        _lib.check(self.lib.obia_b200_slic_begin(_p(self.labels), _p(self.ws), self.h, self.W, self.H, self.Cf, self.n,
                                                 self.step_y, self.step_x, self.start_label, _p(self.status),
                                                 _stream_ptr()), "slic_begin")

    # -- step 3 (x max_num_iter): sweep -> reduce acc over strips -> finish ----------------------------
    def sweep(self):
        sweep = self.lib.obia_b200_slic_sweep if self.exact else self.lib.obia_b200_slic_sweep_fast
        _lib.check(sweep(
            _p(self.feats), None, _p(self.centres), _p(self.labels), _p(self.ws), self.h, self.W, self.pitch, self.Cf,
            self.n, self.step, self.step_y, self.step_x, self.start_label, 0, 0, self.fix_scale, self.row0, self.H,
            _p(self.status), _stream_ptr()), "slic_sweep")

    def acc(self):
        """This strip's centre sums: int64 view (n, 3 + Cf) of the head of the workspace."""
        return self.ws[: self.n * (3 + self.Cf) * 8].view(torch.int64).view(self.n, 3 + self.Cf)

    def finish_sweep(self):
        _lib.check(self.lib.obia_b200_slic_finish_sweep(_p(self.centres), _p(self.ws), self.H, self.W, self.Cf, self.n,
                                                        self.step_y, self.step_x, self.fix_scale, _stream_ptr()),
                   "slic_finish_sweep")

    def check_status(self):
        if int(self.status[0].item()) != 0:
            raise _lib.ObiaB200Error("slic_sweep: a tile collected more than 1024 candidate centres")

    # -- step 4: connectivity on the gathered raster, keep the strip ------------------------------------
    def connect(self, full_labels):
        if not self.enforce:
            self.final = self.labels
            self.n_labels = self.n
            return
        seg = float(self.H * self.W) / self.n
        min_size, max_size = int(self.min_size_factor * seg), int(self.max_size_factor * seg)
        out, self.n_labels = pipeline.enforce_connectivity(full_labels, min_size, max_size, self.start_label)
        self.final = out[self.row0:self.row0 + self.h].contiguous()

    # -- step 5: per-strip statistics (merged by combine_stats) -----------------------------------------
    def strip_stats(self, bands=None):
        return pipeline.zonal_stats(self.final, self.raw, bands, max_label=self.n_labels + 1, resolution=1e-6)


def combine_stats(tables, resolution=1e-6):
    """Merge per-strip statistics tables (L, C, 8) into the statistics of the union of the strips.

    Fields: count, mean, variance, min, max, skewness, kurtosis, sum (pipeline.STAT_FIELDS).
    Central moments are recovered per strip and combined pairwise (Chan, Golub & LeVeque; Pebay).
    """
    def moments(t):
        n, mean, var = t[..., 0], t[..., 1], t[..., 2]
        var0 = torch.nan_to_num(var, nan=0.0)
        M2 = var0 * n
        skew = torch.nan_to_num(t[..., 5], nan=0.0)
        kurt = torch.nan_to_num(t[..., 6], nan=-3.0)
        M3 = skew * var0.pow(1.5) * n
        M4 = (kurt + 3.0) * var0 * var0 * n
        mean0 = torch.where(n > 0, mean, torch.zeros_like(mean))
        inf = torch.full_like(mean, float("inf"))
        return n, mean0, M2, M3, M4, torch.where(n > 0, t[..., 3], inf), torch.where(n > 0, t[..., 4], -inf)

    nA, mA, M2A, M3A, M4A, mnA, mxA = moments(tables[0])
    for t in tables[1:]:
        nB, mB, M2B, M3B, M4B, mnB, mxB = moments(t)
        n = nA + nB
        ns = torch.where(n > 0, n, torch.ones_like(n))
        d = mB - mA
        mean = mA + d * nB / ns
        M2 = M2A + M2B + d * d * nA * nB / ns
        M3 = M3A + M3B + d ** 3 * nA * nB * (nA - nB) / ns ** 2 + 3.0 * d * (nA * M2B - nB * M2A) / ns
        M4 = (M4A + M4B + d ** 4 * nA * nB * (nA * nA - nA * nB + nB * nB) / ns ** 3
              + 6.0 * d * d * (nA * nA * M2B + nB * nB * M2A) / ns ** 2 + 4.0 * d * (nA * M3B - nB * M3A) / ns)
        nA, mA, M2A, M3A, M4A = n, mean, M2, M3, M4
        mnA, mxA = torch.minimum(mnA, mnB), torch.maximum(mxA, mxB)
    n = nA
    ns = torch.where(n > 0, n, torch.ones_like(n))
    nan = torch.full_like(mA, float("nan"))
    var = M2A / ns
    degenerate = var <= (resolution * mA) ** 2
    skew = torch.where(degenerate, nan, (M3A / ns) / (var * var.sqrt()))
    kurt = torch.where(degenerate, nan, (M4A / ns) / (var * var) - 3.0)
    empty = n == 0
    out = torch.stack([n, torch.where(empty, nan, mA), torch.where(empty, nan, var), torch.where(empty, nan, mnA),
                       torch.where(empty, nan, mxA), torch.where(empty, nan, skew), torch.where(empty, nan, kurt),
                       torch.where(empty, torch.zeros_like(mA), mA * n)], dim=-1)
    return out


def split_rows(H_total, world):
    """Contiguous row strips, as equal as possible."""
    base, extra = divmod(H_total, world)
    rows, r = [], 0
    for k in range(world):
        h = base + (1 if k < extra else 0)
        rows.append((r, h))
        r += h
    return rows


def slic_zonal_distributed(raw_strip, row0, H_total, segmentation_bands=None, statistics_bands=None, **slic_kwargs):
    """Run the sharded path with torch.distributed (one process per GPU, NCCL).

    Every rank passes its own strip.  Returns (final labels of the strip, number of segments,
    statistics table (n_segments + 2, Cz, 8) of the WHOLE raster, identical on every rank).
    """
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    s = ShardedSlic(raw_strip, row0, H_total, segmentation_bands, **slic_kwargs)
    mm, fl = s.local_minmax()
    lo, hi = mm[:, 0].clone(), mm[:, 1].clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(fl, op=dist.ReduceOp.MAX)
    mm = torch.stack([lo, hi, lo, hi], dim=1)
    s.prepare(mm, fl)
    for _ in range(s.max_num_iter):
        s.sweep()
        dist.all_reduce(s.acc(), op=dist.ReduceOp.SUM)      # the one exchange per iteration (int64: exact)
        s.finish_sweep()
    s.check_status()
    rows = split_rows(H_total, world)
    if (row0, s.h) != rows[rank]:
        raise ValueError("strips must be the contiguous equal split of split_rows()")
    hmax = max(h for _, h in rows)
    pad = torch.full((hmax, s.W), s.start_label - 1, dtype=torch.int32, device=s.dev)
    pad[:s.h] = s.labels
    gathered = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(gathered, pad)
    full = torch.cat([g[:h] for g, (_, h) in zip(gathered, rows)], dim=0).contiguous()
    s.connect(full)
    st = s.strip_stats(statistics_bands)
    tables = [torch.empty_like(st) for _ in range(world)]
    dist.all_gather(tables, st.contiguous())
    return s.final, s.n_labels, combine_stats(tables)

"""Global SLIC + zonal statistics on ONE raster sharded by row strips across GPUs
("global SLIC on one huge raster"; the reference itself is single-process).

Every rank keeps rows [row0, row0 + h) of the raw raster in HBM; strips start on multiples of
`ROW_ALIGN` rows so that the CTA tiles of the assignment kernel are the same as in a single-GPU run.
Exchanges (NCCL over NVLink via torch.distributed, or tensor copies when several strips live on one
GPU in the tests):

  band ranges      one all-reduce (min / max) of C floats.
  per sweep        every rank assigns its own pixels against the replicated centre table and sums its
                   contribution in 64-bit fixed point; only the BANDS of the table whose centres can
                   receive pixels from two ranks (initial grid row within `band_steps` grid steps of a
                   strip boundary) are exchanged with the neighbour and added (integer addition: exact
                   and order-independent), so both neighbours derive identical centres.  A centre that
                   drifts out of its band raises a flag (`obia_b200_slic_band_check`) and the run is
                   repeated with the full all-reduce of the table.
  connectivity     SLIC components are bounded by the +-2*step windows, so each rank labels its strip
                   plus `halo` rows of its neighbours' labels (`obia_b200_connectivity_strip_*`); kept
                   pieces are counted per rank, one all-gather + exclusive prefix sum gives the label
                   offsets (north_star: "exclusive scan of per-tile label offsets").  A strip whose
                   result could depend on pixels outside the halo reports it; all ranks then fall back
                   to gathering the label raster and running the single-raster kernel.
  statistics       the raster-order numbering makes a rank's labels a contiguous range; statistics are
                   computed over that range only (`obia_b200_zonal_stats_range`) and the rows of the
                   segments that straddle a boundary are merged with the neighbour (pairwise moment
                   combination; counts / min / max exact).  The table stays sharded by label range.

Labels are BIT-IDENTICAL to the single-GPU run for any number of ranks (tests emulate 2-4 ranks on
one GPU; scripts/sharded_check.py runs over NCCL).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib, pipeline, slic_host
from .pipeline import _i32_array, _p, _require_cuda, _stream_ptr

ROW_ALIGN = 128     # strips start on multiples of the tallest CTA tile of the assignment kernels


def split_rows(H_total, world, align=ROW_ALIGN):
    """Contiguous row strips, as equal as possible, every strip starting on a multiple of `align`."""
    units = -(-H_total // align)
    base, extra = divmod(units, world)
    rows, r = [], 0
    for k in range(world):
        h = (base + (1 if k < extra else 0)) * align
        h = max(0, min(h, H_total - r))
        rows.append((r, h))
        r += h
    if any(h == 0 for _, h in rows):
        raise ValueError(f"raster of {H_total} rows is too small for {world} strips of {align}-row tiles")
    return rows


# ------------------------------------------------------------------ communication ---
class DistComm:
    """torch.distributed (NCCL on GPUs): this process holds one strip."""

    def __init__(self, device=None):
        import torch.distributed as dist
        self.dist = dist
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.local = [self.rank]
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())   # (gloo tests pass "cpu")
        self.device = torch.device(device)

    def all_reduce(self, tensors, op):
        ops = {"min": self.dist.ReduceOp.MIN, "max": self.dist.ReduceOp.MAX, "sum": self.dist.ReduceOp.SUM}
        self.dist.all_reduce(tensors[0], op=ops[op])

    def all_gather_host(self, values):
        """values: one list of python ints per local strip -> list over all ranks."""
        t = torch.tensor(values[0], dtype=torch.int64, device=self.device)
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [o.tolist() for o in torch.stack(out).cpu()]

    def all_gather(self, tensors):
        out = [torch.empty_like(tensors[0]) for _ in range(self.world)]
        self.dist.all_gather(out, tensors[0].contiguous())
        return [out]

    def neighbour_exchange(self, send_up, send_down, recv_up_like, recv_down_like):
        """send_up[i] goes to rank-1, send_down[i] to rank+1; returns what the upper / lower neighbour
        sent here (None at the raster edges).  `*_like`: (shape, dtype) of the expected messages."""
        dist, r = self.dist, self.rank
        ops, ru, rd = [], None, None
        dev = self.device
        if r > 0 and recv_up_like[0] is not None:
            ru = torch.empty(recv_up_like[0][0], dtype=recv_up_like[0][1], device=dev)
            ops.append(dist.P2POp(dist.irecv, ru, r - 1))
        if r < self.world - 1 and recv_down_like[0] is not None:
            rd = torch.empty(recv_down_like[0][0], dtype=recv_down_like[0][1], device=dev)
            ops.append(dist.P2POp(dist.irecv, rd, r + 1))
        if r > 0 and send_up[0] is not None:
            ops.append(dist.P2POp(dist.isend, send_up[0].contiguous(), r - 1))
        if r < self.world - 1 and send_down[0] is not None:
            ops.append(dist.P2POp(dist.isend, send_down[0].contiguous(), r + 1))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return [ru], [rd]


class LocalComm:
    """All strips live in this process (tests: several ranks emulated on one GPU)."""

    def __init__(self, world):
        self.world, self.rank = world, 0
        self.local = list(range(world))

    def all_reduce(self, tensors, op):
        st = torch.stack(tensors)
        red = {"min": st.amin(0), "max": st.amax(0), "sum": st.sum(0)}[op]
        for t in tensors:
            t.copy_(red)

    def all_gather_host(self, values):
        return [list(v) for v in values]

    def all_gather(self, tensors):
        return [[t.clone() for t in tensors] for _ in tensors]

    def neighbour_exchange(self, send_up, send_down, recv_up_like, recv_down_like):
        n = self.world
        ru = [None if i == 0 or send_down[i - 1] is None else send_down[i - 1].clone() for i in range(n)]
        rd = [None if i == n - 1 or send_up[i + 1] is None else send_up[i + 1].clone() for i in range(n)]
        return ru, rd


# ------------------------------------------------------------------ one strip ---
class ShardedSlic:
    def __init__(self, raw_strip, row0, H_total, segmentation_bands=None, *, n_segments=100, compactness=10.0,
                 max_num_iter=10, sigma=0, convert2lab=None, enforce_connectivity=True, min_size_factor=0.5,
                 max_size_factor=3, slic_zero=False, start_label=1, mask=None, spacing=None, exact=False):
        self.exact = bool(exact)
        _require_cuda(raw_strip, "raw_strip", torch.float32)
        self.mask = None
        if mask is not None:
            m = mask if isinstance(mask, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(mask)))
            self.mask = (m != 0).to(device=raw_strip.device, dtype=torch.uint8).contiguous()
            if tuple(self.mask.shape) != tuple(raw_strip.shape[:2]):
                raise ValueError("image and mask should have the same shape.")
        sig = np.ravel(np.asarray(sigma, dtype=np.float32))
        self.sigma_y, self.sigma_x = (float(sig[0]), float(sig[0])) if sig.size == 1 else (float(sig[-2]), float(sig[-1]))
        self.smooth = self.sigma_y > 0 or self.sigma_x > 0
        if spacing is not None:
            raise NotImplementedError("anisotropic spacing is not implemented on the B200 path")
        self.slic_zero = bool(slic_zero)
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
        self.top_open, self.bottom_open = self.row0 > 0, self.row0 + self.h < self.H

    # -- step 1: band ranges of this strip, to be min/max-reduced over the strips ----------------
    def local_minmax(self):
        mm, fl = pipeline.band_minmax(self.raw, self.mask)
        return mm, fl          # (C, 4) float32: min, max, masked min, masked max ; (C,) int32 flags

    # -- step 2: features + replicated centres -----------------------------------------------------
    def prepare(self, minmax_global, flags_global, band_steps=4, mask_init=None):
        """`mask_init` = (yx (n, 2) float64, steps (3,), number of mask pixels) of the WHOLE mask
        (pipeline.mask_centroids_device on the gathered mask: identical on every rank)."""
        mm = minmax_global.cpu().numpy()
        fl = flags_global.cpu().numpy()
        f32 = np.float32
        imin, imax = f32(0.0), f32(1.0)
        with np.errstate(invalid="ignore", divide="ignore"):
            nmin, nmax = [], []
            for b in self.bands:
                if fl[b] or not np.isfinite(mm[b, 0]) or not np.isfinite(mm[b, 1]) or mm[b, 1] == mm[b, 0]:
                    raise ValueError("unmasked NaN values in image are not supported")
                if self.mask is not None:
                    # skimage rescales with the range of the MASKED pixels of the normalised bands
                    d = f32(f32(mm[b, 1]) - f32(mm[b, 0]))
                    nmin.append(f32(f32(f32(mm[b, 2]) - f32(mm[b, 0])) / d))
                    nmax.append(f32(f32(f32(mm[b, 3]) - f32(mm[b, 0])) / d))
            if self.mask is not None:
                imin, imax = f32(min(nmin)), f32(max(nmax))
        if self.mask is not None:
            yx, steps, self.n_mask = mask_init
            self.n = int(yx.shape[0])
            self.ny = self.nx = 0
            self.step = float(max(steps))
            self.step_y, self.step_x = slic_host.window_steps(self.H, self.W, self.n)
            c0 = np.zeros((self.n, 2 + self.Cf), dtype=np.float32)
            c0[:, :2] = yx.astype(np.float32)
            self.centres = torch.from_numpy(c0).to(self.dev)
            starts, sy = (0, 0, 0), 1
        else:
            starts, isteps = slic_host.regular_grid_steps((1, self.H, self.W), self.n_segments)
            sy, sx = isteps[1] or 1, isteps[2] or 1
            ys = torch.arange(starts[1], self.H, sy, device=self.dev, dtype=torch.float32)
            xs = torch.arange(starts[2], self.W, sx, device=self.dev, dtype=torch.float32)
            self.ny, self.nx = int(ys.numel()), int(xs.numel())
            self.n = self.ny * self.nx
            self.step = float(max(1.0 if s is None else float(s) for s in isteps))
            self.step_y, self.step_x = slic_host.window_steps(self.H, self.W, self.n)
            self.centres = torch.zeros((self.n, 2 + self.Cf), dtype=torch.float32, device=self.dev)
            self.centres[:, 0] = ys.repeat_interleave(self.nx)
            self.centres[:, 1] = xs.repeat(self.ny)
        ratio = f32(1.0 / self.compactness)
        self.pitch = (self.W + 31) // 32 * 32
        self.feats = torch.empty((self.Cf, self.h, self.pitch), dtype=torch.float32, device=self.dev)
        bmin = np.ascontiguousarray(mm[:, 0], dtype=np.float32)
        bmax = np.ascontiguousarray(mm[:, 1], dtype=np.float32)
        # with sigma > 0 the features are smoothed first and scaled by 1/compactness afterwards (slic()'s
        # order); the driver exchanges the rows the Gaussian reaches across the strip boundaries
        _lib.check(self.lib.obia_b200_slic_features(
            _p(self.raw), self.h, self.W, self.C, _i32_array(self.bands), len(self.bands),
            bmin.ctypes.data_as(ctypes.c_void_p), bmax.ctypes.data_as(ctypes.c_void_p), float(imin), float(imax),
            int(self.to_lab), 1.0 if self.smooth else float(ratio), _p(self.feats), self.pitch, _stream_ptr()),
            "slic_features")
        self.ratio = float(ratio)
        self.fix_scale = slic_host.fixed_point_scale(float(ratio) * (256.0 if self.to_lab else 4.0), self.H, self.W,
                                                     self.step_y, self.step_x)
        nbytes = self.lib.obia_b200_slic_workspace_bytes(self.H, self.W, self.Cf, self.n, self.step_y, self.step_x)
        self.ws = torch.empty((nbytes,), dtype=torch.uint8, device=self.dev)
        self._maxdc_tail = (self.n * 4 + 255) // 256 * 256      # workspace layout: ..., maxdc [n] float32 (256-byte blocks)
        self.labels = torch.empty((self.h, self.W), dtype=torch.int32, device=self.dev)
        self.status = torch.zeros((4,), dtype=torch.int32, device=self.dev)
        self.begin()
        # centre-index bands at the strip boundaries: grid rows within `band_steps` steps of the boundary
        # (masked runs place their centres by k-means, not on a grid: they exchange the whole table)
        def band(yb):
            if self.mask is not None:
                return (0, 0)
            lo = int(np.ceil((yb - band_steps * sy - starts[1]) / sy))
            hi = int(np.floor((yb + band_steps * sy - starts[1]) / sy))
            lo, hi = max(lo, 0), min(hi, self.ny - 1)
            return (lo * self.nx, (hi + 1) * self.nx) if hi >= lo else (0, 0)
        self.band_up = band(self.row0) if self.top_open else (0, 0)
        self.band_down = band(self.row0 + self.h) if self.bottom_open else (0, 0)

    def begin(self):
        """Start of a run of sweeps (`_slic_cython` entry): labels = mask label, sums and maxima reset."""
        _lib.check(self.lib.obia_b200_slic_begin(_p(self.labels), _p(self.ws), self.h, self.W, self.H, self.Cf, self.n,
                                                 self.step_y, self.step_x, self.start_label, _p(self.status),
                                                 _stream_ptr()), "slic_begin")

    # -- step 2b (sigma > 0): Gaussian over the strip extended by the neighbours' feature rows -----------
    def smooth_radius(self):
        return int(4.0 * self.sigma_y + 0.5) if self.sigma_y > 0 else 0

    def feature_rows(self, top):
        """The `smooth_radius()` unsmoothed feature rows next to the upper / lower strip boundary."""
        r = self.smooth_radius()
        return (self.feats[:, :r] if top else self.feats[:, self.h - r:]).contiguous()

    def smooth_features(self, rows_up, rows_down):
        """scipy's reflecting Gaussian on [neighbour rows, strip, neighbour rows]: the reflection only
        reaches the rows that are cropped away again, except at the true raster edges."""
        wy, ry = slic_host.gaussian_taps(self.sigma_y) if self.sigma_y > 0 else (np.ones(1), 0)
        wx, rx = slic_host.gaussian_taps(self.sigma_x) if self.sigma_x > 0 else (np.ones(1), 0)
        parts = [p for p in (rows_up, self.feats, rows_down) if p is not None]
        ext = torch.cat(parts, dim=1).contiguous() if len(parts) > 1 else self.feats
        He = int(ext.shape[1])
        out = torch.empty_like(ext)
        tmp = torch.empty_like(ext) if max(ry, rx) > 63 else torch.empty((4,), dtype=torch.float32, device=self.dev)
        _lib.check(self.lib.obia_b200_gaussian_planar(
            _p(ext), _p(tmp), _p(out), He, self.W, self.pitch, self.Cf, wy.ctypes.data_as(ctypes.c_void_p), ry,
            wx.ctypes.data_as(ctypes.c_void_p), rx, self.ratio, _stream_ptr()), "gaussian_planar")
        top = 0 if rows_up is None else int(rows_up.shape[1])
        self.feats = out[:, top:top + self.h].contiguous()

    # -- step 3 (x max_num_iter): sweep -> combine acc over strips -> finish ---------------------------
    def sweep(self, ignore_color=False):
        sweep = self.lib.obia_b200_slic_sweep if self.exact else self.lib.obia_b200_slic_sweep_fast
        _lib.check(sweep(
            _p(self.feats), _p(self.mask), _p(self.centres), _p(self.labels), _p(self.ws), self.h, self.W, self.pitch,
            self.Cf, self.n, self.step, self.step_y, self.step_x, self.start_label, int(ignore_color),
            int(self.slic_zero), self.fix_scale, self.row0, self.H, _p(self.status), _stream_ptr()), "slic_sweep")

    def maxdc(self):
        """SLICO: the per-centre running maxima of the colour distance (float32 view (n,), the last table
        of the workspace); the strips' tables are combined with an element-wise maximum."""
        return self.ws[self.ws.numel() - self._maxdc_tail:][: self.n * 4].view(torch.float32)

    def update_max_color(self):
        _lib.check(self.lib.obia_b200_slic_update_max_color(
            _p(self.feats), _p(self.mask), _p(self.centres), _p(self.labels), _p(self.ws), self.h, self.W, self.H, self.pitch,
            self.Cf, self.n, self.step_y, self.step_x, self.start_label, _stream_ptr()), "slic_update_max_color")

    def acc(self):
        """This strip's centre sums: int64 view (n, 3 + Cf) of the head of the workspace."""
        return self.ws[: self.n * (3 + self.Cf) * 8].view(torch.int64).view(self.n, 3 + self.Cf)

    def finish_sweep(self, check_bands=False):
        _lib.check(self.lib.obia_b200_slic_finish_sweep(_p(self.centres), _p(self.ws), self.H, self.W, self.Cf, self.n,
                                                        self.step_y, self.step_x, self.fix_scale, _stream_ptr()),
                   "slic_finish_sweep")
        if check_bands:
            _lib.check(self.lib.obia_b200_slic_band_check(
                _p(self.centres), self.n, self.Cf, self.row0, self.row0 + self.h, self.H, self.band_up[0],
                self.band_up[1], self.band_down[0], self.band_down[1], self.step_y, _p(self.status), _stream_ptr()),
                "slic_band_check")

    def sizes(self):
        seg = float(self.n_mask if self.mask is not None else self.H * self.W) / self.n
        return int(self.min_size_factor * seg), int(self.max_size_factor * seg)

    def default_halo(self):
        """Rows of neighbour labels on each side: three window heights (a SLIC component spans at most
        4*step+1 rows).  The outer third is booked unknown by the kernel, the inner two thirds hold the
        pieces that reach the core and the pieces they can merge into."""
        return 12 * self.step_y + 12

    # -- step 4a: connectivity on the gathered raster (fallback) ----------------------------------------
    def connect_full(self, full_labels):
        min_size, max_size = self.sizes()
        out, n_labels = pipeline.enforce_connectivity(full_labels, min_size, max_size, self.start_label)
        self.final = out[self.row0:self.row0 + self.h].contiguous()
        self.n_labels = n_labels
        self.has_zero = None

    # -- step 4b: connectivity on the strip + halo ---------------------------------------------------
    def strip_begin(self, halo_up, halo_down):
        parts = [t for t in (halo_up, self.labels, halo_down) if t is not None]
        self.ext = torch.cat(parts, dim=0).contiguous() if len(parts) > 1 else self.labels
        self.core_row0 = 0 if halo_up is None else int(halo_up.shape[0])
        H_ext = int(self.ext.shape[0])
        # (kept per strip: the workspace lives from strip_begin to strip_finish, and strips emulated in one process
        #  must not share it)
        self.cc_ws = pipeline.scratch(self.dev, self.lib.obia_b200_connectivity_workspace_bytes(H_ext, self.W),
                                      ("connectivity-strip", self.row0))
        counts = (ctypes.c_int64 * 5)()
        min_size, max_size = self.sizes()
        _lib.check(self.lib.obia_b200_connectivity_strip_begin(
            _p(self.ext), _p(self.cc_ws), H_ext, self.W, self.core_row0, self.h, int(halo_up is not None),
            int(halo_down is not None), min_size, max_size, self.start_label, counts, _stream_ptr()),
            "connectivity_strip_begin")
        self.k_before, self.k_core, self.cc_rounds = int(counts[0]), int(counts[1]), int(counts[3])
        self.k_shared = int(counts[4])     # rows of the statistics table shared with the rank above
        return self.k_core

    def strip_finish(self, labels_before):
        """labels_before = number of kept pieces that start in the core rows of the ranks above."""
        self.label_base = int(labels_before)
        out = torch.empty((self.h, self.W), dtype=torch.int32, device=self.dev)
        flags = (ctypes.c_int32 * 2)()
        min_size, max_size = self.sizes()
        _lib.check(self.lib.obia_b200_connectivity_strip_finish(
            _p(self.ext), _p(out), _p(self.cc_ws), int(self.ext.shape[0]), self.W, self.core_row0, self.h, min_size,
            max_size, self.start_label, self.label_base - self.k_before, flags, _stream_ptr()),
            "connectivity_strip_finish")
        self.final = out
        self.has_zero = bool(flags[1])
        del self.cc_ws, self.ext
        return int(flags[0])

    # -- step 5: statistics over this rank's label range -------------------------------------------------
    def range_stats(self, bands=None, resolution=1e-6, with_zero=False):
        """Rows = labels [start_label + label_base - k_shared, start_label + label_base + k_core); with
        `with_zero` (label 0 outside that range) one extra leading row for label 0, returned separately.
        Returns (table, zero row or None)."""
        bands = list(range(self.C)) if bands is None else [int(b) for b in bands]
        lo = self.start_label + self.label_base - self.k_shared
        zero = 1 if (with_zero and lo > 0) else 0
        n_rows = max(1, self.k_shared + self.k_core) + zero
        stats = torch.empty((n_rows, len(bands), 8), dtype=torch.float64, device=self.dev)
        ws = torch.empty((self.lib.obia_b200_zonal_workspace_bytes(n_rows - 1, 8),), dtype=torch.uint8, device=self.dev)
        _lib.check(self.lib.obia_b200_zonal_stats_range(_p(self.final), _p(self.raw), self.h, self.W, self.C,
                                                        _i32_array(bands), len(bands), lo, n_rows, zero,
                                                        float(resolution), _p(stats), _p(ws), _stream_ptr()),
                   "zonal_stats_range")
        if zero:
            return stats[1:], stats[:1]
        if with_zero:      # label 0 lies inside the rank's own range (first rank): nothing extra to report
            empty = torch.full((1, len(bands), 8), float("nan"), dtype=torch.float64, device=self.dev)
            empty[..., 0] = 0.0
            empty[..., 7] = 0.0
            return stats, empty
        return stats, None


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


# ------------------------------------------------------------------ driver ---
class ShardedResult:
    """Per local strip: final labels (core rows), the rank's slice of the statistics table (rows =
    labels [label_lo, label_lo + rows)), the statistics of label 0 (or None), the global segment count."""

    def __init__(self):
        self.labels, self.stats, self.label_lo, self.zero_row = [], [], [], None
        self.n_labels, self.timings, self.mode = 0, {}, {}


def run_sharded(strips, comm, statistics_bands=None, exchange="band", halo=None, timings=False, stats=True):
    """Drive the strips held by this process (`comm.local`) through the sharded path.

    exchange: "band" (neighbour exchange of the boundary bands of the centre sums, falls back to
    "allreduce" when a centre leaves its band) or "allreduce" (whole table every sweep)."""
    res = ShardedResult()
    ev = []

    def mark(name):
        if timings:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev.append((name, e))

    mark("start")
    masked = strips[0].mask is not None
    mms = [s.local_minmax() for s in strips]
    lo = [m[0][:, 0].clone() for m in mms]
    hi = [m[0][:, 1].clone() for m in mms]
    mlo = [torch.nan_to_num(m[0][:, 2], nan=float("inf")) for m in mms]     # masked range (NaN: no mask pixel here)
    mhi = [torch.nan_to_num(m[0][:, 3], nan=float("-inf")) for m in mms]
    fl = [m[1] for m in mms]
    comm.all_reduce(lo, "min")
    comm.all_reduce(hi, "max")
    comm.all_reduce(fl, "max")
    mask_init = [None] * len(strips)
    if masked:
        comm.all_reduce(mlo, "min")
        comm.all_reduce(mhi, "max")
        # maskSLIC initialisation needs the whole mask (RandomState(123) sample of its pixels + k-means):
        # 1 byte per pixel is gathered and every rank derives the same centres
        rows_m = [h for _, h in split_rows(strips[0].H, comm.world)] if comm.world > 1 else [strips[0].h]
        hmax = max(rows_m)
        pads = []
        for s in strips:
            pad = torch.zeros((hmax, s.W), dtype=torch.uint8, device=s.dev)
            pad[:s.h] = s.mask
            pads.append(pad)
        gathered = comm.all_gather(pads)
        for i, (s, g) in enumerate(zip(strips, gathered)):
            full = torch.cat([t[:h] for t, h in zip(g, rows_m)], dim=0).contiguous()
            if i == 0 or len(strips) == 1:
                init = pipeline.mask_centroids_device(full, int(s.n_segments))
            mask_init[i] = init
            del full

    def prepare_all():
        for i, s in enumerate(strips):
            s.prepare(torch.stack([lo[i], hi[i], mlo[i], mhi[i]], dim=1), fl[i], mask_init=mask_init[i])
        if strips[0].smooth:
            r = strips[0].smooth_radius()
            if any(s.h < r for s in strips):
                raise ValueError("strips are thinner than the Gaussian radius")
            up = [s.feature_rows(True) if s.top_open and r > 0 else None for s in strips]
            down = [s.feature_rows(False) if s.bottom_open and r > 0 else None for s in strips]
            like = ((strips[0].Cf, r, strips[0].pitch), torch.float32)
            ru, rd = comm.neighbour_exchange(up, down, [like if u is not None else None for u in up],
                                             [like if d is not None else None for d in down])
            for s, a, b in zip(strips, ru, rd):
                s.smooth_features(a, b)

    prepare_all()
    mark("preprocess")

    def slic(mode, ignore_color=False):
        for it in range(strips[0].max_num_iter):
            for s in strips:
                s.sweep(ignore_color)
            if mode == "allreduce" or comm.world == 1:
                comm.all_reduce([s.acc() for s in strips], "sum")
            else:
                up = [s.acc()[s.band_up[0]:s.band_up[1]] if s.top_open else None for s in strips]
                down = [s.acc()[s.band_down[0]:s.band_down[1]] if s.bottom_open else None for s in strips]
                like_u = [None if u is None else (tuple(u.shape), u.dtype) for u in up]
                like_d = [None if d is None else (tuple(d.shape), d.dtype) for d in down]
                ru, rd = comm.neighbour_exchange(up, down, like_u, like_d)
                for s, u, d, a, b in zip(strips, up, down, ru, rd):
                    if a is not None:
                        u += a
                    if b is not None:
                        d += b
            for s in strips:
                s.finish_sweep(check_bands=(mode == "band" and comm.world > 1))
            if strips[0].slic_zero:
                # SLICO: every strip raises the maxima of the centres its pixels belong to; the band
                # centres take the larger of the two ranks' values (max is exact and order-free)
                for s in strips:
                    s.update_max_color()
                if comm.world > 1 and mode == "allreduce":
                    comm.all_reduce([s.maxdc() for s in strips], "max")
                elif comm.world > 1:
                    up = [s.maxdc()[s.band_up[0]:s.band_up[1]] if s.top_open else None for s in strips]
                    down = [s.maxdc()[s.band_down[0]:s.band_down[1]] if s.bottom_open else None for s in strips]
                    like_u = [None if u is None else (tuple(u.shape), u.dtype) for u in up]
                    like_d = [None if d is None else (tuple(d.shape), d.dtype) for d in down]
                    ru, rd = comm.neighbour_exchange(up, down, like_u, like_d)
                    for u, d, a, b in zip(up, down, ru, rd):
                        if a is not None:
                            torch.maximum(u, a, out=u)
                        if b is not None:
                            torch.maximum(d, b, out=d)
        st = [s.status.clone() for s in strips]
        comm.all_reduce(st, "max")
        flags = st[0].cpu().tolist()
        if flags[0]:
            raise _lib.ObiaB200Error("slic_sweep: a tile collected more than 1024 candidate centres")
        return flags[1]

    mode = "allreduce" if masked else exchange
    if masked:
        slic(mode, ignore_color=True)     # maskSLIC step 2: spatial-only k-means moves the centres first
        for s in strips:
            s.begin()
    if slic(mode):
        # a centre left its band: redo with the whole table (prepare() resets centres, labels, sums)
        mode = "allreduce"
        prepare_all()
        slic(mode)
    res.mode["exchange"] = mode
    mark("slic")

    s0 = strips[0]
    if not s0.enforce:
        for s in strips:
            s.final, s.k_before, s.k_shared, s.k_core, s.label_base, s.has_zero = s.labels, 0, 0, s.n, 0, False
        res.n_labels = s0.n
        res.mode["connectivity"] = "none"
        kcores, kshared, any_zero = None, None, 0
    else:
        h_rows = int(halo) if halo is not None else s0.default_halo()
        incomplete = 1
        rows_all = [[h] for _, h in split_rows(s0.H, comm.world)] if comm.world > 1 else [[s0.h]]
        if comm.world > 1 and min(r[0] for r in rows_all) >= h_rows:
            up = [s.labels[:h_rows] if s.top_open else None for s in strips]
            down = [s.labels[s.h - h_rows:] if s.bottom_open else None for s in strips]
            like = ((h_rows, s0.W), torch.int32)
            ru, rd = comm.neighbour_exchange(up, down, [like if s.top_open else None for s in strips],
                                             [like if s.bottom_open else None for s in strips])
            for s, a, b in zip(strips, ru, rd):
                s.strip_begin(a, b)
            counts = comm.all_gather_host([[s.k_core, s.k_shared] for s in strips])   # one exchange for both
            kcores, kshared = [c[0] for c in counts], [c[1] for c in counts]
            prefix = np.concatenate([[0], np.cumsum(kcores)])
            inc = [torch.tensor([s.strip_finish(int(prefix[r])), int(bool(s.has_zero))], dtype=torch.int32,
                                device=s.dev) for s, r in zip(strips, comm.local)]
            comm.all_reduce(inc, "max")
            inc_host = inc[0].cpu().tolist()
            incomplete, any_zero = int(inc_host[0]), int(inc_host[1])
            res.n_labels = int(prefix[-1])
            res.mode["connectivity"] = "strip+halo"
        elif comm.world == 1:
            s0.strip_begin(None, None)
            incomplete = s0.strip_finish(0)
            res.n_labels = s0.k_core
            kcores, kshared, any_zero = [s0.k_core], [0], int(bool(s0.has_zero))
            res.mode["connectivity"] = "strip+halo"
        if incomplete:
            # some strip's result could depend on pixels outside its halo (or the strips are thinner than
            # the halo): gather the label raster, every rank runs the single-raster kernel
            hmax = max(r[0] for r in rows_all)
            pads = []
            for s in strips:
                pad = torch.full((hmax, s.W), s.start_label - 1, dtype=torch.int32, device=s.dev)
                pad[:s.h] = s.labels
                pads.append(pad)
            gathered = comm.all_gather(pads)
            for s, g in zip(strips, gathered):
                full = torch.cat([t[:r[0]] for t, r in zip(g, rows_all)], dim=0).contiguous()
                s.connect_full(full)
            res.n_labels = s0.n_labels
            res.mode["connectivity"] = "gathered"
            kcores = None
    mark("connectivity")

    if masked:
        for s in strips:
            s.final.masked_fill_(s.mask == 0, -1)      # segment_boundaries.py:55-57
    res.labels = [s.final for s in strips]
    if stats:
        if kcores is None:
            # labels are not a contiguous range per rank: whole-table statistics merged over all ranks
            tabs = [pipeline.zonal_stats(s.final, s.raw, statistics_bands, max_label=res.n_labels + 1) for s in strips]
            allt = comm.all_gather(tabs)
            merged = [combine_stats(list(g)) for g in allt]
            res.stats, res.label_lo = merged, [0 for _ in strips]
            res.mode["stats"] = "replicated"
        else:
            both = [s.range_stats(statistics_bands, with_zero=bool(any_zero)) for s in strips]
            tabs = [b[0] for b in both]
            kb = [[k] for k in kshared]
            # rows of pieces that start above the core go UP to their owner
            up = [t[:s.k_shared] if s.top_open and s.k_shared > 0 else None for s, t in zip(strips, tabs)]
            Cz = int(tabs[0].shape[1])
            like_d = []
            for s, r in zip(strips, comm.local):
                nb = kb[r + 1][0] if r + 1 < comm.world else 0
                like_d.append(((nb, Cz, 8), torch.float64) if nb > 0 else None)
            _, rd = comm.neighbour_exchange(up, [None] * len(strips), [None] * len(strips), like_d)
            out = []
            for s, t, b in zip(strips, tabs, rd):
                own = t[s.k_shared:s.k_shared + s.k_core]
                if b is not None and b.shape[0] > 0:
                    nb = int(b.shape[0])
                    if nb > own.shape[0]:
                        raise _lib.ObiaB200Error("a segment spans more than two strips: strips are too thin")
                    tail = combine_stats([own[own.shape[0] - nb:], b])
                    own = torch.cat([own[:own.shape[0] - nb], tail], dim=0)
                out.append(own)
            res.stats = out
            res.label_lo = [s.start_label + s.label_base for s in strips]
            res.mode["stats"] = "label-range"
            if any_zero:
                zero = combine_stats(list(comm.all_gather([b[1] for b in both])[0]))
                if s0.start_label == 1:
                    res.zero_row = zero
                else:
                    # start_label 0: label 0 is also the first kept segment, owned by the first rank
                    for s, t in zip(strips, res.stats):
                        if s.row0 == 0 and t.shape[0] > 0:
                            t[:1] = combine_stats([t[:1], zero])
    mark("stats")
    if timings:
        torch.cuda.synchronize()
        res.timings = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(ev[:-1], ev[1:])}
    return res


def slic_zonal_distributed(raw_strip, row0, H_total, segmentation_bands=None, statistics_bands=None,
                           exchange="band", halo=None, timings=False, stats=True, **slic_kwargs):
    """Run the sharded path with torch.distributed (one process per GPU, NCCL).

    Every rank passes its own strip (`split_rows(H_total, world)`).  Returns a ShardedResult whose
    lists hold this rank's single entry: final labels of the strip, the rank's rows of the statistics
    table and the label of its first row."""
    comm = DistComm()
    rows = split_rows(H_total, comm.world)
    if (int(row0), int(raw_strip.shape[0])) != rows[comm.rank]:
        raise ValueError("strips must be the aligned contiguous split of split_rows()")
    s = ShardedSlic(raw_strip, row0, H_total, segmentation_bands, **slic_kwargs)
    return run_sharded([s], comm, statistics_bands, exchange=exchange, halo=halo, timings=timings, stats=stats)

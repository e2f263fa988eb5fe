"""Window batches of the tiled driver: every window of a batch (all black tiles of a chunk, or all white
windows of one tile-row of obia/utils/tiling.py:103-287) goes through min/max, features, maskSLIC
initialisation, the SLIC sweeps and connectivity in ONE launch per stage (csrc/batch.cuh), instead of one
`create_segments(image=tile, mask=..., n_segments=...)` call per tile (tiling.py:137-143, :275-281).

The result inside every window is what `pipeline.slic_labels(window, mask=..., n_segments=...)` gives for
that window alone (tests/test_tiling_batched.py compares the two drivers pixel for pixel).
"""
from __future__ import annotations

import ctypes
import functools
import math

import numpy as np
import torch

from . import _lib, slic_host

# mirrors struct WinDesc (csrc/batch.cuh)
WIN_DESC = np.dtype([
    ("y0", "<i4"), ("x0", "<i4"), ("h", "<i4"), ("w", "<i4"),
    ("row0", "<i4"), ("valid", "<i4"), ("n", "<i4"), ("c0", "<i4"),
    ("cell0", "<i4"), ("ncy", "<i4"), ("ncx", "<i4"), ("step_y", "<i4"),
    ("step_x", "<i4"), ("min_size", "<i4"), ("max_size", "<i4"), ("n_mask", "<i4"),
    ("sw", "<f4"), ("inv_w", "<f4"), ("fix_scale32", "<f4"), ("imin", "<f4"),
    ("idiff", "<f4"), ("rescale", "<i4"), ("p0", "<i4"), ("m", "<i4"),
    ("fix_scale", "<f8"), ("fix_ratio", "<i8"), ("km_cs", "<f8"),
    ("km_ncy", "<i4"), ("km_ncx", "<i4"), ("km_cell0", "<i4"), ("pad", "<i4"),
])
assert WIN_DESC.itemsize == 136

MAX_BATCH = 4096                 # windows per batch (grid.y / grid.z stay far below 65535)
MAX_SLAB_PIXELS = 160_000_000    # slab pixels per batch (int32 positions; about 12 GB of scratch at 4 bands)

SUPPORTED_KWARGS = {"compactness", "max_num_iter", "convert2lab", "enforce_connectivity", "min_size_factor",
                    "max_size_factor", "start_label", "sigma", "slic_zero", "spacing", "channel_axis", "exact",
                    "segmentation_bands"}


def supports(slic_kwargs):
    """True when the batched path reproduces `slic_labels(**slic_kwargs)`: tolerance-mode kernel, no Gaussian
    pre-smoothing, no SLICO, connectivity enforced."""
    kw = slic_kwargs
    if set(kw) - SUPPORTED_KWARGS:
        return False
    sigma = kw.get("sigma", 0)
    if np.ndim(sigma) != 0 or float(sigma) != 0.0:
        return False
    if kw.get("slic_zero", False) or kw.get("exact", False) or not kw.get("enforce_connectivity", True):
        return False
    if kw.get("spacing") is not None or kw.get("channel_axis", -1) not in (-1, 2):
        return False
    return kw.get("start_label", 1) in (0, 1)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


TIMINGS = None      # set to a dict (section -> seconds) to collect synchronised per-stage times (debugging aid)


_T_LAST = [0.0]


def _tick(name):
    """Attribute the (synchronised) time since the previous tick to `name` when TIMINGS is a dict."""
    if TIMINGS is None:
        return
    import time
    torch.cuda.synchronize()
    now = time.perf_counter()
    if name is not None:
        TIMINGS[name] = TIMINGS.get(name, 0.0) + now - _T_LAST[0]
    _T_LAST[0] = now


@functools.lru_cache(maxsize=65536)
def _geometry(h, w, n):
    """(step_y, step_x) of the +-2*step windows for n centres on an h x w window."""
    return slic_host.window_steps(h, w, n)


@functools.lru_cache(maxsize=65536)
def _fix_scale(max_abs, h, w, step_y, step_x):
    return slic_host.fixed_point_scale(max_abs, h, w, step_y, step_x)


@functools.lru_cache(maxsize=4096)
def _grid_init(h, w, n_segments):
    """Regular-grid centres of an unmasked h x w window: (ys, xs, step) as pipeline.slic_labels lays them out."""
    starts, isteps = slic_host.regular_grid_steps((1, h, w), n_segments)
    ys = np.arange(starts[1], h, isteps[1] or 1, dtype=np.float32)
    xs = np.arange(starts[2], w, isteps[2] or 1, dtype=np.float32)
    steps = [1.0 if s is None else float(s) for s in isteps]
    return ys, xs, float(max(steps))


class WindowBatch:
    """One batch of disjoint windows of `raw` ((H, Wl, C) float32 on the device).

    windows: (B, 4) int array of (y0, x0, h, w), x0 in LOCAL columns of `raw`.
    After `segment(...)`: `labels` = (slab_rows, slab_w) int32 slab of final labels (-1 on masked pixels,
    kept pieces numbered start_label.. window by window), `usable` = (B,) bool, `n_labels`.
    """

    def __init__(self, raw, windows):
        self.lib = _lib.load()
        self.raw = raw
        self.dev = raw.device
        self.H, self.Wl, self.C = (int(s) for s in raw.shape)
        win = np.asarray(windows, dtype=np.int64).reshape(-1, 4)
        self.B = B = int(win.shape[0])
        if B == 0 or B > MAX_BATCH:
            raise ValueError("window batch size")
        self.hmax, self.wmax = int(win[:, 2].max()), int(win[:, 3].max())
        self.win_rows = self.hmax + 1                       # one masked row between windows
        self.slab_rows = B * self.win_rows
        self.slab_w = (self.wmax + 3) // 4 * 4
        self.pitch = (self.wmax + 31) // 32 * 32
        if self.slab_rows * self.slab_w >= 2 ** 31 - 1:
            raise ValueError("window batch slab exceeds int32 pixels")
        d = np.zeros(B, dtype=WIN_DESC)
        d["y0"], d["x0"], d["h"], d["w"] = win[:, 0], win[:, 1], win[:, 2], win[:, 3]
        d["row0"] = np.arange(B, dtype=np.int64) * self.win_rows
        self.desc = d
        self.desc_dev = None
        self.mask_slab = None
        self.upload()

    # ------------------------------------------------------------------ helpers
    def upload(self):
        host = torch.from_numpy(self.desc.view(np.uint8).reshape(-1))
        if self.desc_dev is None:
            self.desc_dev = host.to(self.dev)
        else:
            self.desc_dev.copy_(host)

    def new_mask_slab(self):
        self.mask_slab = torch.zeros((self.slab_rows, self.slab_w), dtype=torch.uint8, device=self.dev)
        return self.mask_slab

    def mask_from(self, mask):
        """Slab mask = the windows of a (H, Wl) uint8 / bool mask raster (black tiles with a user mask)."""
        m = mask if mask.dtype == torch.uint8 else mask.to(torch.uint8)
        m = m.contiguous()
        self.new_mask_slab()
        _lib.check(self.lib.obia_b200_window_mask_copy(_p(m), int(m.shape[1]), _p(self.desc_dev), self.B, self.hmax,
                                                       self.wmax, _p(self.mask_slab), self.slab_w, _stream()),
                   "window_mask_copy")

    # ------------------------------------------------------------------ the path
    def segment(self, **kw):
        return self.begin(**kw).finish()

    def begin(self, *, n_segments=None, pixel_area=1.0, crown_radius=5, segmentation_bands=None, compactness=10.0,
              max_num_iter=10, convert2lab=None, min_size_factor=0.5, max_size_factor=3, start_label=1, **_ignored):
        """First half: per-window band ranges and mask counts (one read-back), the per-window parameters, and the
        start of the host-side sample draws (background threads) -- nothing here waits for earlier device work
        that was queued after the caller's last synchronisation."""
        self._st = None
        lib, dev, B, C = self.lib, self.dev, self.B, self.C
        d = self.desc
        masked = self.mask_slab is not None
        bands = list(range(C)) if segmentation_bands is None else [int(b) for b in segmentation_bands]
        for band in bands:
            if band >= C or band < 0:
                raise IndexError(f"Band index {band} out of range. Available bands indices: 0 to {C - 1}.")
        Cs = len(bands)
        to_lab = False
        if convert2lab or convert2lab is None:
            if Cs != 3 and convert2lab:
                raise ValueError("Lab colorspace conversion requires a RGB image.")
            to_lab = Cs == 3
        Cf = 3 if to_lab else Cs
        f32 = np.float32

        _tick(None)
        # ---- K1a per window: band ranges, mask counts (one read-back) ---------------------------------
        stats = torch.empty((B, C, 4), dtype=torch.float32, device=dev)
        flags = torch.empty((B, C), dtype=torch.int32, device=dev)
        counts = torch.empty((B,), dtype=torch.int32, device=dev)
        _lib.check(lib.obia_b200_window_stats(_p(self.raw), self.Wl, C, _p(self.desc_dev), B, self.hmax,
                                              _p(self.mask_slab), self.slab_w, _p(stats), _p(flags), _p(counts),
                                              _stream()), "window_stats")
        packed = torch.cat([stats.reshape(-1), flags.reshape(-1).to(torch.float32),
                            counts.to(torch.float32)]).cpu().numpy()
        mm = packed[:B * C * 4].reshape(B, C, 4)
        fl = packed[B * C * 4:B * C * 5].reshape(B, C) != 0
        n_mask = packed[B * C * 5:].astype(np.int64)
        hw = d["h"].astype(np.int64) * d["w"].astype(np.int64)
        n_coord = n_mask if masked else hw

        _tick("stats + read-back")
        # ---- per-window parameters (what slic_labels derives on the host, vectorised) ------------------
        if n_segments is not None:
            n_seg = np.full(B, int(n_segments), dtype=np.int64)
        else:
            if not masked:
                raise ValueError("create_tiled_segments needs `input_mask` or an explicit n_segments")
            crown_area = math.pi * (crown_radius ** 2)
            n_seg = np.rint(n_mask.astype(np.float64) * pixel_area / crown_area).astype(np.int64)   # tiling.py:126-135
        valid = n_seg > 0
        sel = mm[:, bands, :]                                   # (B, Cs, 4)
        mn, mx, mmn, mmx = (sel[:, :, i].astype(f32) for i in range(4))
        with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
            bad = fl[:, bands] | ~np.isfinite(mn) | ~np.isfinite(mx) | (mx == mn)
            valid &= ~bad.any(axis=1)
            if masked:
                valid &= (n_mask > 0) & ~np.isnan(mmn).any(axis=1)
            dd = (mx - mn).astype(f32)
            nmin = ((mmn - mn).astype(f32) / dd).astype(f32)
            nmax = ((mmx - mn).astype(f32) / dd).astype(f32)
            imin = nmin.min(axis=1).astype(f32)
            imax = nmax.max(axis=1).astype(f32)
            idiff = (imax - imin).astype(f32)
        d["imin"], d["idiff"] = np.where(valid, imin, 0), np.where(valid, idiff, 1)
        d["rescale"] = ((imax != imin) & (idiff != f32(1.0)) & valid).astype(np.int32)
        d["n_mask"] = np.minimum(n_coord, 2 ** 31 - 1)

        ratio = f32(1.0 / compactness)
        max_abs = float(ratio) * (256.0 if to_lab else 4.0)
        n = np.zeros(B, dtype=np.int64)
        grids = {}
        # the derivation depends on (h, w, n_segments / n) only: once per distinct triple, broadcast to the windows
        vi = np.nonzero(valid)[0]
        hh, ww = d["h"][vi].astype(np.int64), d["w"][vi].astype(np.int64)
        if masked:
            nn = np.minimum(n_seg[vi], n_coord[vi])
            step_u = None
        else:
            uk, inv = np.unique((hh << 44) | (ww << 24) | np.minimum(n_seg[vi], (1 << 24) - 1), return_inverse=True)
            nn_u, step_u_ = np.zeros(len(uk), np.int64), np.zeros(len(uk), np.float64)
            for u, k in enumerate(uk.tolist()):
                ys, xs, step = _grid_init(k >> 44, (k >> 24) & 0xfffff, k & 0xffffff)
                nn_u[u], step_u_[u] = len(ys) * len(xs), step
            nn, step_u = nn_u[inv], step_u_[inv]
            for i_, u in zip(vi.tolist(), inv.tolist()):
                k = int(uk[u])
                grids[i_] = _grid_init(k >> 44, (k >> 24) & 0xfffff, k & 0xffffff)
        ok = (nn > 0) & (nn < 2 ** 24)
        if step_u is not None:
            ok &= step_u > 0
        nn_safe = np.where(ok, nn, 1)
        uk, inv = np.unique((hh << 44) | (ww << 24) | nn_safe, return_inverse=True)
        cols = np.zeros((len(uk), 8), dtype=np.float64)     # sy, sx, fix_scale, ncy, ncx, km_cs, km_ncy, km_ncx
        for u, k in enumerate(uk.tolist()):
            h, w, nk = k >> 44, (k >> 24) & 0xfffff, k & 0xffffff
            sy, sx = _geometry(h, w, nk)
            cs = max(1.0, math.sqrt(float(h) * float(w) / float(nk)))
            cols[u] = (sy, sx, _fix_scale(max_abs, h, w, sy, sx), -(-h // sy), -(-w // sx), cs,
                       int(float(h) / cs) + 1, int(float(w) / cs) + 1)
        cw = cols[inv]
        seg_size = n_coord[vi].astype(np.float64) / nn_safe.astype(np.float64)
        mins = np.minimum(np.trunc(min_size_factor * seg_size), 2 ** 31 - 1).astype(np.int64)
        maxs = np.minimum(np.trunc(max_size_factor * seg_size), 2 ** 31 - 1).astype(np.int64)
        ok &= maxs >= 1
        valid[vi] = ok
        n[vi] = np.where(ok, nn, 0)
        for name, c in (("step_y", 0), ("step_x", 1), ("ncy", 3), ("ncx", 4), ("km_ncy", 6), ("km_ncx", 7)):
            d[name][vi] = np.where(ok, cw[:, c], 0).astype(np.int32)
        d["fix_scale"][vi] = np.where(ok, cw[:, 2], 0.0)
        d["km_cs"][vi] = np.where(ok, cw[:, 5], 0.0)
        d["min_size"][vi], d["max_size"][vi] = np.where(ok, mins, 0), np.where(ok, maxs, 0)
        if step_u is not None:
            stepf = step_u.astype(f32)
            with np.errstate(divide="ignore"):
                sw = (1.0 / (stepf * stepf).astype(f32).astype(np.float64)).astype(f32)
                d["sw"][vi], d["inv_w"][vi] = np.where(ok, sw, 0), np.where(ok, (f32(1.0) / sw).astype(f32), 0)
        n = np.where(valid, n, 0)
        d["valid"] = valid.astype(np.int32)
        d["n"] = n
        c0 = np.concatenate([[0], np.cumsum(n)[:-1]])
        d["c0"] = c0
        cells = np.where(valid, d["ncy"].astype(np.int64) * d["ncx"], 0)
        d["cell0"] = np.concatenate([[0], np.cumsum(cells)[:-1]])
        kmc = np.where(valid, d["km_ncy"].astype(np.int64) * d["km_ncx"], 0)
        d["km_cell0"] = np.concatenate([[0], np.cumsum(kmc)[:-1]])
        n_total, cells_total, km_cells_total = int(n.sum()), int(cells.sum()), int(kmc.sum())
        self.usable = valid.copy()
        self.n_labels = 0
        self.labels = None
        if n_total == 0:
            return self
        _lib.check(lib.obia_b200_slic_batch_prepare(d.ctypes.data_as(ctypes.c_void_p), B, Cf), "slic_batch_prepare")
        variants = int(np.bitwise_or.reduce(d["pad"][valid])) if valid.any() else 0

        draws = None
        if masked:
            # the RandomState(123) draws of the maskSLIC initialisation start now, on host threads
            draws = slic_host.submit_mask_samples([(int(n_coord[i]), int(n_seg[i])) for i in np.nonzero(valid)[0]])
        _tick("host parameters")
        self._st = dict(masked=masked, bands=bands, Cs=Cs, Cf=Cf, to_lab=to_lab, ratio=ratio, valid=valid, n=n, c0=c0,
                        n_coord=n_coord, n_seg=n_seg, n_mask=n_mask, grids=grids, mn=mn, dd=dd, n_total=n_total,
                        cells_total=cells_total, km_cells_total=km_cells_total, max_num_iter=max_num_iter,
                        start_label=start_label, variants=variants, draws=draws)
        return self

    def finish(self):
        """Second half of `segment`: centre initialisation, features, sweeps, connectivity."""
        st = getattr(self, "_st", None)
        if st is None:
            return self
        self._st = None
        lib, dev, B, C, d = self.lib, self.dev, self.B, self.C, self.desc
        masked, bands, Cs, Cf, to_lab, ratio, valid, n, c0 = (st[k] for k in (
            "masked", "bands", "Cs", "Cf", "to_lab", "ratio", "valid", "n", "c0"))
        n_coord, n_seg, n_mask, grids, mn, dd = (st[k] for k in ("n_coord", "n_seg", "n_mask", "grids", "mn", "dd"))
        n_total, cells_total, km_cells_total = st["n_total"], st["cells_total"], st["km_cells_total"]
        max_num_iter, start_label = st["max_num_iter"], st["start_label"]
        _tick(None)
        cwin = torch.from_numpy(np.repeat(np.arange(B, dtype=np.int32), n)).to(dev)
        centres = torch.empty((n_total, 2 + Cf), dtype=torch.float32, device=dev)
        if masked:
            # ---- maskSLIC initialisation: RandomState(123) draws (host threads), k-means on the device ------
            # position of the r-th mask pixel of the slab = first index whose inclusive mask count reaches r + 1
            # (np.nonzero order; no host round trip, unlike torch.nonzero)
            mask_rank = torch.cumsum(self.mask_slab.reshape(-1), 0, dtype=torch.int32)

            def positions(ranks):
                want = torch.from_numpy(np.ascontiguousarray(ranks + 1, dtype=np.int32)).to(dev)
                return torch.searchsorted(mask_rank, want).to(torch.int32)
            _tick("centre init: mask ranks")
            base = np.concatenate([[0], np.cumsum(n_mask)[:-1]])
            seeds, dense, p0, m = [], [], np.zeros(B, dtype=np.int64), np.zeros(B, dtype=np.int64)
            all_dense = True
            draws, futures = {}, st["draws"]
            for i in np.nonzero(valid)[0]:
                got = futures[(int(n_coord[i]), int(n_seg[i]))]
                idx, idx_dense = got if isinstance(got, tuple) else got.result()
                draws[i] = idx_dense
                seeds.append(idx + base[i])
                if idx_dense is not None:
                    all_dense = False
            if all_dense:
                # coord[idx_dense] is every mask pixel of the window: the k-means kernel walks the mask slab itself
                pts = None
                p0, m = base, n_mask
            else:
                off = 0
                for i in np.nonzero(valid)[0]:
                    di = np.arange(n_coord[i], dtype=np.int64) if draws[i] is None else draws[i]
                    dense.append(di + base[i])
                    p0[i], m[i] = off, len(di)
                    off += len(di)
                pts = positions(np.concatenate(dense))
            _tick("centre init: sample draws (host)")
            d["p0"], d["m"] = np.where(valid, p0, 0), np.where(valid, m, 0)
            self.upload()
            seed_pos = positions(np.concatenate(seeds))
            _tick("centre init: seed upload")
            cent = torch.empty((n_total, 2), dtype=torch.float64, device=dev)
            km_ws = torch.empty((lib.obia_b200_mask_kmeans_batch_workspace_bytes(n_total, km_cells_total),),
                                dtype=torch.uint8, device=dev)
            _lib.check(lib.obia_b200_mask_kmeans_batch(
                _p(pts), 0 if pts is None else int(pts.numel()), _p(self.mask_slab) if all_dense else None, self.hmax,
                self.wmax, _p(seed_pos),
                _p(cwin), _p(self.desc_dev), B, n_total, km_cells_total, self.slab_w, self.win_rows, 5, Cf, _p(cent),
                _p(centres), _p(km_ws), _stream()), "mask_kmeans_batch")
            del km_ws, pts, mask_rank
        else:
            self.upload()
            rows = np.zeros((n_total, 2 + Cf), dtype=np.float32)
            for i, (ys, xs, _) in grids.items():
                if not valid[i]:
                    continue
                k0 = int(c0[i])
                rows[k0:k0 + n[i], 0] = np.repeat(ys, len(xs))
                rows[k0:k0 + n[i], 1] = np.tile(xs, len(ys))
            centres.copy_(torch.from_numpy(rows))

        _tick("centre init: k-means kernels")
        # ---- K1b per window ---------------------------------------------------------------------------
        feats = torch.empty((Cf, self.slab_rows, self.pitch), dtype=torch.float32, device=dev)
        bands_dev = torch.tensor(bands, dtype=torch.int32, device=dev)
        bmin = torch.from_numpy(np.ascontiguousarray(np.where(valid[:, None], mn, 0), dtype=np.float32)).to(dev)
        bdiff = torch.from_numpy(np.ascontiguousarray(np.where(valid[:, None], dd, 1), dtype=np.float32)).to(dev)
        _lib.check(lib.obia_b200_window_features(
            _p(self.raw), self.Wl, C, _p(bands_dev), Cs, _p(bmin), _p(bdiff), _p(self.desc_dev), B, self.hmax,
            self.wmax, int(to_lab), float(ratio), _p(feats), self.slab_rows, self.pitch, _stream()), "window_features")

        _tick("features")
        # ---- K2 per window ----------------------------------------------------------------------------
        ws = torch.empty((lib.obia_b200_slic_batch_workspace_bytes(n_total, cells_total, Cf),), dtype=torch.uint8,
                         device=dev)
        labels = torch.empty((self.slab_rows, self.slab_w), dtype=torch.int32, device=dev)
        status = torch.zeros((4 + B,), dtype=torch.int32, device=dev)
        overflow = torch.zeros((B,), dtype=torch.int32, device=dev)

        def run(ignore_color):
            _lib.check(lib.obia_b200_slic_iterate_batch(
                _p(feats), _p(self.mask_slab), _p(centres), _p(labels), _p(ws), _p(self.desc_dev), _p(cwin), B, n_total,
                cells_total, self.hmax, self.wmax, self.slab_rows, self.slab_w, self.pitch, Cf, int(max_num_iter),
                int(start_label), int(ignore_color), st["variants"], _p(status), _stream()), "slic_iterate_batch")
            overflow.bitwise_or_(status[4:])

        if masked:
            run(True)     # maskSLIC step 2: spatial-only k-means moves the centres first
        run(False)
        del ws, feats

        _tick("slic sweeps")
        # ---- K3 on the slab, sizes per window -------------------------------------------------------------
        wsizes = np.ones((B, 2), dtype=np.int32)
        wsizes[valid, 0] = d["min_size"][valid]
        wsizes[valid, 1] = d["max_size"][valid]
        wsizes_dev = torch.from_numpy(wsizes).to(dev)
        cc_ws = torch.empty((lib.obia_b200_connectivity_workspace_bytes(self.slab_rows, self.slab_w),),
                            dtype=torch.uint8, device=dev)
        out = torch.empty_like(labels)
        nl = ctypes.c_int64(0)
        _lib.check(lib.obia_b200_enforce_connectivity_windows(
            _p(labels), _p(out), _p(cc_ws), self.slab_rows, self.slab_w, _p(wsizes_dev), self.win_rows,
            int(start_label), ctypes.byref(nl), _stream()), "enforce_connectivity_windows")
        del cc_ws, labels
        _tick("connectivity")
        if masked:
            out.masked_fill_(self.mask_slab == 0, -1)       # segment_boundaries.py:55-57
        # windows dropped on the device (degenerate step) or that overflowed the candidate staging
        dev_valid = self.desc_dev.view(torch.int32).reshape(B, WIN_DESC.itemsize // 4)[:, 5]
        back = torch.stack([dev_valid, overflow]).cpu().numpy()
        _tick("final read-back")
        self.usable = valid & (back[0] != 0) & (back[1] == 0)
        self.labels = out
        self.n_labels = int(nl.value)
        self.centres = centres
        return self

"""Host model of the GPU connectivity algorithm (K3) -- test helper only.

The reference's enforce-connectivity step is a sequential raster scan
(SURVEY.md 3.4 step 9).  The CUDA implementation in
obia_b200/csrc/connectivity.cu reproduces it EXACTLY with data-parallel
phases; this file is the same decomposition written as slow, obviously
order-free Python so the decomposition itself can be checked against the
sequential oracle on the CPU (tests/test_cc_model.py), independent of CUDA.

Phases (every decision taken at a scan position only reads what pieces had been
given at EARLIER scan positions, so the per-piece steps can run in any order /
in parallel and are iterated to the unique fixed point):

  1. 4-connected components of equal label; T[p] = smallest raster index of
     p's component (union-find with min-index roots on the GPU).
  2. components larger than max_size are replayed sequentially *per component*
     to reproduce the BFS cap: they split into pieces, T[p] = piece start.
  3. every piece smaller than min_size replays its own BFS to find `adjacent`
     = the last-seen neighbour pixel q that was already labelled when the
     sequential scan reached the piece:  T[q] < t  (plus the start_label=1
     corner case where a merged piece got label 0 == mask label and is
     re-scanned; tracked by `tfix`, iterated to a fixed point).
  4. kept pieces are numbered by the rank of their start pixel; merged pieces
     follow their `adjacent` chain.
"""
from __future__ import annotations

import numpy as np

INF = np.iinfo(np.int64).max
DIRS = ((0, 1), (0, -1), (1, 0), (-1, 0))  # x+1, x-1, y+1, y-1  (dy, dx)


def _bfs_piece(labels, T, t, L, max_size, visit, H, W, labelled_at, start=None):
    """Replay of the reference BFS restricted to piece t.  Returns (queue, adj_pixel)."""
    s = t if start is None else start
    queue = [s]
    visit[s] = 1
    adj = -1
    head = 0
    while head < len(queue) and len(queue) < max_size:
        p = queue[head]
        py, px = divmod(p, W)
        for dy, dx in DIRS:
            yy, xx = py + dy, px + dx
            if 0 <= xx < W and 0 <= yy < H:
                q = yy * W + xx
                if labels[q] == L and T[q] == t and not visit[q]:
                    visit[q] = 1
                    queue.append(q)
                    if len(queue) >= max_size:
                        break
                elif not (labels[q] == L and T[q] == t) and labelled_at(q):
                    adj = q
        head += 1
    return queue, adj


def enforce_connectivity_model(labels_hw, min_size, max_size, start_label=1):
    labels_hw = np.asarray(labels_hw, dtype=np.int64)
    H, W = labels_hw.shape
    N = H * W
    lab = labels_hw.ravel()
    mask_label = start_label - 1

    # ---- phase 1: components, T = min index -----------------------------
    T = np.full(N, -1, dtype=np.int64)
    for p in range(N):
        if lab[p] == mask_label or T[p] >= 0:
            continue
        stack = [p]
        T[p] = p
        while stack:
            u = stack.pop()
            uy, ux = divmod(u, W)
            for dy, dx in DIRS:
                yy, xx = uy + dy, ux + dx
                if 0 <= xx < W and 0 <= yy < H:
                    q = yy * W + xx
                    if lab[q] == lab[u] and T[q] < 0:
                        T[q] = p
                        stack.append(q)
    psize = np.zeros(N, dtype=np.int64)
    roots, counts = np.unique(T[T >= 0], return_counts=True)
    psize[roots] = counts

    # ---- phase 2: oversized components split by the BFS cap -------------
    visit = np.zeros(N, dtype=np.uint8)
    for r in roots[counts > max_size]:
        members = np.flatnonzero(T == r)          # ascending raster order
        assigned = np.zeros(N, dtype=bool)
        L = lab[r]
        for s in members:
            if assigned[s]:
                continue
            # capped BFS over not-yet-assigned members
            queue = [s]
            assigned[s] = True
            head = 0
            while head < len(queue) and len(queue) < max_size:
                py, px = divmod(queue[head], W)
                for dy, dx in DIRS:
                    yy, xx = py + dy, px + dx
                    if 0 <= xx < W and 0 <= yy < H:
                        q = yy * W + xx
                        if lab[q] == L and not assigned[q]:
                            assigned[q] = True
                            queue.append(q)
                            if len(queue) >= max_size:
                                break
                head += 1
            for q in queue:
                T[q] = s
            psize[s] = len(queue)

    starts = np.flatnonzero((T == np.arange(N)) & (lab != mask_label))
    kept = np.zeros(N, dtype=bool)
    kept[starts] = psize[starts] >= min_size

    # ---- phase 3: adjacent of small pieces, fixed point over tfix -------
    adj = np.full(N, -1, dtype=np.int64)
    tfix = np.full(N, INF, dtype=np.int64)
    small = [t for t in starts if not kept[t]]
    tfix[small] = small                      # optimistic: labelled at own time
    changed = True
    rounds = 0
    while changed:
        changed = False
        rounds += 1
        new_adj = adj.copy()
        new_tfix = tfix.copy()
        for t in small:                      # order-free: reads only old state
            L = lab[t]

            def labelled_at_time(now):
                def f(q):
                    if lab[q] == mask_label:
                        return False
                    tq = T[q]
                    if tq == t or tq > now:
                        return False
                    if kept[tq]:
                        return tq < now
                    if start_label == 0:
                        return tq < now      # merged pieces always carry a label >= 0
                    return tfix[tq] < now    # start_label == 1: label 0 == mask label
                return f

            queue, a = _bfs_piece(lab, T, t, L, max_size, visit, H, W, labelled_at_time(t))
            for q in queue:
                visit[q] = 0
            fix = t
            if a < 0 and start_label == 1:
                # merged to 0 == mask label: the scan re-enters the piece at each
                # later pixel (raster order) until a labelled neighbour is seen
                fix = INF
                for s in sorted(queue):
                    if s <= t:
                        continue
                    q2, a2 = _bfs_piece(lab, T, t, L, max_size, visit, H, W,
                                        labelled_at_time(s), start=s)
                    for q in q2:
                        visit[q] = 0
                    if a2 >= 0:
                        a, fix = a2, s
                        break
            new_adj[t] = a
            new_tfix[t] = fix
        # adj depends on other pieces only through tfix, so an unchanged tfix
        # means the assumptions this round was computed under were right
        if not np.array_equal(new_tfix, tfix):
            changed = True
        adj, tfix = new_adj, new_tfix

    # ---- phase 4: numbering + chain resolution --------------------------
    newlab = np.full(N, -1, dtype=np.int64)
    kstarts = starts[kept[starts]]
    newlab[kstarts] = start_label + np.arange(len(kstarts))
    out = np.full(N, mask_label, dtype=np.int64)
    for p in range(N):
        if lab[p] == mask_label:
            continue
        t = T[p]
        while True:
            if kept[t]:
                out[p] = newlab[t]
                break
            if adj[t] < 0:
                out[p] = 0
                break
            t = T[adj[t]]
    return out.reshape(H, W), rounds

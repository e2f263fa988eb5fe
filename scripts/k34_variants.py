"""K3 (connectivity) and K4 (zonal statistics) alone on the c2 labels, CUDA-event timed; prints checksums so that
variant libraries (OBIA_B200_LIB=...) can be compared with the main one.  GPU box only."""
import hashlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from obia_b200 import pipeline

tag = sys.argv[1] if len(sys.argv) > 1 else "main"
dev = torch.device("cuda", 0)
raw = bench.synth_raster_cuda(10000, 10000, 8, 2, dev)


def ev_time(fn, reps=7):
    """median of per-call CUDA-event times (K3 has host read-backs: single calls jitter)"""
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], out


for comp in (0.1, 10.0):
    res = pipeline.slic_labels(raw, None, n_segments=200000, compactness=comp, max_num_iter=10, keep_intermediates=True)
    seg = 1e8 / res.n_centres
    t3, (lab, nl) = ev_time(lambda: pipeline.enforce_connectivity(res.pre_connectivity, int(0.5 * seg), int(3 * seg), 1))
    assert torch.equal(lab, res.labels)
    t4, st = ev_time(lambda: pipeline.zonal_stats(res.labels, raw, None, max_label=res.n_labels))
    h3 = hashlib.sha1(lab.cpu().numpy().tobytes()).hexdigest()[:12]
    s = st.cpu().numpy()
    h4 = hashlib.sha1(s[:, :, [0, 3, 4]].tobytes()).hexdigest()[:12]     # count, min, max: exact fields
    import numpy as np
    print(f"{tag} compactness {comp}: K3 {t3:.3f} ms ({nl} segments, labels {h3})  K4 {t4:.3f} ms (exact fields {h4}, "
          f"sum of means {np.nansum(s[:, :, 1]):.9e}, kurt {np.nansum(s[:, :, 6]):.9e})", flush=True)
    del res, lab, st

"""K4 on c4 (8192 x 8192 x 64, many-band gather kernel), CUDA-event timed; for variant libraries.  GPU box only."""
import hashlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from obia_b200 import pipeline

tag = sys.argv[1] if len(sys.argv) > 1 else "main"
dev = torch.device("cuda", 0)
raw = bench.synth_raster_cuda(8192, 8192, 64, 4, dev)
res = pipeline.slic_labels(raw, None, n_segments=50000, compactness=0.3, max_num_iter=10)
ts = []
for _ in range(6):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    st = pipeline.zonal_stats(res.labels, raw, None, max_label=res.n_labels)
    b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
s = st.cpu().numpy()
gb = 8192 * 8192 * (64 * 4 + 4) / 1e9
t = sorted(ts[1:])[len(ts[1:]) // 2]
print(f"{tag} c4: K4 {t:.3f} ms = {gb / t:.2f} TB/s algorithmic ({res.n_labels} segments, exact fields "
      f"{hashlib.sha1(s[:, :, [0, 3, 4]].tobytes()).hexdigest()[:12]}, means {np.nansum(s[:, :, 1]):.9e}, "
      f"kurt {np.nansum(s[:, :, 6]):.9e})", flush=True)

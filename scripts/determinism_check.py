import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from obia_b200 import pipeline
dev = torch.device("cuda")
for (S, C, n, comp) in [(2048, 64, 3100, 0.3), (3000, 8, 18000, 0.1), (2048, 16, 3100, 0.3), (2048, 32, 3100, 0.3)]:
    raw = bench.synth_raster_cuda(S, S, C, seed=1, device=dev)
    outs = []
    for rep in range(3):
        res = pipeline.slic_labels(raw, None, n_segments=n, compactness=comp, keep_intermediates=True)
        outs.append((res.features.clone(), res.pre_connectivity.clone(), res.labels.clone(), res.centres.clone(), res.n_labels))
    for rep in (1, 2):
        f = (outs[rep][0] == outs[0][0]).all().item()
        p = (outs[rep][1] != outs[0][1]).sum().item()
        l = (outs[rep][2] != outs[0][2]).sum().item()
        c = (outs[rep][3] != outs[0][3]).sum().item()
        print(f"S={S} C={C}: rep{rep} features equal={f} pre-cc diff px={p} final diff px={l} centre words diff={c} n_labels {outs[rep][4]} vs {outs[0][4]}", flush=True)

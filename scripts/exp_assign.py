import sys, os, ctypes, torch, numpy as np, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from obia_b200 import _lib, pipeline, slic_host
import bench
variant = sys.argv[1]
_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'obia_b200', '_lib', f'libobia_exp_{variant}.so')
lib = _lib.load()
H=W=10000; C=8
raw = bench.synth_raster_cuda(H,W,C,2,torch.device('cuda'))
for comp in (0.1, 10.0):
    res = pipeline.slic_labels(raw, None, n_segments=200000, compactness=comp, max_num_iter=2, enforce_connectivity=False)
    torch.cuda.synchronize()
    lib.obia_b200_profile_enable(1)
    for _ in range(3):
        res = pipeline.slic_labels(raw, None, n_segments=200000, compactness=comp, max_num_iter=1, enforce_connectivity=False)
    torch.cuda.synchronize()
    lib.obia_b200_profile_enable(0)
    ms, n = ctypes.c_double(0), ctypes.c_int64(0)
    lib.obia_b200_profile_read(ctypes.byref(ms), ctypes.byref(n))
    print(variant, 'compactness', comp, 'assign avg ms', ms.value/n.value, 'launches', n.value, flush=True)

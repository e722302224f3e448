"""Diagnostic: one 2-site H_eff application at the ab initio shape (M = 1024, MPO bond w, d = 2, 8 % dense MPO
sites): kernel-by-kernel times under `ncu --metrics gpu__time_duration.sum`, or the whole application with CUDA
events.   python tools/hop_qc_shape.py [M] [w]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from renormalizer_b200 import ops, _lib
from renormalizer_b200.backend import asxp
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
w = int(sys.argv[2]) if len(sys.argv) > 2 else 326
_lib.get()
rng = np.random.default_rng(0)
L, R = asxp(rng.standard_normal((M, w, M))), asxp(rng.standard_normal((M, w, M)))
def site():
    s = rng.standard_normal((w, 2, 2, w)) * (rng.random((w, 2, 2, w)) < 0.08)
    return ops.MpoSite(s)
C = asxp(rng.standard_normal((M, 2, 2, M)))
plan = ops.HopPlan(L, R, [site(), site()], (M, 2, 2, M), torch.float64)
out = plan.apply(C); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.cudart().cudaProfilerStart()
e0.record()
for _ in range(3):
    plan.apply(C)
e1.record(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
flops = 2.0 * (M * w) * (4 * M) * M * 2
print(f"M={M} w={w}: hop {e0.elapsed_time(e1) / 3:.2f} ms; the two GEMMs alone would take {flops / 100e12 * 1e3:.2f} ms at 100 TFLOP/s")

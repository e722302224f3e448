"""Diagnostic: H_eff.C timing at a given bond dimension on both GEMM paths (not part of the product)."""
import sys, time, ctypes
import numpy as np
import torch
sys.path.insert(0, ".")
from renormalizer_b200 import models, ops, _lib
from renormalizer_b200.backend import backend, asxp
from renormalizer_b200.hop_expr import hop_expr_dtype
from renormalizer_b200.mpo import Mpo

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cplx = (sys.argv[2] != "real") if len(sys.argv) > 2 else True
d, w = 8, 3
lib = _lib.get()
rng = np.random.default_rng(0)
dt = torch.complex128 if cplx else torch.float64
def t(shape):
    a = rng.standard_normal(shape)
    if cplx: a = a + 1j * rng.standard_normal(shape)
    return asxp(a)
L, R, C = t((M, w, M)), t((M, w, M)), t((M, d, M))
omega, g = models.ohmic_modes(20)
site = Mpo(models.spin_boson_mpo(0.0, 1.0, omega, g, d))[5]
es = 2 if cplx else 1
flops = 2.0 * (M * w) * (d * M * es) * (M * es) + 2.0 * (M * d) * (M * es) * (w * M * es)
ref = None
for path in (0, 1):
    plan = ops.HopPlan(L, R, [site], (M, d, M), dt, path=path)
    out = plan.apply(C); torch.cuda.synchronize()
    if ref is None: ref = out.clone()
    else: print("   path1 vs path0 rel err", float((out - ref).abs().max() / ref.abs().max()))
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): plan.apply(C)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    lib.rn_profile_begin(); plan.apply(C); torch.cuda.synchronize()
    pm, pf, pc = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_long(0)
    lib.rn_profile_end(ctypes.byref(pm), ctypes.byref(pf), ctypes.byref(pc))
    print(f"M={M} cplx={cplx} path={path}: hop {ms:.3f} ms  ({flops/ms/1e9:.1f} TFLOP/s whole hop); "
          f"GEMM launches {pc.value}: {pm.value:.3f} ms -> {pf.value/pm.value/1e9:.1f} TFLOP/s")
    plan.close()

"""TEST DOUBLE for the host-logic tests -- not part of the product, never imported by it.

A pytest plugin (`-p _host_logic_stub`, with tests/ on PYTHONPATH) that replaces the C-ABI-backed
operations of `renormalizer_b200.ops` (GEMM, QR, SVD, environment update, H_eff, Krylov, the Davidson
vector kernels) by NumPy / torch-CPU equivalents built on the oracle, so that the HOST side of the
sweeps -- quantum-number bookkeeping, truncation, the propagate-and-compress / expansion / variational
compression drivers, the DMRG sweep schedule -- can be exercised against the reference's golden vectors
on a machine without a GPU.  It says nothing about the CUDA kernels: those are covered by `-m gpu`.
`tests/test_host_logic.py::test_sweep_host_logic_with_stubbed_kernels` runs it in a subprocess; the
product itself still refuses to run without the CUDA library (`test_compute_without_gpu_fails_loudly`).
"""
import sys
import numpy as np, torch
import os
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import renormalizer_b200.backend
bk = sys.modules["renormalizer_b200.backend"]
bk.Backend.device = property(lambda self: torch.device("cpu"))
def asxp(array, dtype=None):
    if array is None: return None
    if hasattr(array, "array") and not isinstance(array, (np.ndarray, torch.Tensor)): array = array.array
    t = array if isinstance(array, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(array))
    if dtype is not None and t.dtype != dtype: t = t.to(dtype)
    elif t.dtype not in (torch.float64, torch.complex128): t = t.to(torch.complex128 if t.is_complex() else torch.float64)
    return t.contiguous()
bk.asxp = asxp
from renormalizer_b200 import ops
def _prom(a, b):
    if a.is_complex() != b.is_complex():
        a, b = a.to(torch.complex128), b.to(torch.complex128)
    return a, b
ops.matmul = lambda a, b: torch.matmul(*_prom(a, b))
ops.tensordot1 = lambda a, b: torch.tensordot(*_prom(a, b), dims=1)
def qr(a, lq=False):
    if not lq: return torch.linalg.qr(a)
    q, r = torch.linalg.qr(a.conj().T)
    return r.conj().T.contiguous(), q.conj().T.contiguous()
ops.qr = qr
def svd(a, **k):
    u, s, vh = torch.linalg.svd(a, full_matrices=False)
    svd.last_sweeps = 1
    return u, s, vh
ops.svd = svd
from oracle.contract import env_update as _eu, hop_apply as _hop
def env_update(environ, bra, ket, site, domain, path=None):
    return torch.from_numpy(np.ascontiguousarray(_eu(environ.resolve_conj().numpy(), ket.resolve_conj().numpy(), site.array, domain, ms_conj=bra.resolve_conj().numpy().conj())))
ops.env_update = env_update
ops.MpoSite.dense = property(lambda self: torch.from_numpy(self.array))
import renormalizer_b200.svd_qn as sq, renormalizer_b200.mps as mpsmod, renormalizer_b200.lib as libmod, renormalizer_b200.hop_expr as hopmod, renormalizer_b200.gs as gsmod
for mod in (sq, mpsmod, libmod, hopmod, gsmod):
    if hasattr(mod, "asxp"): mod.asxp = asxp
class _Hop:
    plan = None
    def __init__(self, l, r, cmo, shape, dtype):
        self.l, self.r, self.w, self.shape, self.dtype = l.numpy(), r.numpy(), [np.asarray(ops.as_mpo_site(m).array) for m in cmo], tuple(shape), dtype
    def __call__(self, c):
        out = _hop(self.l, self.r, self.w, asxp(c).resolve_conj().numpy().reshape(self.shape))
        return torch.from_numpy(np.ascontiguousarray(out)).to(self.dtype)
    def close(self): pass
mpsmod.hop_expr_dtype = lambda l, r, cmo, shape, dtype: _Hop(asxp(l), asxp(r), cmo, shape, dtype)
from oracle.krylov import expm_krylov as _ek
def expm_krylov(afunc, dt, v):
    res, j = _ek(lambda y: afunc(torch.from_numpy(np.ascontiguousarray(y))).numpy().ravel(), dt, v.resolve_conj().numpy())
    return torch.from_numpy(np.ascontiguousarray(res)), j
mpsmod.expm_krylov = expm_krylov

# ---- Davidson / DMRG stubs
class _WS:
    def __init__(self, device, nvec_max=64): self.nvec_max = nvec_max
ops.VecWorkspace = _WS
def multi_dot(V, x, nvec, n, cplx, ws, out=None):
    V2 = V.reshape(-1, n)[:nvec]
    r = (V2.conj() @ x.reshape(-1).to(V2.dtype))
    o = torch.view_as_real(r.to(torch.complex128)).reshape(-1).clone() if cplx else torch.stack([r, torch.zeros_like(r)], 1).reshape(-1)
    if out is not None:
        out.reshape(-1)[:o.numel()] = o
        return out
    return o
ops.multi_dot = multi_dot
def lincomb(V, coef, nvec, n, cplx, out):
    out.copy_((coef.reshape(-1)[:nvec].to(V.dtype)[:, None] * V[:nvec]).sum(0))
    return out
ops.lincomb = lincomb
import renormalizer_b200.davidson as davmod
gsmod.hop_expr_dtype = mpsmod.hop_expr_dtype

"""Diagnostic: ONE mid-chain TDVP-PS site update (forward Krylov, QR, environment update, backward
bond Krylov, absorb) at the bench shape, bracketed by cudaProfilerStart/Stop for
`ncu --profile-from-start off` launch lists (not part of the product)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from renormalizer_b200 import models, ops, _lib
from renormalizer_b200.backend import backend, asxp
from renormalizer_b200.hop_expr import hop_expr_dtype
from renormalizer_b200.krylov import expm_krylov
from renormalizer_b200.lib import contract_one_site
from renormalizer_b200.mpo import Mpo
from renormalizer_b200.svd_qn import svd_qn, add_outer

M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
d, w = 8, 3
_lib.get()
rng = np.random.default_rng(0)
def c(shape):
    return asxp(rng.standard_normal(shape) + 1j * rng.standard_normal(shape))
def herm(M, w):
    e = rng.standard_normal((M, w, M)) + 1j * rng.standard_normal((M, w, M))
    return asxp((e + e.conj().transpose(2, 1, 0)) / (4 * M))
L, R = herm(M, w), herm(M, w)
omega, g = models.ohmic_modes(20, alpha=0.05, omega_c=20.0)
mpo = Mpo(models.spin_boson_mpo(0.0, 1.0, omega, g, d))
site = mpo[5]
C = c((M, d, M)); C = C / torch.linalg.vector_norm(C)
nxt = c((M, d, M))
qn0 = np.zeros((M, 1), dtype=int); sq = np.zeros((d, 1), dtype=int)

def site_update():
    hop = hop_expr_dtype(L, R, [site], (M, d, M), torch.complex128)
    ct, j1 = expm_krylov(hop, -0.025j, C.reshape(-1))
    hop.close()
    u, _, v, _ = svd_qn(ct.reshape(M, d, M), add_outer(qn0, sq), qn0, np.array([0]), QR=True, system="L",
                        full_matrices=False)
    a = u.contiguous().reshape(M, d, -1)
    vt = v.transpose(0, 1).contiguous()
    l2 = contract_one_site(L, a, site, "L")
    hop0 = hop_expr_dtype(l2, R, [], tuple(vt.shape), torch.complex128)
    back, j2 = expm_krylov(hop0, 0.025j, vt.reshape(-1))
    hop0.close()
    out = ops.tensordot1(back.reshape(vt.shape), nxt)
    return j1, j2

for _ in range(2):
    steps = site_update()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
t0 = time.perf_counter()
steps = site_update()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
torch.cuda.cudart().cudaProfilerStop()
print(f"M={M}: one site update {dt*1e3:.3f} ms, Krylov steps fwd/bwd {steps}")

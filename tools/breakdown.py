"""Diagnostic: time the pieces of one mid-chain TDVP-PS site update (not part of the product)."""
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
from renormalizer_b200.svd_qn import svd_qn

path = int(sys.argv[1]) if len(sys.argv) > 1 else 0
M = int(sys.argv[2]) if len(sys.argv) > 2 else 256
d, w = 8, 3
backend.gemm_path = path
_lib.get()
rng = np.random.default_rng(0)
def c(shape):
    return asxp(rng.standard_normal(shape) + 1j * rng.standard_normal(shape))
def herm(M, w):
    e = rng.standard_normal((M, w, M)) + 1j * rng.standard_normal((M, w, M))
    return asxp(e + e.conj().transpose(2, 1, 0))
L, R = herm(M, w), herm(M, w)
omega, g = models.ohmic_modes(20)
mpo = Mpo(models.spin_boson_mpo(0.0, 1.0, omega, g, d))
site = mpo[5]
C = c((M, d, M)); C = C / torch.linalg.vector_norm(C)

def timeit(name, fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n * 1e3
    print(f"{name:28s} {dt:8.3f} ms")
    return r

hop = hop_expr_dtype(L, R, [site], (M, d, M), torch.complex128)
timeit("hop1 apply", lambda: hop(C), 20)
hop0 = hop_expr_dtype(L, R, [], (M, M), torch.complex128)
C0 = c((M, M))
timeit("hop0 apply", lambda: hop0(C0), 20)
scale = 1.0 / float(torch.linalg.vector_norm(hop(C)))
res = timeit("krylov fwd (site)", lambda: expm_krylov(lambda y: hop(y) * scale, -0.025j, C.reshape(-1)), 3)
print("   steps", res[1])
res = timeit("krylov bwd (bond)", lambda: expm_krylov(lambda y: hop0(y) * scale, 0.025j, C0.reshape(-1)), 3)
print("   steps", res[1])
qn0 = np.zeros((M, 1), dtype=int); sq = np.zeros((d, 1), dtype=int)
from renormalizer_b200.svd_qn import add_outer
timeit("svd_qn QR (L)", lambda: svd_qn(C, add_outer(qn0, sq), qn0, np.array([0]), QR=True, system="L", full_matrices=False), 5)
timeit("svd_qn QR (R)", lambda: svd_qn(C, qn0, add_outer(sq, qn0), np.array([0]), QR=True, system="R", full_matrices=False), 5)
timeit("env update L", lambda: contract_one_site(L, C, site, "L"), 10)
timeit("env update R", lambda: contract_one_site(R, C, site, "R"), 10)
timeit("tensordot1", lambda: ops.tensordot1(C0, C), 10)
timeit("hop plan create+close", lambda: hop_expr_dtype(L, R, [site], (M, d, M), torch.complex128).close(), 10)
a = asxp(rng.standard_normal((M * d, M)))

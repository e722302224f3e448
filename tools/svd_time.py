"""Diagnostic: time and sweep count of the SVD (bare Jacobi vs QR-preconditioned) on a few shapes
(not part of the product).   python tools/svd_time.py"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from renormalizer_b200 import _lib, ops
_lib.get()
rng = np.random.default_rng(0)
def run(m, n, cplx=False, decay=None, rank=None, modes=(False, True)):
    a = rng.standard_normal((m, n))
    if cplx:
        a = a + 1j * rng.standard_normal((m, n))
    if decay is not None or rank is not None:
        u, s, vh = np.linalg.svd(a, full_matrices=False)
        k = len(s)
        s = np.exp(-decay * np.arange(k)) if decay is not None else s
        if rank is not None:
            s[rank:] = 0
        a = (u * s) @ vh
    ad = torch.from_numpy(a).cuda()
    k = min(m, n)
    sref = np.linalg.svd(a, compute_uv=False)
    for pre in modes:
        for it in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            u, s, vh = ops.svd(ad, precondition=pre)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        rec = ((u * s) @ vh).cpu().numpy()
        uo = (u.conj().T @ u).cpu().numpy()
        print(f"m={m} n={n} cplx={cplx} decay={decay} rank={rank} precond={pre}: {dt*1e3:.1f} ms, sweeps {ops.svd.last_sweeps}, "
              f"sigma err {np.abs(s.cpu().numpy() - sref).max() / sref[0]:.1e}, recon {np.abs(rec - a).max():.1e}, "
              f"|UhU-I| {np.abs(uo - np.eye(k)).max():.1e}", flush=True)
for args in [(1432, 512), (512, 1432, False, 0.05), (1432, 512, False, 0.05), (1432, 512, False, None, 300),
             (2048, 2048, False, 0.02), (1432, 512, True, 0.05), (256, 256, False, 0.1), (4096, 1024, False, 0.03)]:
    run(*args)

"""Krylov approximation of expm(dt*A) v on the device.

Mirror of renormalizer/lib/krylov/krylov.py:28-84 (expm_krylov): same Lanczos recurrence and the
same stopping rules, so the number of H_eff applications matches the reference.  The Krylov
vectors, alpha and beta never leave HBM; the host only sees the tridiagonal coefficients at the
reference's own convergence check points (every second step from the fifth on).
"""
import numpy as np
import torch
from scipy.linalg import eigh_tridiagonal

from . import ops


def _coef(alpha, beta, nrm, dt):
    """krylov.py:15-25 (_expm_krylov) without the final V @ ... product."""
    try:
        w, u = eigh_tridiagonal(alpha, beta)
    except np.linalg.LinAlgError:
        w, u = np.linalg.eigh(np.diag(alpha) + np.diag(beta, -1) + np.diag(beta, 1))
    return u @ (nrm * np.exp(dt * w) * u[0])


class _Lanczos:
    def __init__(self, n, dtype, device, block_size):
        self.n, self.dtype, self.device = n, dtype, device
        self.cplx = dtype == torch.complex128
        self.V = torch.empty((block_size, n), dtype=dtype, device=device)
        self.alpha = torch.zeros((block_size, 2), dtype=torch.float64, device=device)
        self.beta = torch.zeros((block_size, 2), dtype=torch.float64, device=device)
        self.ws = ops.VecWorkspace(device, nvec_max=1)
        self.block_size = block_size

    def grow(self):
        bs = self.block_size
        self.V = torch.cat([self.V, torch.empty((bs, self.n), dtype=self.dtype, device=self.device)])
        z = torch.zeros((bs, 2), dtype=torch.float64, device=self.device)
        self.alpha = torch.cat([self.alpha, z])
        self.beta = torch.cat([self.beta, z.clone()])

    def combine(self, coef, m):
        out = torch.empty(self.n, dtype=self.dtype, device=self.device)
        c = np.asarray(coef)
        if self.cplx:
            c = c.astype(np.complex128)
        else:
            c = c.astype(np.float64)
        cd = torch.from_numpy(c).to(self.device)
        ops.lincomb(self.V, cd, m, self.n, self.cplx, out)
        return out


def expm_krylov(Afunc, dt, vstart, block_size=50):
    """Return (expm(dt*A) @ vstart, number of A applications); A Hermitian, given as
    Afunc(device vector) -> device vector.  When Afunc is an H_eff callable of this package the
    whole Lanczos iteration runs as one fused C call (rn_lanczos_step)."""
    plan = getattr(Afunc, "plan", None)
    if not np.iscomplex(dt):
        dt = dt.real
    vstart = vstart.reshape(-1).contiguous()
    if np.iscomplex(dt) and not vstart.is_complex():
        vstart = vstart.to(torch.complex128)
    n = vstart.numel()
    if plan is not None and plan.dtype == vstart.dtype and (vstart.is_complex() or not np.iscomplex(dt)):
        got = ops.expm_krylov_plan(plan, vstart, dt)
        if got is not None:
            return got
    st = _Lanczos(n, vstart.dtype, vstart.device, block_size)
    nrmv = float(torch.linalg.vector_norm(vstart))
    assert nrmv > 0
    st.V[0] = vstart / nrmv
    eps_break = 100 * n * np.finfo(float).eps
    res = None
    alpha_h = beta_h = None

    def finish(m):
        a = alpha_h[:m, 0].copy()
        b = beta_h[:m - 1, 0].copy()
        return st.combine(_coef(a, b, nrmv, dt), m), m

    fused = plan is not None and plan.dtype == st.dtype and n > 1
    wbuf = torch.empty(n, dtype=st.dtype, device=vstart.device) if fused else None
    for j in range(n):
        if fused and j < n - 1:
            if st.V.shape[0] == j + 1:
                st.grow()
            ops.lanczos_step(plan, n, st.V, j, st.alpha, st.beta, wbuf, st.ws)
            w = wbuf
        else:
            w = Afunc(st.V[j])
            w = w.reshape(-1)
            if w.dtype != st.dtype:
                w = w.to(st.dtype)
            if not w.is_contiguous():
                w = w.contiguous()
            # alpha_j = Re <w, v_j>  (== Re <v_j, w>)
            ops.multi_dot(st.V[j], w, 1, n, st.cplx, st.ws, out=st.alpha[j])
        if j == n - 1:
            alpha_h = st.alpha[:j + 1].cpu().numpy()
            beta_h = st.beta[:j + 1].cpu().numpy()
            first_bad = _first_breakdown(beta_h[:j, 0], eps_break)
            return finish(j + 1 if first_bad is None else first_bad + 1)
        if not fused:
            if st.V.shape[0] == j + 1:
                st.grow()
            ops.lanczos_update(w, st.V[j], st.V[j - 1] if j > 0 else None, st.alpha[j],
                               st.beta[j - 1] if j > 0 else None, st.ws, st.beta[j])
        check = 3 < j and j % 2 == 0
        if check or j < 4 and n <= 8:
            alpha_h = st.alpha[:j + 1].cpu().numpy()
            beta_h = st.beta[:j + 1].cpu().numpy()
            first_bad = _first_breakdown(beta_h[:j + 1, 0], eps_break)
            if first_bad is not None:
                return finish(first_bad + 1)
        if check:
            new_res = st.combine(_coef(alpha_h[:j + 1, 0].copy(), beta_h[:j, 0].copy(), nrmv, dt), j + 1)
            if res is not None and ops.allclose(res, new_res):
                return new_res, j + 1
            res = new_res
        if not fused:
            ops.scale_inv(w, st.beta[j], st.V[j + 1])
    raise RuntimeError("unreachable")


def _first_breakdown(beta, eps_break):
    """Index j of the first beta_j below the reference's breakdown threshold (krylov.py:75)."""
    bad = np.nonzero(~(beta >= eps_break))[0]
    return int(bad[0]) if len(bad) else None

"""Oracle: Lanczos approximation of expm(dt*A) v for Hermitian A."""
import numpy as np
from scipy.linalg import eigh_tridiagonal


def _project_back(alpha, beta, vt, nrm, dt):
    """Reference: lib/krylov/krylov.py:15-25 (_expm_krylov)."""
    try:
        w, u = eigh_tridiagonal(alpha, beta)
    except np.linalg.LinAlgError:
        w, u = np.linalg.eigh(np.diag(alpha) + np.diag(beta, -1) + np.diag(beta, 1))
    return vt @ (u @ (nrm * np.exp(dt * w) * u[0]))


def expm_krylov(afunc, dt, vstart, block_size=50):
    """Returns (expm(dt*A) @ vstart, number of A applications).

    Reference: renormalizer/lib/krylov/krylov.py:28-84.  Stops when the Krylov space is
    exhausted, when beta underflows (100*n*eps), or when two successive approximations taken
    every second step (from step 4 on) agree to numpy.allclose defaults.
    """
    if not np.iscomplex(dt):
        dt = dt.real
    vstart = np.asarray(vstart)
    n = len(vstart)
    nrm = float(np.linalg.norm(vstart))
    assert nrm > 0
    alpha = np.zeros(block_size)
    beta = np.zeros(block_size - 1)
    V = np.empty((block_size, n), dtype=vstart.dtype)
    V[0] = vstart / nrm
    last = None
    for j in range(n):
        w = afunc(V[j])
        alpha[j] = np.vdot(w, V[j]).real
        if j == n - 1:
            return _project_back(alpha[:j + 1], beta[:j], V[:j + 1].T, nrm, dt), j + 1
        if len(V) == j + 1:
            V = np.concatenate([V, np.empty((block_size, n), dtype=V.dtype)])
            alpha = np.concatenate([alpha, np.zeros(block_size)])
            beta = np.concatenate([beta, np.zeros(block_size)])
        w = w - (alpha[j] * V[j] + (beta[j - 1] * V[j - 1] if j > 0 else 0))
        beta[j] = np.linalg.norm(w)
        if beta[j] < 100 * n * np.finfo(float).eps:
            return _project_back(alpha[:j + 1], beta[:j], V[:j + 1].T, nrm, dt), j + 1
        if 3 < j and j % 2 == 0:
            cur = _project_back(alpha[:j + 1], beta[:j], V[:j + 1].T, nrm, dt)
            if last is not None and np.allclose(last, cur):
                return cur, j + 1
            last = cur
        V[j + 1] = w / beta[j]

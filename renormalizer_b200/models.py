"""Synthetic models of the shapes BASELINE.json names: MPO site tensors built by hand as finite
state machines (same sparse structure and bond dimensions as the reference's symbolic MPOs) and
random / product initial states.  Host-side set-up code (NumPy); nothing here is on the hot path.

Spin-boson (renormalizer/model/model.py:410-439, SpinBosonModel; example/sbm.py):
    H = eps sigma_z + delta sigma_x + sum_i w_i b_i^+ b_i + sigma_z sum_i g_i (b_i^+ + b_i)
Holstein chain (model.py:236-345, HolsteinModel scheme 1/2; renormalizer/tests/parameter.py):
    H = sum_i e_i a_i^+ a_i + J sum_i (a_i^+ a_{i+1} + h.c.) + sum_i w b_i^+ b_i
        + g w sum_i a_i^+ a_i (b_i^+ + b_i)
"""
import numpy as np


def _boson_ops(d):
    b = np.diag(np.sqrt(np.arange(1, d)), 1)
    return np.eye(d), b.T @ b, b + b.T


def spin_boson_mpo(eps, delta, omegas, couplings, nlevels):
    """Site tensors W[b, up, down, f]; sites = [spin, ph_1, ..., ph_N]; bond dimension 3."""
    n = len(omegas)
    assert n >= 1 and len(couplings) == n
    sz = np.diag([1.0, -1.0])
    sx = np.array([[0.0, 1.0], [1.0, 0.0]])
    i2 = np.eye(2)
    sites = []
    w = np.zeros((1, 2, 2, 3))
    w[0, :, :, 0] = i2                       # nothing placed yet
    w[0, :, :, 1] = sz                       # sigma_z waiting for a coordinate
    w[0, :, :, 2] = eps * sz + delta * sx    # complete term
    sites.append(w)
    for k in range(n):
        d = nlevels[k] if hasattr(nlevels, "__len__") else nlevels
        idn, num, x = _boson_ops(d)
        last = k == n - 1
        w = np.zeros((3, d, d, 1 if last else 3))
        done = 0 if last else 2
        w[0, :, :, done] = omegas[k] * num
        w[1, :, :, done] = couplings[k] * x
        w[2, :, :, done] = idn
        if not last:
            w[0, :, :, 0] = idn
            w[1, :, :, 1] = idn
        sites.append(w)
    return sites


def spin_boson_sigma_z_mpo(nmodes, nlevels):
    sites = [np.diag([1.0, -1.0]).reshape(1, 2, 2, 1)]
    for k in range(nmodes):
        d = nlevels[k] if hasattr(nlevels, "__len__") else nlevels
        sites.append(np.eye(d).reshape(1, d, d, 1))
    return sites


def ohmic_modes(nmodes, alpha=0.05, omega_c=20.0, delta=1.0):
    """Trapezoid discretisation of an Ohmic bath (reference renormalizer/sbm/lib.py:126-135)."""
    x0, x1 = 0.0, omega_c
    dw = (x1 - x0) / nmodes

    def j(w):
        return np.pi / 2 * alpha * w * np.exp(-w / omega_c)
    xs = x0 + dw * np.arange(nmodes + 1)
    omega = (xs[:-1] + xs[1:]) / 2
    c2 = (j(xs[:-1]) + j(xs[1:])) / 2 * 2 / np.pi * omega * dw
    g = np.sqrt(c2) / np.sqrt(2 * omega)      # c_i x_i = c_i (b + b^+) / sqrt(2 w_i)
    return omega, g


def holstein_mpo(nmol, nlevels, e0=0.0, j=-0.1, omega=0.2, g=1.0):
    """Site tensors for sites [e_1, ph_1, ..., e_N, ph_N]; bond dimensions alternate 5 / 4."""
    adag = np.array([[0.0, 0.0], [1.0, 0.0]])
    a = adag.T
    num_e = adag @ a
    i2 = np.eye(2)
    idn, num_b, x = _boson_ops(nlevels)
    # bond states: 0 start, 1 done, 2 a^+ pending, 3 a pending, 4 n pending (own phonon)
    sites = []
    for m in range(nmol):
        first, last = m == 0, m == nmol - 1
        we = np.zeros((1 if first else 4, 2, 2, 5))
        we[0, :, :, 0] = i2
        we[0, :, :, 1] = e0 * num_e
        if not last:
            we[0, :, :, 2] = j * adag
            we[0, :, :, 3] = j * a
        we[0, :, :, 4] = g * omega * num_e
        if not first:
            we[1, :, :, 1] = i2
            we[2, :, :, 1] = a
            we[3, :, :, 1] = adag
        sites.append(we)
        wp = np.zeros((5, nlevels, nlevels, 1 if last else 4))
        if last:
            wp[0, :, :, 0] = omega * num_b
            wp[1, :, :, 0] = idn
            wp[4, :, :, 0] = x
        else:
            wp[0, :, :, 0] = idn
            wp[0, :, :, 1] = omega * num_b
            wp[1, :, :, 1] = idn
            wp[2, :, :, 2] = idn
            wp[3, :, :, 3] = idn
            wp[4, :, :, 1] = x
        sites.append(wp)
    return sites


def holstein_sigmaqn(nmol, nlevels):
    """Exciton-number quantum numbers of the physical indices."""
    out = []
    for _ in range(nmol):
        out.append(np.array([[0], [1]]))
        out.append(np.zeros((nlevels, 1), dtype=int))
    return out


def mpo_to_dense(sites):
    """Dense matrix of a small MPO (tests only)."""
    t = sites[0][0]                                  # (up, down, f)
    for w in sites[1:]:
        t = np.tensordot(t, w, axes=(-1, 0))         # (..., up, down, f)
    t = t[..., 0]
    n = len(sites)
    ups = list(range(0, 2 * n, 2))
    downs = list(range(1, 2 * n, 2))
    t = t.transpose(ups + downs)
    dim = int(np.prod(t.shape[:n]))
    return t.reshape(dim, dim)


def random_mps_sites(pdims, m_max, rng, dtype=np.float64):
    """Left-canonical random MPS without quantum numbers: bond dimensions min(m_max, exact)."""
    n = len(pdims)
    left = [1]
    for d in pdims:
        left.append(left[-1] * d)
    right = [1]
    for d in reversed(pdims):
        right.append(right[-1] * d)
    right = right[::-1]
    dims = [int(min(m_max, l, r)) for l, r in zip(left, right)]
    sites = []
    for i in range(n):
        dl, d, dr = dims[i], pdims[i], dims[i + 1]
        a = rng.standard_normal((dl * d, dr))
        if dtype == np.complex128:
            a = a + 1j * rng.standard_normal((dl * d, dr))
        if i < n - 1:
            q, _ = np.linalg.qr(a)
            sites.append(q.reshape(dl, d, dr).astype(dtype))
        else:
            a = a / np.linalg.norm(a)
            sites.append(a.reshape(dl, d, dr).astype(dtype))
    return sites


def random_mps_qn(sigmaqn, qntot, m_max, rng):
    """Random left-canonical MPS with a conserved quantum number, built block by block like the
    reference's Mps.random (mps/mps.py:120-185).  Returns (sites, qn list)."""
    qntot = np.asarray(qntot)
    qn_size = len(qntot)
    n = len(sigmaqn)
    qn = [np.zeros((1, qn_size), dtype=int)]
    sites = []
    dim_prev = 1
    for i in range(n - 1):
        sq = np.asarray(sigmaqn[i])
        big = (qn[i][:, None, :] + sq[None, :, :]).reshape(-1, qn_size)
        cols, col_qn = [], []
        sectors = sorted(set(tuple(t) for t in big))
        sectors = [s for s in sectors if not np.all(qntot < np.array(s))]
        blocks = {}
        for s in sectors:
            idx = np.where(np.all(big == np.array(s), axis=1))[0]
            a = rng.standard_normal((len(idx), len(idx)))
            q, _ = np.linalg.qr(a)
            blocks[s] = (idx, q)
        # spread the retained states evenly over the sectors, then fill up
        total = sum(len(v[0]) for v in blocks.values())
        target = min(m_max, total)
        quota = {s: min(len(blocks[s][0]), target // len(sectors)) for s in sectors}
        left = target - sum(quota.values())
        while left > 0:
            progressed = False
            for s in sectors:
                if left > 0 and quota[s] < len(blocks[s][0]):
                    quota[s] += 1
                    left -= 1
                    progressed = True
            if not progressed:
                break
        for s in sectors:
            idx, q = blocks[s]
            for c in range(quota[s]):
                v = np.zeros(len(big))
                v[idx] = q[:, c]
                cols.append(v)
                col_qn.append(s)
        mt = np.stack(cols, axis=1)
        dim = mt.shape[1]
        sites.append(mt.reshape(dim_prev, sq.shape[0], dim))
        qn.append(np.array(col_qn))
        dim_prev = dim
    qn.append(np.zeros((1, qn_size), dtype=int))
    sq = np.asarray(sigmaqn[-1])
    last = rng.standard_normal((dim_prev, sq.shape[0], 1)) - 0.0
    big = qn[-2][:, None, None, :] + sq[None, :, None, :]
    mask = np.all(big == qntot, axis=-1)
    last[~mask] = 0
    last /= np.linalg.norm(last)
    sites.append(last)
    return sites, qn

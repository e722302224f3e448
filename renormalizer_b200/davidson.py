"""Davidson eigensolver with device-resident subspace.

Mirror of renormalizer/lib/davidson/davidson.py:73-455 (PySCF davidson1 with the reference's
switches): same subspace schedule, convergence test (|de| < tol and |r| < sqrt(tol)), restart and
linear-dependency rules.  Trial vectors and their images live in two HBM stacks; inner products
are batched (rn_multi_dot), linear combinations are rn_lincomb, and only the small Rayleigh
matrix is diagonalised on the host.
"""
import numpy as np
import scipy.linalg
import torch

from . import ops


def _dots(stack, nvec, x, n, cplx, ws):
    """<stack_i, x> for i < nvec as a host complex/real array (one device->host copy)."""
    out = ops.multi_dot(stack, x, nvec, n, cplx, ws).cpu().numpy().reshape(nvec, 2)
    return out[:, 0] + 1j * out[:, 1] if cplx else out[:, 0].copy()


def _norm(x, n, cplx, ws):
    return float(np.sqrt(max(_dots(x.reshape(1, -1), 1, x, n, cplx, ws)[0].real, 0.0)))


def _orthonormalise(vecs, n, cplx, ws, lindep):
    """davidson.py:467-491 (_qr): Gram-Schmidt, dropping dependent vectors."""
    out = []
    for x in vecs:
        x = x.clone()
        for q in out:
            x -= q * complex(_dots(q.reshape(1, -1), 1, x, n, cplx, ws)[0]) if cplx else \
                q * float(_dots(q.reshape(1, -1), 1, x, n, cplx, ws)[0])
        nrm = _norm(x, n, cplx, ws)
        if nrm ** 2 > lindep:
            out.append(x / nrm)
    return out


def davidson(aop, x0, precond, tol=1e-12, max_cycle=50, max_space=12, lindep=1e-14, nroots=1):
    """Lowest `nroots` eigenpairs of the Hermitian operator aop(device vector) -> device vector.
    x0: device vector or list of them; precond(dx, e, x0) -> device vector.
    Returns (e, c) like the reference: float / vector for nroots == 1, arrays / list otherwise."""
    toloose = np.sqrt(tol)
    if isinstance(x0, torch.Tensor):
        x0 = [x0]
    x0 = [x.reshape(-1) for x in x0]
    n = x0[0].numel()
    dtype, dev = x0[0].dtype, x0[0].device
    cplx = dtype == torch.complex128
    max_space = max_space + (nroots - 1) * 3
    cap = max_space + nroots + 40
    XS = torch.empty((cap, n), dtype=dtype, device=dev)
    AX = torch.empty((cap, n), dtype=dtype, device=dev)
    ws = ops.VecWorkspace(dev, nvec_max=cap)
    heff = np.zeros((cap, cap), dtype=np.complex128 if cplx else np.float64)
    fresh_start = True
    e = 0
    v = None
    conv = [False] * nroots
    space = 0
    xt = None

    def combine(stack, vk, m):
        out = torch.empty(n, dtype=dtype, device=dev)
        c = torch.from_numpy(np.ascontiguousarray(vk.astype(np.complex128 if cplx else np.float64))).to(dev)
        return ops.lincomb(stack, c, m, n, cplx, out)

    for icyc in range(max_cycle):
        if fresh_start:
            space = 0
            xt = _orthonormalise(x0, n, cplx, ws, lindep)
            if len(xt) == 0:
                raise RuntimeError("davidson: initial guess is empty or zero")
            x0 = None
        elif len(xt) > 1:
            xt = _orthonormalise(xt, n, cplx, ws, lindep)[:40]
        head = space
        for x in xt:
            XS[space] = x
            ax = aop(x).reshape(-1)
            AX[space] = ax
            space += 1
        elast, vlast, conv_last = e, v, conv
        # new rows / columns of the Rayleigh matrix (davidson.py:56-70)
        for j in range(head, space):
            d = _dots(AX, space, XS[j], n, cplx, ws)       # <ax_i, x_j> = conj(<x_j, ax_i>)
            row = np.conj(d)                               # heff[j, i] = <x_j, ax_i>
            for i in range(space):
                if i < head or i <= j:
                    heff[j, i] = row[i]
                    heff[i, j] = np.conj(row[i])
            heff[j, j] = row[j].real
        xt = None
        w, v = scipy.linalg.eigh(heff[:space, :space])
        e = w[:nroots]
        v = v[:, :nroots]
        x0 = [combine(XS, v[:, k], space) for k in range(v.shape[1])]
        ax0 = [combine(AX, v[:, k], space) for k in range(v.shape[1])]
        if not fresh_start:
            hd = vlast.shape[0]
            idx = np.argmax(abs(np.dot(v[:hd].conj().T, vlast)), axis=1)
            elast = [elast[i] for i in idx]
            conv_last = [conv_last[i] for i in idx]
        de = e - elast
        dx_norm, xt, conv = [], [], [False] * len(e)
        for k, ek in enumerate(e):
            r = ax0[k] - x0[k] * float(ek)
            xt.append(r)
            dx_norm.append(_norm(r, n, cplx, ws))
            conv[k] = abs(de[k]) < tol and dx_norm[k] < toloose
        ax0 = None
        if all(conv):
            break
        if any((not conv[k]) and nr ** 2 > lindep for k, nr in enumerate(dx_norm)):
            keep = [(not conv[k]) and dx_norm[k] ** 2 > lindep for k in range(len(e))]
        else:
            keep = [dx_norm[k] ** 2 > lindep for k in range(len(e))]
        new = []
        for k in range(len(e)):
            if keep[k]:
                t = precond(xt[k], e[0], x0[k])
                t = t * (1 / _norm(t, n, cplx, ws))
                new.append(t)
        xt = new
        # project out the current subspace (davidson.py:407-411), batched over the stack
        for t in xt:
            c = _dots(XS, space, t, n, cplx, ws)
            t -= combine(XS, c, space)
        new = []
        for t in xt:
            nrm = _norm(t, n, cplx, ws)
            if nrm ** 2 > lindep:
                new.append(t * (1 / nrm))
        xt = new
        if len(xt) == 0:
            break
        fresh_start = space + nroots > max_space
    if nroots == 1:
        return float(e[0]), x0[0]
    return e, x0

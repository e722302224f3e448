"""Oracle: Davidson eigensolver as used by the reference's DMRG (restated from the algorithm in
renormalizer/lib/davidson/davidson.py:73-455, itself PySCF's lib.linalg_helper.davidson1)."""
import numpy as np
import scipy.linalg


def _orthonormalise(vecs, lindep):
    """Modified Gram-Schmidt dropping dependent vectors.  Reference: davidson.py:467-491 (_qr)."""
    out = []
    for x in vecs:
        x = np.array(x, copy=True)
        for q in out:
            x -= q * np.dot(q.conj(), x)
        nrm2 = np.dot(x.conj(), x).real
        if nrm2 > lindep:
            out.append(x / np.sqrt(nrm2))
    return out


def _combine(v, xs):
    """Reference: davidson.py:493-500 (_gen_x0): x0[k] = sum_i v[i, k] xs[i]."""
    space, nroots = v.shape
    x0 = np.einsum("c,x->cx", v[space - 1], np.asarray(xs[space - 1]))
    for i in reversed(range(space - 1)):
        xi = np.asarray(xs[i])
        for k in range(nroots):
            x0[k] += v[i, k] * xi
    return x0


def davidson(aop, x0, precond, tol=1e-12, max_cycle=50, max_space=12, lindep=1e-14, nroots=1):
    """Lowest `nroots` eigenpairs of a Hermitian operator given as a matvec.

    Reference: renormalizer/lib/davidson/davidson.py:73-148 (davidson) and :151-455 (davidson1),
    default switches (SORT_EIG_BY_SIMILARITY=False, follow_state=False, in-core, no pick).
    Convergence per root: |de| < tol and |residual| < sqrt(tol).
    Returns (e, c) -- scalars / 1-D array for nroots == 1, lists otherwise.
    """
    toloose = np.sqrt(tol)
    if isinstance(x0, np.ndarray) and x0.ndim == 1:
        x0 = [x0]
    max_space = max_space + (nroots - 1) * 3
    heff = None
    fresh_start = True
    e = 0
    v = None
    conv = [False] * nroots

    for icyc in range(max_cycle):
        if fresh_start:
            xs, ax = [], []
            space = 0
            xt = _orthonormalise(x0, lindep)
            if len(xt) == 0:
                raise RuntimeError("davidson: empty or linearly dependent initial guess")
            x0 = None
        elif len(xt) > 1:
            xt = _orthonormalise(xt, lindep)[:40]

        axt = [aop(x) for x in xt]
        xs.extend(xt)
        ax.extend(axt)
        rnow = len(xt)
        head, space = space, space + rnow
        if heff is None:
            dtype = np.result_type(axt[0], xt[0])
            heff = np.empty((max_space + nroots, max_space + nroots), dtype=dtype)

        elast, vlast, conv_last = e, v, conv
        # Rayleigh matrix: new rows/columns only (davidson.py:56-70, _fill_heff_hermitian)
        for ip, i in enumerate(range(head, space)):
            for jp, j in enumerate(range(head, i)):
                heff[i, j] = np.dot(xt[ip].conj(), axt[jp])
                heff[j, i] = heff[i, j].conj()
            heff[i, i] = np.dot(xt[ip].conj(), axt[ip]).real
        for i in range(head):
            for jp, j in enumerate(range(head, space)):
                heff[j, i] = np.dot(xt[jp].conj(), ax[i])
                heff[i, j] = heff[j, i].conj()
        xt = axt = None

        w, v = scipy.linalg.eigh(heff[:space, :space])
        e = w[:nroots]
        v = v[:, :nroots]
        x0 = _combine(v, xs)
        ax0 = _combine(v, ax)

        # reorder last energies by overlap (davidson.py:503-521, _sort_elast)
        if not fresh_start:
            hd = vlast.shape[0]
            idx = np.argmax(abs(np.dot(v[:hd].conj().T, vlast)), axis=1)
            elast = [elast[i] for i in idx]
            conv_last = [conv_last[i] for i in idx]
        de = e - elast
        dx_norm, xt, conv = [], [], [False] * nroots
        for k, ek in enumerate(e):
            r = ax0[k] - ek * x0[k]
            xt.append(r)
            dx_norm.append(np.sqrt(np.dot(r.conj(), r).real))
            conv[k] = abs(de[k]) < tol and dx_norm[k] < toloose
        ax0 = None
        if all(conv):
            break

        if any((not conv[k]) and n ** 2 > lindep for k, n in enumerate(dx_norm)):
            keep = [(not conv[k]) and dx_norm[k] ** 2 > lindep for k in range(len(e))]
        else:
            keep = [dx_norm[k] ** 2 > lindep for k in range(len(e))]
        new = []
        for k in range(len(e)):
            if keep[k]:
                t = precond(xt[k], e[0], x0[k])
                t *= 1 / np.sqrt(np.dot(t.conj(), t).real)
                new.append(t)
        xt = new
        for i in range(space):
            xi = np.asarray(xs[i])
            for t in xt:
                t -= xi * np.dot(xi.conj(), t)
        new = []
        for t in xt:
            nrm = np.sqrt(np.dot(t.conj(), t).real)
            if nrm ** 2 > lindep:
                new.append(t * (1 / nrm))
        xt = new
        if len(xt) == 0:
            break
        fresh_start = space + nroots > max_space

    x0 = [x for x in x0]
    if nroots == 1:
        return e[0], x0[0]
    return e, x0

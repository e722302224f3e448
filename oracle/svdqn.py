"""Oracle: quantum-number blocked SVD / QR bond decomposition and basis selection."""
import numpy as np
import scipy.linalg


def add_outer(a, b):
    """Outer sum over all but the last (qn component) axis.  Reference: mps/svd_qn.py:302-310."""
    assert a.shape[-1] == b.shape[-1]
    sa, sb = a.shape[:-1], b.shape[:-1]
    return a.reshape(sa + (1,) * len(sb) + (-1,)) + b.reshape((1,) * len(sa) + sb + (-1,))


def get_qn_mask(qnmat, qntot):
    """Reference: mps/svd_qn.py:313-314."""
    return np.all(qnmat == np.array(qntot), axis=-1)


def _completed_svd(a, full_matrices, opt_full_matrices, rng_rand):
    """SVD; for very unbalanced blocks under full_matrices add only min(m, n) extra random
    orthonormal vectors.  Reference: mps/svd_qn.py:13-66 (optimized_svd, add_orthonormal_basis).
    """
    m, n = a.shape
    if not full_matrices:
        opt_full_matrices = False
    opt = opt_full_matrices and not (1 / 3 < m / n < 3)
    try:
        u, s, vt = scipy.linalg.svd(a, full_matrices=full_matrices and not opt,
                                    lapack_driver="gesdd")
    except scipy.linalg.LinAlgError:
        u, s, vt = scipy.linalg.svd(a, full_matrices=full_matrices and not opt,
                                    lapack_driver="gesvd")
    if not opt:
        return u, s, vt

    def extend(q):
        rows, cols = q.shape
        assert 2 * cols < rows
        x = rng_rand(rows, cols)
        x = x - q @ (q.T.conj() @ x)
        extra, _ = scipy.linalg.qr(x, mode="economic")
        return np.concatenate([q, extra], axis=1)

    if m < n:
        vt = extend(vt.T).T
    else:
        u = extend(u)
    return u, s, vt


def _scatter_rows(indices, block, nrows):
    """Reference: mps/svd_qn.py:87-94 (blockrecover)."""
    out = np.zeros((nrows, block.shape[1]), dtype=block.dtype)
    out[indices, :] = block
    return out


def svd_qn(coef_array, qnbigl, qnbigr, qntot, QR=False, system=None, full_matrices=True,
           opt_full_matrices=True, rng_rand=None):
    """Block-wise SVD (or QR / RQ) of a centre tensor, one block per conserved quantum number.

    Reference: renormalizer/mps/svd_qn.py:97-246.  Returns the same tuples:
      SVD: (U, S_u, qnl_new, V, S_v, qnr_new);  QR: (U, qnl_new, V, qnr_new), V = (R factor).T
    """
    if rng_rand is None:
        rng_rand = np.random.rand
    nl = int(np.prod(qnbigl.shape[:-1]))
    nr = int(np.prod(qnbigr.shape[:-1]))
    mat = coef_array.reshape(nl, nr)
    qn_size = len(qntot)
    lqn = qnbigl.reshape(-1, qn_size)
    rqn = qnbigr.reshape(-1, qn_size)

    u_nz, u_z, v_nz, v_z, s_nz, su_z, sv_z = [], [], [], [], [], [], []
    ql_nz, ql_z, qr_nz, qr_z = [], [], [], []
    for ql in set([tuple(t) for t in lqn]):
        qr = qntot - ql
        rset = np.where(get_qn_mask(rqn, qr))[0]
        if len(rset) == 0:
            continue
        lset = np.where(get_qn_mask(lqn, ql))[0]
        block = mat[np.ix_(lset, rset)]
        dim = min(block.shape)
        if not QR:
            bu, bs, bvt = _completed_svd(block, full_matrices, opt_full_matrices, rng_rand)
            s_nz.append(bs)
        else:
            mode = "full" if full_matrices else "economic"
            if system == "R":
                bu, bvt = scipy.linalg.rq(block, mode=mode)
            elif system == "L":
                bu, bvt = scipy.linalg.qr(block, mode=mode)
            else:
                raise ValueError("system must be 'L' or 'R' for QR")
        bv = bvt.T
        u_nz.append(_scatter_rows(lset, bu[:, :dim], nl))
        ql_nz += [ql] * dim
        v_nz.append(_scatter_rows(rset, bv[:, :dim], nr))
        qr_nz += [tuple(qr)] * dim
        if full_matrices:
            u_z.append(_scatter_rows(lset, bu[:, dim:], nl))
            ql_z += [ql] * (bu.shape[1] - dim)
            su_z.append(np.zeros(bu.shape[1] - dim))
            v_z.append(_scatter_rows(rset, bv[:, dim:], nr))
            qr_z += [tuple(qr)] * (bv.shape[1] - dim)
            sv_z.append(np.zeros(bv.shape[1] - dim))
    if len(u_nz) + len(u_z) == 0:
        raise ValueError("Invalid quantum number")
    u = np.concatenate(u_nz + u_z, axis=1)
    v = np.concatenate(v_nz + v_z, axis=1)
    qnl_new = ql_nz + ql_z
    qnr_new = qr_nz + qr_z
    if QR:
        return u, qnl_new, v, qnr_new
    su = np.concatenate(s_nz + su_z)
    sv = np.concatenate(s_nz + sv_z)
    if not full_matrices:
        order = np.argsort(su)[::-1]
        u, v = u[:, order], v[:, order]
        su = sv = su[order]
        qnl_new = np.array(qnl_new)[order].tolist()
        qnr_new = np.array(qnr_new)[order].tolist()
    return u, su, qnl_new, v, sv, qnr_new


def eigh_qn(dm, qnbigl, qnbigr, qntot, system):
    """Block diagonalisation of the averaged reduced density matrix of the multi-state algorithm.
    Reference: renormalizer/mps/svd_qn.py:243-302.  Returns (U, sqrt(eigenvalues), new qn)."""
    assert system in ("L", "R")
    qnbig, comp = (qnbigl, qnbigr) if system == "L" else (qnbigr, qnbigl)
    qn_size = len(qntot)
    localqn = qnbig.reshape(-1, qn_size)
    n = len(localqn)
    us, ss, new_qn = [], [], []
    for nl in set([tuple(t) for t in localqn]):
        nr = qntot - nl
        if np.sum(get_qn_mask(comp, nr)) == 0:
            continue
        lset = np.where(get_qn_mask(localqn, nl))[0]
        block = dm.reshape(n, n)[np.ix_(lset, lset)]
        s2, bu = scipy.linalg.eigh(block)
        s2[s2 < 0] = 0
        ss.append(np.sqrt(s2))
        us.append(_scatter_rows(lset, bu, n))
        new_qn += [nl] * len(lset)
    return np.concatenate(us, axis=1), np.concatenate(ss), new_qn


def select_basis(vset, sset, qnlist, compset, mmax, percent=0):
    """Pick the retained renormalised basis: `percent` of it evenly from each quantum-number
    sector, the rest by singular value.  Reference: renormalizer/mps/lib.py:265-335.
    Returns (ms, dim, qn, compms) with compms columns scaled by their singular values.
    """
    qnlist = [tuple(q) for q in qnlist]
    sectors = set(qnlist)
    pool = {i: (qnlist[i], sset[i]) for i in range(len(qnlist))}

    def take_from_sector(qn, n):
        members = sorted(((i, v) for i, v in pool.items() if v[0] == qn),
                         key=lambda x: x[1][1], reverse=True)
        chosen = [i for i, _ in members[:min(n, len(members))]]
        for i in chosen:
            del pool[i]
        return chosen

    nbasis = min(len(pool), mmax)
    picked = []
    if percent != 0:
        per_sector = int(nbasis * percent / len(sectors))
        for qn in sectors:
            picked += take_from_sector(qn, per_sector)
    rest = nbasis - len(picked)
    ranked = sorted(pool.items(), key=lambda x: x[1][1], reverse=True)
    picked += [i for i, _ in ranked[:rest]]
    assert len(picked) == len(set(picked))

    dim = len(picked)
    ms = np.zeros((vset.shape[0], dim), dtype=vset.dtype)
    compms = None if compset is None else np.zeros((compset.shape[0], dim), dtype=compset.dtype)
    qn_out = []
    for j, i in enumerate(picked):
        ms[:, j] = vset[:, i]
        if compset is not None and i < compset.shape[1]:
            compms[:, j] = compset[:, i] * sset[i]
        qn_out.append(qnlist[i])
    return ms, dim, np.array(qn_out), compms

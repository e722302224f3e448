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
import scipy.linalg


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
    # number of basis states of sites i+1 .. n-1 per quantum number: a bond sector cannot hold more
    # states than its complement offers
    right_count = [None] * n
    acc = {tuple([0] * qn_size): 1}
    for i in range(n - 1, 0, -1):
        nxt = {}
        for t in np.asarray(sigmaqn[i]).reshape(-1, qn_size):
            for k, c in acc.items():
                key = tuple(np.array(k) + t)
                if np.all(np.array(key) <= qntot):
                    nxt[key] = min(nxt.get(key, 0) + c, 1 << 40)
        acc = nxt
        right_count[i - 1] = acc
    for i in range(n - 1):
        sq = np.asarray(sigmaqn[i])
        big = (qn[i][:, None, :] + sq[None, :, :]).reshape(-1, qn_size)
        cols, col_qn = [], []
        sectors = sorted(set(map(tuple, np.unique(big, axis=0))))
        sectors = [s for s in sectors if np.all(np.array(s) <= qntot)]
        # also drop sectors from which qntot cannot be reached by the remaining sites
        rest_max = np.sum([np.asarray(q).max(axis=0) for q in sigmaqn[i + 1:]], axis=0)
        sectors = [s for s in sectors if np.all(np.array(s) + rest_max >= qntot)]
        cap = {s: right_count[i].get(tuple(qntot - np.array(s)), 0) for s in sectors}
        sectors = [s for s in sectors if cap[s] > 0]
        members = {s: np.where(np.all(big == np.array(s), axis=1))[0] for s in sectors}
        # spread the retained states evenly over the sectors, then fill up
        total = sum(len(v) for v in members.values())
        target = min(m_max, total)
        room = {s: min(len(members[s]), cap[s]) for s in sectors}
        target = min(target, sum(room.values()))
        quota = {s: min(room[s], target // len(sectors)) for s in sectors}
        left = target - sum(quota.values())
        while left > 0:
            progressed = False
            for s in sectors:
                if left > 0 and quota[s] < room[s]:
                    quota[s] += 1
                    left -= 1
                    progressed = True
            if not progressed:
                break
        blocks = {}
        for s in sectors:
            idx = members[s]
            if quota[s] == 0:
                blocks[s] = (idx, np.zeros((len(idx), 0)))
                continue
            # orthonormal columns; a tall Gaussian block is well conditioned, so two rounds of
            # Cholesky QR (BLAS-3, on the transposed block so that the products are row-major) are as
            # good as Householder QR and several times faster
            qt = rng.standard_normal((quota[s], len(idx)))
            for _ in range(2):
                qt = scipy.linalg.solve_triangular(np.linalg.cholesky(qt @ qt.T), qt, lower=True)
            q = qt.T
            blocks[s] = (idx, q)
        mt = np.zeros((len(big), sum(quota.values())))
        c0 = 0
        for s in sectors:
            idx, q = blocks[s]
            mt[idx, c0:c0 + q.shape[1]] = q
            col_qn += [s] * q.shape[1]
            c0 += q.shape[1]
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


# ---------------------------------------------------------------------------------------------
# Ab initio (quantum chemistry) Hamiltonians in spin orbitals, renormalizer/model/h_qc.py:146-221:
#     H = sum_pq h1e[p,q] a+_p a_q + sum_pqrs h2e[p,q,r,s] a+_p a+_q a_r a_s
# after the Jordan-Wigner transformation of h_qc.py:150-159 (a+_j = Z_0 .. Z_{j-1} sigma-_j with
# index 1 = occupied), one half-spin site per spin orbital, quantum numbers (N_alpha, N_beta).
# The reference builds the MPO symbolically (bipartite-graph algorithm, mpo.py); here the operator
# sum is turned into an MPO numerically: site by site, the coefficient matrix between the left
# bond states extended by the local operator and the distinct remaining operator strings is
# rank-factorised (quantum-number block by block), which yields the same minimal bond dimensions.
# ---------------------------------------------------------------------------------------------
_JW_MATS = [np.eye(2), np.diag([1.0, -1.0]), np.diag([1.0], k=1), np.diag([1.0], k=-1),
            np.diag([1.0, 0.0]), np.diag([0.0, 1.0])]          # I, Z, sigma+ (a), sigma- (a+), 1-n, n
_JW_I, _JW_Z, _JW_P, _JW_M = 0, 1, 2, 3


def _jw_table():
    """Multiplication table of the local alphabet: mats[i] @ mats[j] = sign * mats[k] (sign 0: zero)."""
    nsym = len(_JW_MATS)
    idx = np.zeros((nsym, nsym), dtype=np.int8)
    sgn = np.zeros((nsym, nsym), dtype=np.int8)
    for i in range(nsym):
        for j in range(nsym):
            prod = _JW_MATS[i] @ _JW_MATS[j]
            for k in range(nsym):
                for s in (1, -1):
                    if np.array_equal(prod, s * _JW_MATS[k]):
                        idx[i, j], sgn[i, j] = k, s
    return idx, sgn


def qc_operator_strings(h1e, h2e):
    """Operator strings of the Jordan-Wigner transformed Hamiltonian: (strings, coefs) with
    strings[t, l] the index into the local alphabet (I, Z, sigma+, sigma-, 1-n, n) of term t on
    spin orbital l; equal strings are merged."""
    n = h1e.shape[0]
    tidx, tsgn = _jw_table()
    sites = np.arange(n)[None, :]
    all_str, all_c = [], []
    for ints, kinds in ((h1e, (_JW_M, _JW_P)), (h2e, (_JW_M, _JW_M, _JW_P, _JW_P))):
        orb = np.argwhere(ints != 0)
        if len(orb) == 0:
            continue
        coef = ints[tuple(orb.T)].astype(float)
        cur = np.zeros((len(orb), n), dtype=np.int8)
        sign = np.ones(len(orb), dtype=np.int64)
        for f, kind in enumerate(kinds):
            j = orb[:, f:f + 1]
            fac = np.where(sites < j, _JW_Z, np.where(sites == j, kind, _JW_I)).astype(np.int8)
            sign = sign * np.prod(tsgn[cur, fac].astype(np.int64), axis=1)
            cur = tidx[cur, fac]
        keep = sign != 0
        all_str.append(cur[keep])
        all_c.append(coef[keep] * sign[keep])
    strings = np.concatenate(all_str)
    coefs = np.concatenate(all_c)
    uniq, inv = np.unique(strings, axis=0, return_inverse=True)
    summed = np.zeros(len(uniq))
    np.add.at(summed, inv.reshape(-1), coefs)
    keep = np.abs(summed) > 1e-15 * np.abs(summed).max()
    return uniq[keep], summed[keep]


def qc_sigmaqn(norbs):
    """h_qc.py:209-216: even spin orbitals carry (1, 0) when occupied, odd ones (0, 1)."""
    return [np.array([[0, 0], [1, 0]]) if i % 2 == 0 else np.array([[0, 0], [0, 1]]) for i in range(norbs)]


def operator_sum_mpo(strings, coefs, site_mats, site_qn, tol=1e-13):
    """MPO site tensors W[b, up, down, f] of sum_t coefs[t] prod_l site_mats[l][strings[t, l]].

    site_mats[l] is the local alphabet of site l (array nsym_l x d_l x d_l), site_qn[l][s, :] the
    change of the conserved quantum numbers produced by symbol s; every term must conserve them.
    Returns (sites, bond quantum numbers)."""
    strings = np.asarray(strings)
    nterm, n = strings.shape
    nq = np.asarray(site_qn[0]).shape[1]
    suf, A = strings, np.asarray(coefs, dtype=float)[None, :].copy()
    state_qn = np.zeros((1, nq), dtype=int)
    sites, bond_qn = [], [state_qn]
    smax = 0.0
    for i in range(n):
        mats, sym_qn = np.asarray(site_mats[i]), np.asarray(site_qn[i])
        ops = suf[:, 0].astype(int)
        ns = A.shape[0]
        if i == n - 1:
            w = np.zeros((ns, mats.shape[1], mats.shape[2], 1))
            for u, o in enumerate(ops):
                w[:, :, :, 0] += A[:, u, None, None] * mats[o][None]
            sites.append(w)
            bond_qn.append(np.zeros((1, nq), dtype=int))
            break
        urest, inv = np.unique(suf[:, 1:], axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        # rows (state, op) grouped by their quantum number; each group only meets the columns its
        # (state, op) pairs point to
        groups = {}
        for o in np.unique(ops):
            cols_o = np.nonzero(ops == o)[0]
            live = np.nonzero(np.abs(A[:, cols_o]).max(axis=1) > 0)[0]
            for s in live:
                groups.setdefault(tuple(state_qn[s] + sym_qn[o]), []).append((s, o, cols_o))
        q_cols, q_rows, r_blocks, new_qn = [], [], [], []
        for g in sorted(groups):
            members = groups[g]
            cols = np.unique(np.concatenate([inv[c] for _, _, c in members]))
            pos = np.full(len(urest), -1)
            pos[cols] = np.arange(len(cols))
            sub = np.zeros((len(members), len(cols)))
            for r, (s, o, cols_o) in enumerate(members):
                sub[r, pos[inv[cols_o]]] = A[s, cols_o]
            # rank factorisation sub = Q R through the LQ factorisation of the (short, wide) block
            if sub.shape[1] > sub.shape[0]:
                qt, lt = np.linalg.qr(sub.T)              # sub = lt.T @ qt.T
                u, sv, vh = np.linalg.svd(lt.T, full_matrices=False)
                vh = vh @ qt.T
            else:
                u, sv, vh = np.linalg.svd(sub, full_matrices=False)
            smax = max(smax, sv[0] if len(sv) else 0.0)
            k = int(np.count_nonzero(sv > tol * smax))
            if k == 0:
                continue
            q_rows.append(members)
            q_cols.append(u[:, :k])
            r_blocks.append((cols, sv[:k, None] * vh[:k]))
            new_qn += [g] * k
        ktot = len(new_qn)
        w = np.zeros((ns, mats.shape[1], mats.shape[2], ktot))
        a_new = np.zeros((ktot, len(urest)))
        k0 = 0
        for members, q, (cols, r) in zip(q_rows, q_cols, r_blocks):
            k = q.shape[1]
            for row, (s, o, _) in enumerate(members):
                w[s, :, :, k0:k0 + k] += mats[o][:, :, None] * q[row][None, None, :]
            a_new[k0:k0 + k, cols] = r
            k0 += k
        sites.append(w)
        state_qn = np.array(new_qn, dtype=int).reshape(ktot, -1)
        bond_qn.append(state_qn)
        suf, A = urest, a_new
    return sites, bond_qn


def qc_mpo(h1e, h2e, tol=1e-13):
    """MPO of the ab initio Hamiltonian (spin orbitals, Jordan-Wigner) with (N_alpha, N_beta) bond
    quantum numbers; the nuclear repulsion is not included (as in the reference's model)."""
    n = h1e.shape[0]
    strings, coefs = qc_operator_strings(np.asarray(h1e), np.asarray(h2e))
    mats = np.stack(_JW_MATS)
    qn_a = np.zeros((len(_JW_MATS), 2), dtype=int)
    qn_a[_JW_M], qn_a[_JW_P] = (1, 0), (-1, 0)         # sigma- creates, sigma+ annihilates
    qn_b = qn_a[:, ::-1].copy()
    return operator_sum_mpo(strings, coefs, [mats] * n, [qn_a if i % 2 == 0 else qn_b for i in range(n)], tol)


def exciton_phonon_mpo(energies, jmat, omegas, couplings, nlevels, tol=1e-13):
    """Frenkel-Holstein Hamiltonian with long-range excitonic couplings (example/fmo.py:45-52,
    model.py:236-345, HolsteinModel scheme 3-like: any J_ij):
        H = sum_i e_i a+_i a_i + sum_{i!=j} J_ij a+_i a_j
            + sum_{i,k} w_k b+_ik b_ik + sum_{i,k} g_k w_k a+_i a_i (b+_ik + b_ik)
    sites [e_1, ph_11 .. ph_1K, e_2, ...]; one conserved exciton number.  Returns (sites, bond qn)."""
    nmol, nmode = len(energies), len(omegas)
    adag = np.array([[0.0, 0.0], [1.0, 0.0]])
    e_mats = np.stack([np.eye(2), adag, adag.T, adag @ adag.T])           # I, a+, a, n
    e_qn = np.array([[0], [1], [-1], [0]])
    idn, num, x = _boson_ops(nlevels)
    p_mats = np.stack([idn, num, x])
    p_qn = np.zeros((3, 1), dtype=int)
    n = nmol * (1 + nmode)
    pos_e = [i * (1 + nmode) for i in range(nmol)]
    terms, coefs = [], []

    def add(c, ops):
        if c == 0:
            return
        row = np.zeros(n, dtype=np.int8)
        for site, sym in ops:
            row[site] = sym
        terms.append(row)
        coefs.append(c)
    for i in range(nmol):
        add(energies[i], [(pos_e[i], 3)])
        for j in range(nmol):
            if i != j:
                add(jmat[i][j], [(pos_e[i], 1), (pos_e[j], 2)])
        for k in range(nmode):
            add(omegas[k], [(pos_e[i] + 1 + k, 1)])
            add(couplings[k] * omegas[k], [(pos_e[i], 3), (pos_e[i] + 1 + k, 2)])
    site_mats = [e_mats if l in pos_e else p_mats for l in range(n)]
    site_qn = [e_qn if l in pos_e else p_qn for l in range(n)]
    return operator_sum_mpo(np.stack(terms), np.array(coefs), site_mats, site_qn, tol)


def exciton_phonon_sigmaqn(nmol, nmode, nlevels):
    out = []
    for _ in range(nmol):
        out.append(np.array([[0], [1]]))
        out += [np.zeros((nlevels, 1), dtype=int)] * nmode
    return out


def random_mpdm_qn(sigmaqn, qntot, m_max, rng):
    """Random left-canonical density-operator MPS (sites (l, up, ancilla, r)) with the quantum
    numbers of MpDm: the ancilla index carries none (mpdm.py: _get_sigmaqn = add_outer(sigmaqn, 0)).
    Returns (sites, qn, sigmaqn of the (up, ancilla) pairs)."""
    pair = [np.repeat(np.asarray(sq)[:, None, :], len(sq), axis=1) for sq in sigmaqn]
    flat, qn = random_mps_qn([p.reshape(-1, p.shape[-1]) for p in pair], qntot, m_max, rng)
    sites = [s.reshape(s.shape[0], len(sq), len(sq), s.shape[-1]) for s, sq in zip(flat, sigmaqn)]
    return sites, qn, pair


def random_qc_integrals(nspatial, rng, decay=0.35):
    """Synthetic integrals of the shape read_fcidump returns (h_qc.py:15-72): symmetric one-electron
    matrix, two-electron integrals with the 8-fold symmetry of real orbitals, spin-orbital form and
    antisymmetrised as int_to_h does.  Off-diagonal elements decay with the orbital distance so the
    ground state has the area-law structure of a localised-orbital calculation."""
    n = nspatial
    dist = np.abs(np.arange(n)[:, None] - np.arange(n)[None, :])
    h = rng.standard_normal((n, n)) * np.exp(-decay * dist)
    h = 0.5 * (h + h.T) - np.diag(np.linspace(1.0, 0.0, n))
    # (pq|rs) = sum_k L[k,p,q] L[k,r,s] with symmetric L: positive, 8-fold symmetric
    nk = 2 * n
    lvec = rng.standard_normal((nk, n, n)) * np.exp(-decay * dist)[None] / np.sqrt(nk)
    lvec = 0.5 * (lvec + lvec.transpose(0, 2, 1))
    eri = np.einsum("kpq,krs->pqrs", lvec, lvec)
    ns = 2 * n
    p = np.arange(ns)
    sh = np.where((p[:, None] % 2) == (p[None, :] % 2), h[p[:, None] // 2, p[None, :] // 2], 0.0)
    P, Q, R, S = np.meshgrid(p, p, p, p, indexing="ij")
    seri = np.where((P % 2 == S % 2) & (Q % 2 == R % 2), eri[P // 2, S // 2, Q // 2, R // 2], 0.0)
    aseri = np.where((P < Q) & (R < S), seri - seri.transpose(0, 1, 3, 2), 0.0)
    return sh, aseri


def seeded_mps_qn(sigmaqn, qntot, m_max, rng, occupation, noise=1e-3):
    """A product state (physical index occupation[i] on site i) plus `noise` times a random state
    of bond dimension m_max - 1 -- the direct sum of the two bond spaces, as the reference seeds
    its ab initio runs (mps/tests/test_gs.py:131-134: `mps.scale(1e-8) + hf`).  The first sweep then
    starts from a sensible guess instead of a random vector.  Not canonical; returns (sites, qn)."""
    qntot = np.asarray(qntot)
    n, nq = len(sigmaqn), len(qntot)
    rs, rqn = random_mps_qn(sigmaqn, qntot, m_max - 1, rng)
    acc = np.zeros(nq, dtype=int)
    pqn = [acc.copy().reshape(1, nq)]
    for i in range(n):
        acc = acc + np.asarray(sigmaqn[i])[occupation[i]]
        pqn.append(acc.copy().reshape(1, nq))
    assert np.array_equal(acc, qntot), "occupation does not give qntot"
    pqn[-1] = np.zeros((1, nq), dtype=int)
    sites, qn = [], [np.zeros((1, nq), dtype=int)]
    for i in range(n):
        r = rs[i] * (noise if i == 0 else 1.0)
        dl, d, dr = r.shape
        first, last = i == 0, i == n - 1
        t = np.zeros((1 if first else dl + 1, d, 1 if last else dr + 1))
        t[(0 if first else dl), occupation[i], (0 if last else dr)] = 1.0
        if first:
            t[0, :, :dr] = r[0]
        elif last:
            t[:dl, :, 0] = r[:, :, 0]
        else:
            t[:dl, :, :dr] = r
        sites.append(t)
        qn.append(pqn[i + 1] if last else np.concatenate([rqn[i + 1], pqn[i + 1]]))
    return sites, qn

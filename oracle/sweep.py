"""Oracle: MPS container, environments, two sweep drivers (DMRG ground state, TDVP-PS).

NumPy restatement of the reference's sweep path; see oracle/__init__.py for the rules.
"""
import numpy as np
import scipy.linalg

from .contract import hop_apply, hop_diag, env_update
from .svdqn import add_outer, get_qn_mask, svd_qn, select_basis, eigh_qn
from .krylov import expm_krylov
from .davidson import davidson


class Mps:
    """Site tensors (l, d, r) + quantum-number bookkeeping.

    Reference: renormalizer/mps/mp.py:34-80 (MatrixProduct state), mps/mps.py:118.
    """

    def __init__(self, sites, qn, sigmaqn, qntot, qnidx, to_right, coeff=1.0, is_mpo=False):
        self.sites = [np.asarray(s) for s in sites]
        self.qn = [np.asarray(q) for q in qn]
        self.sigmaqn = [np.asarray(s) for s in sigmaqn]
        self.qntot = np.asarray(qntot)
        self.qnidx = int(qnidx)
        self.to_right = bool(to_right)
        self.coeff = coeff
        # an operator handled as a matrix product (mpo.py:297-303): canonicalisation balances the
        # norm between the two factors and the singular values go to the other side (mp.py:258-275)
        self.is_mpo = bool(is_mpo)

    def __len__(self):
        return len(self.sites)

    def copy(self):
        return Mps([s.copy() for s in self.sites], [q.copy() for q in self.qn], self.sigmaqn,
                   self.qntot.copy(), self.qnidx, self.to_right, self.coeff, self.is_mpo)

    def to_complex(self):
        m = self.copy()
        m.sites = [s.astype(np.complex128) for s in m.sites]
        return m

    @property
    def bond_dims(self):
        return [s.shape[0] for s in self.sites] + [self.sites[-1].shape[-1]]

    # -- reference: mp.py:230-243
    def iter_idx_list(self, full, stop_idx=None):
        n = len(self)
        if self.to_right:
            last = stop_idx if stop_idx is not None else (n if full else n - 1)
            return range(self.qnidx, last)
        last = stop_idx if stop_idx is not None else (-1 if full else 0)
        return range(self.qnidx, last, -1)

    # -- reference: mp.py:297-306
    def switch_direction(self):
        if self.to_right:
            self.qnidx = len(self) - 1
            self.to_right = False
        else:
            self.qnidx = 0
            self.to_right = True

    # -- reference: mp.py:159-172
    def move_qnidx(self, dstidx):
        n = len(self)
        for idx in range(self.qnidx + 1, n + 1):
            self.qn[idx] = self.qntot - self.qn[idx]
        for idx in range(n, dstidx, -1):
            self.qn[idx] = self.qntot - self.qn[idx]
        self.qnidx = dstidx

    # -- reference: mp.py:308-352
    def big_qn(self, cidx):
        sq = [self.sigmaqn[i] for i in cidx]
        qnl = self.qn[cidx[0]]
        qnr = self.qn[cidx[-1] + 1]
        if len(cidx) == 1:
            if self.to_right:
                qnbigl, qnbigr = add_outer(qnl, sq[0]), qnr
            else:
                qnbigl, qnbigr = qnl, add_outer(sq[0], qnr)
        else:
            qnbigl, qnbigr = add_outer(qnl, sq[0]), add_outer(sq[1], qnr)
        return qnbigl, qnbigr, add_outer(qnbigl, qnbigr)

    # -- reference: mp.py:890-908 (_push_cano) + mp.py:245-295 (_update_ms, sigma=None, MPS)
    def push_cano(self, idx):
        qnbigl, qnbigr, _ = self.big_qn([idx])
        system = "L" if self.to_right else "R"
        shape = self.sites[idx].shape
        u, qnlset, v, qnrset = svd_qn(self.sites[idx], qnbigl, qnbigr, self.qntot, QR=True,
                                      system=system, full_matrices=False)
        vt = v.T
        m = u.shape[1]
        if self.is_mpo:                               # mp.py:258-267
            if self.to_right:
                nrm = np.linalg.norm(vt)
                u, vt = u * nrm, vt / nrm
            else:
                nrm = np.linalg.norm(u)
                u, vt = u / nrm, vt * nrm
        if self.to_right:
            self.sites[idx + 1] = np.tensordot(vt, self.sites[idx + 1], axes=1)
            self.sites[idx] = u.reshape(shape[:-1] + (m,))
            self.qn[idx + 1] = np.array(qnlset)
            self.qnidx = idx + 1
        else:
            self.sites[idx - 1] = np.tensordot(self.sites[idx - 1], u, axes=1)
            self.sites[idx] = vt.reshape((m,) + shape[1:])
            self.qn[idx] = np.array(qnrset)
            self.qnidx = idx - 1

    # -- reference: mp.py:910-922
    def canonicalise(self):
        idx = None
        for idx in self.iter_idx_list(full=False):
            self.push_cano(idx)
        if (not self.to_right and idx == 1) or (self.to_right and idx == len(self) - 2):
            self.switch_direction()
        return self

    def check_left_canonical(self, atol=1e-8, rtol=1e-5):
        for s in self.sites[:-1]:
            m = s.reshape(-1, s.shape[-1])
            if not np.allclose(m.T.conj() @ m, np.eye(m.shape[1]), rtol=rtol, atol=atol):
                return False
        return True

    def check_right_canonical(self, atol=1e-8, rtol=1e-5):
        for s in self.sites[1:]:
            m = s.reshape(s.shape[0], -1)
            if not np.allclose(m @ m.T.conj(), np.eye(m.shape[0]), rtol=rtol, atol=atol):
                return False
        return True

    # -- reference: mp.py:206-228
    def ensure_left_canonical(self):
        if self.to_right or self.qnidx != len(self) - 1 or not self.check_left_canonical():
            self.move_qnidx(0)
            self.to_right = True
            return self.canonicalise()
        return self

    def ensure_right_canonical(self):
        if (not self.to_right) or self.qnidx != 0 or not self.check_right_canonical():
            self.move_qnidx(len(self) - 1)
            self.to_right = False
            return self.canonicalise()
        return self

    # -- reference: mp.py:933-958 (dot), mp.py:355-372 (mp_norm)
    def dot_conj(self, other):
        """<self|other>"""
        e0 = np.eye(1)
        for a, b in zip(self.sites, other.sites):
            e0 = np.tensordot(e0, b, 1)
            phys = list(range(a.ndim - 1))       # left bond + physical (+ ancilla) indices
            e0 = np.tensordot(e0, a.conj(), (phys, phys)).T
        return complex(e0[0, 0])

    @property
    def mp_norm(self):
        res = self.dot_conj(self).real
        return float(np.sqrt(max(res, 0.0)))

    # -- reference: mps.py:2025-2058 ("mps_only") + mp.py:984-994 (scale at qnidx)
    def normalize_mps_only(self):
        self.sites[self.qnidx] = self.sites[self.qnidx] * (1.0 / self.mp_norm)
        return self

    # -- reference: mps.py:471-525 (expectation through a right environment)
    def expectation(self, mpo):
        r = np.ones((1, 1, 1))
        for i in range(len(self) - 1, -1, -1):
            r = env_update(r, self.sites[i], mpo[i], "R")
        val = complex(r[0, 0, 0])
        return val.real if np.isclose(val.imag, 0) else val


class Environ:
    """Left/right environment store.  Reference: renormalizer/mps/lib.py:12-129."""

    def __init__(self, mps, mpo, domain=None, mps_conj=None):
        """`mps_conj` (optional): the already-conjugated bra state (lib.py:19-31)."""
        self.disk = {}
        self.sentinel = np.ones((1, 1, 1))
        self.disk[("L", -1)] = self.sentinel
        self.disk[("R", len(mps))] = self.sentinel
        for dom in (["L", "R"] if domain is None else [domain]):
            n = len(mps)
            rng = range(0, n - 1) if dom == "L" else range(n - 1, 0, -1)
            t = self.sentinel
            for i in rng:
                t = env_update(t, mps.sites[i], mpo[i], dom,
                               ms_conj=None if mps_conj is None else mps_conj.sites[i])
                self.disk[(dom, i)] = t

    def read(self, domain, idx):
        return self.disk[(domain, idx)]

    def get_lr(self, domain, idx, mps, mpo, method, mps_conj=None):
        if idx < 0 or idx >= len(mps):
            return self.sentinel
        if method == "Enviro":
            return self.read(domain, idx)
        assert method == "System"
        prev = self.read(domain, idx + (-1 if domain == "L" else 1))
        t = env_update(prev, mps.sites[idx], mpo[idx], domain,
                       ms_conj=None if mps_conj is None else mps_conj.sites[idx])
        self.disk[(domain, idx)] = t
        return t


def m_trunc_fixed(sigma, m_max):
    """Reference: utils/configs.py:202-205 (_fixed_m_trunc with a uniform max bond dimension)."""
    return min(int(m_max), len(sigma))


def m_trunc_threshold(sigma, threshold):
    """Reference: utils/configs.py:196-200 (_threshold_m_trunc)."""
    return int(np.sum(sigma / scipy.linalg.norm(sigma) > threshold))


def update_mps(mps, cstruct, cidx, qnbigl, qnbigr, m_max, percent=0.0):
    """Decompose the optimised centre tensor and move the centre one site on.

    Reference: renormalizer/mps/mp.py:651-888 (_update_mps), single-state SVD branch (no OFS).
    """
    system = "L" if mps.to_right else "R"
    multi = isinstance(cstruct, list)
    rotated_c, averaged_ms = [], []
    if not multi:
        u, su, qnlnew, v, sv, qnrnew = svd_qn(cstruct, qnbigl, qnbigr, mps.qntot, system=system)
        if mps.to_right:
            ms, msdim, msqn, compms = select_basis(u, su, qnlnew, v, m_trunc_fixed(su, m_max), percent)
            ms = ms.reshape(list(qnbigl.shape[:-1]) + [msdim])
            compms = np.moveaxis(compms.reshape(list(qnbigr.shape[:-1]) + [msdim]), -1, 0)
        else:
            ms, msdim, msqn, compms = select_basis(v, sv, qnrnew, u, m_trunc_fixed(sv, m_max), percent)
            ms = np.moveaxis(ms.reshape(list(qnbigr.shape[:-1]) + [msdim]), -1, 0)
            compms = compms.reshape(list(qnbigl.shape[:-1]) + [msdim])
    else:
        # state-averaged method (mp.py:780-838): basis from the averaged reduced density matrix
        nl_axes = qnbigl.ndim - 1
        ddm = 0.0
        for c in cstruct:
            if mps.to_right:
                ax = list(range(nl_axes, c.ndim))
            else:
                ax = list(range(nl_axes))
            ddm = ddm + np.tensordot(c, c, axes=(ax, ax))
        ddm = ddm / len(cstruct)
        uset, sset, qnnew = eigh_qn(ddm, qnbigl, qnbigr, mps.qntot, system)
        ms, msdim, msqn, _ = select_basis(uset, sset, qnnew, None, m_trunc_fixed(sset, m_max), percent)
        if mps.to_right:
            ms = ms.reshape(list(qnbigl.shape[:-1]) + [msdim])
            for c in cstruct:
                rotated_c.append(np.tensordot(ms, c, axes=(list(range(nl_axes)), list(range(nl_axes)))))
            compms = rotated_c[0]
        else:
            ms = ms.reshape(list(qnbigr.shape[:-1]) + [msdim])
            for c in cstruct:
                rotated_c.append(np.tensordot(c, ms, axes=(list(range(nl_axes, c.ndim)),
                                                          list(range(qnbigr.ndim - 1)))))
            compms = rotated_c[0]
            ms = np.moveaxis(ms, -1, 0)
    n = len(mps)
    if len(cidx) == 1:
        i = cidx[0]
        mps.sites[i] = ms
        if mps.to_right:
            if i != n - 1:
                averaged_ms = [np.tensordot(c, mps.sites[i + 1], axes=1) for c in rotated_c]
                mps.sites[i + 1] = np.tensordot(compms, mps.sites[i + 1], axes=1)
                mps.qn[i + 1] = msqn
                mps.qnidx = i + 1
            else:
                averaged_ms = [np.tensordot(mps.sites[i], c, axes=1) for c in rotated_c]
                mps.sites[i] = np.tensordot(mps.sites[i], compms, axes=1)
                mps.qnidx = n - 1
        else:
            if i != 0:
                averaged_ms = [np.tensordot(mps.sites[i - 1], c, axes=1) for c in rotated_c]
                mps.sites[i - 1] = np.tensordot(mps.sites[i - 1], compms, axes=1)
                mps.qn[i] = msqn
                mps.qnidx = i - 1
            else:
                averaged_ms = [np.tensordot(c, mps.sites[i], axes=1) for c in rotated_c]
                mps.sites[i] = np.tensordot(compms, mps.sites[i], axes=1)
                mps.qnidx = 0
    else:
        if mps.to_right:
            mps.sites[cidx[0]], mps.sites[cidx[1]] = ms, compms
            mps.qnidx = cidx[1]
        else:
            mps.sites[cidx[1]], mps.sites[cidx[0]] = ms, compms
            mps.qnidx = cidx[0]
        averaged_ms = rotated_c
        mps.qn[cidx[1]] = msqn
    return averaged_ms if multi else None


def _sign_fix(c):
    """Reference: mps/gs.py:372-380."""
    return c / np.sign(c[np.abs(c).argmax()])


def _scatter(c, mask):
    """Reference: mps/lib.py:438-457 (cvec2cmat, one root)."""
    out = np.zeros(mask.shape, dtype=c.dtype)
    out[mask] = c
    return out


class StackedMpo:
    """Sum of Hamiltonians kept as separate MPOs.  Reference: renormalizer/mps/mpo.py:483-494;
    consumed by gs.py:113-114 (one Environ per member) and gs.py:226-241, 321-343, 499-502
    (H_eff and its diagonal are the sums over the members)."""

    def __init__(self, mpos):
        self.mpos = list(mpos)


def dmrg_single_sweep(mps, mpo, environ, method, m_max, percent, last_opt_idx, stats=None, nroots=1,
                      site_filter=None, site_done=None):
    """One DMRG sweep over all sites.  Reference: renormalizer/mps/gs.py:174-304 (omega=None,
    algo="davidson"; nroots > 1 is the state-averaged algorithm).  Returns (micro results
    [(e, cidx)], res_mps) with res_mps a list of Mps when nroots > 1."""
    n = len(mps)
    micro = []
    res_mps = None
    averaged_ms = []
    for imps in mps.iter_idx_list(full=True):
        if method == "2site" and ((mps.to_right and imps == n - 1) or
                                  (not mps.to_right and imps == 0)):
            break
        lmethod, rmethod = ("System", "Enviro") if mps.to_right else ("Enviro", "System")
        if method == "1site":
            lidx, cidx, ridx = imps - 1, [imps], imps + 1
        elif mps.to_right:
            lidx, cidx, ridx = imps - 1, [imps, imps + 1], imps + 2
        else:
            lidx, cidx, ridx = imps - 2, [imps - 1, imps], imps + 1
        members = mpo.mpos if isinstance(mpo, StackedMpo) else [mpo]
        environs = environ if isinstance(environ, list) else [environ]
        lts = [env_i.get_lr("L", lidx, mps, op_i, lmethod) for env_i, op_i in zip(environs, members)]
        rts = [env_i.get_lr("R", ridx, mps, op_i, rmethod) for env_i, op_i in zip(environs, members)]
        cmos = [[op_i[i] for i in cidx] for op_i in members]
        if site_filter is not None and imps not in site_filter:
            # measurement aid (bench.py): pass over this site with the QR alone
            mps.push_cano(imps)
            if site_done is not None:
                site_done(imps, None)
            continue
        qnbigl, qnbigr, qnmat = mps.big_qn(cidx)
        mask = get_qn_mask(qnmat, mps.qntot)
        cshape = mask.shape
        if np.prod(cshape) < 1000:
            # direct diagonalisation.  Reference: gs.py:307-407
            ham = 0
            for ltensor, rtensor, cmo in zip(lts, rts, cmos):
                if len(cidx) == 1:
                    h_i = np.einsum("abc,bdef,lfk->adlcek", ltensor, cmo[0], rtensor, optimize=True)
                    ham = ham + h_i[:, :, :, mask][mask, :]
                else:
                    h_i = np.einsum("abc,bdef,fghj,ljk->adglcehk", ltensor, cmo[0], cmo[1], rtensor,
                                    optimize=True)
                    ham = ham + h_i[:, :, :, :, mask][mask, :]
            w, vec = scipy.linalg.eigh(ham)
            if nroots == 1:
                e, c = w[0], _sign_fix(vec[:, 0])
            else:
                e = w[:nroots]
                c = [_sign_fix(vec[:, i]) for i in range(min(nroots, vec.shape[1]))]
            nhop = 0
        else:
            # Davidson.  Reference: gs.py:410-576
            if nroots == 1:
                if len(cidx) == 1:
                    guess = mps.sites[cidx[0]]
                else:
                    guess = np.tensordot(mps.sites[cidx[0]], mps.sites[cidx[1]], axes=1)
                cguess = [guess[mask]]
            else:
                cguess = []
                for ms in averaged_ms:
                    if len(cidx) == 1:
                        raw = ms
                    elif mps.to_right:
                        raw = np.tensordot(ms, mps.sites[cidx[1]], axes=1)
                    else:
                        raw = np.tensordot(mps.sites[cidx[0]], ms, axes=1)
                    cguess.append(raw[mask])
                guess_dim = int(np.sum(mask))
                cguess.extend([np.random.rand(guess_dim) - 0.5 for _ in range(len(cguess), nroots)])
            hdiag = sum(hop_diag(l, r, c) for l, r, c in zip(lts, rts, cmos))[mask]
            count = [0]

            def hop(x):
                count[0] += 1
                xs = _scatter(x, mask)
                return sum(hop_apply(l, r, c, xs) for l, r, c in zip(lts, rts, cmos))[mask]

            def precond(x, e, *args):
                return x / (hdiag - e + 1e-4)
            e, c = davidson(hop, cguess, precond, max_cycle=100, nroots=nroots)
            if nroots == 1:
                c = _sign_fix(c)
            else:
                c = [_sign_fix(ci) for ci in c]
            nhop = count[0]
        if stats is not None:
            stats.append(nhop)
        if nroots > 1:
            e = np.asarray(e).tolist()
        micro.append((e, cidx))
        if nroots == 1:
            cstruct = _scatter(c, mask)
        else:
            cstruct = [_scatter(ci, mask) for ci in c]
        if cidx == last_opt_idx:
            if nroots == 1:
                res_mps = mps.copy()
                update_mps(res_mps, cstruct, cidx, qnbigl, qnbigr, m_max, percent)
            else:
                res_mps = [mps.copy() for _ in cstruct]
                for r, ci in zip(res_mps, cstruct):
                    update_mps(r, ci, cidx, qnbigl, qnbigr, m_max, percent)
        averaged_ms = update_mps(mps, cstruct, cidx, qnbigl, qnbigr, m_max, percent)
        if site_done is not None:
            site_done(imps, e)
    mps.switch_direction()
    return micro, res_mps


def optimize_mps(mps, mpo, procedure, method="2site", e_rtol=1e-6, e_atol=1e-8, stats=None,
                 micro_out=None, nroots=1):
    """DMRG ground state.  Reference: renormalizer/mps/gs.py:54-171.  `mps` is overwritten.
    Returns (energy per sweep, optimised Mps)."""
    if mps.qnidx == len(mps) - 1:     # is_left_canonical
        mps.ensure_right_canonical()
        env = "R"
    else:
        mps.ensure_left_canonical()
        env = "L"
    if isinstance(mpo, StackedMpo):
        environ = [Environ(mps, item, env) for item in mpo.mpos]       # gs.py:113-114
    else:
        environ = Environ(mps, mpo, env)
    macro = []
    opt_idx = None
    res_mps = None
    for isweep, (m_max, percent) in enumerate(procedure):
        micro, res_mps, = dmrg_single_sweep(mps, mpo, environ, method, int(m_max), percent,
                                            opt_idx, stats, nroots)
        if micro_out is not None:
            micro_out.append(np.array([e for e, _ in micro]))
        opt_e = min(micro)
        macro.append(opt_e[0])
        opt_idx = opt_e[1]
        if isweep > 0 and percent == 0:
            v1, v2 = sorted(macro)[:2]
            if np.allclose(v1, v2, rtol=e_rtol, atol=e_atol):
                break
    assert res_mps is not None
    for r in (res_mps if isinstance(res_mps, list) else [res_mps]):
        r.normalize_mps_only()
        r.ensure_left_canonical()
        r.canonicalise()
    return macro, res_mps


def evolve_tdvp_ps(mps_in, mpo, dt, normalize=True, stats=None):
    """One time step of one-site projector-splitting TDVP (two half sweeps, Krylov local solver).

    Reference: renormalizer/mps/mps.py:1268-1404 (_evolve_tdvp_ps, ivp_solver == "krylov") and
    mps.py:644-662 (evolve -> normalize "mps_only" for real dt).
    """
    mps = mps_in.to_complex() if not np.iscomplex(dt) else mps_in.copy()
    n = len(mps)
    environ = Environ(mps, mpo)
    for _ in range(2):
        for imps in mps.iter_idx_list(full=True):
            system = "L" if mps.to_right else "R"
            l_array = environ.read("L", imps - 1)
            r_array = environ.read("R", imps + 1)
            shape = list(mps.sites[imps].shape)
            w = mpo[imps]
            mps_t, j = expm_krylov(
                lambda y: hop_apply(l_array, r_array, [w], y.reshape(shape)).ravel(),
                -1j * dt / 2, mps.sites[imps].ravel())
            if stats is not None:
                stats.append(j)
            mps_t = mps_t.reshape(shape)
            qnbigl, qnbigr, _ = mps.big_qn([imps])
            u, qnlset, v, qnrset = svd_qn(mps_t, qnbigl, qnbigr, mps.qntot, QR=True,
                                          system=system, full_matrices=False)
            vt = v.T
            if not mps.to_right and imps != 0:
                mps.sites[imps] = vt.reshape([-1] + shape[1:])
                mps.qn[imps] = np.array(qnrset)
                mps.qnidx = imps - 1
                r_array = environ.get_lr("R", imps, mps, mpo, "System")
                su = u.shape
                back, j = expm_krylov(
                    lambda y: hop_apply(l_array, r_array, [], y.reshape(su)).ravel(),
                    1j * dt / 2, u.ravel())
                if stats is not None:
                    stats.append(j)
                mps.sites[imps - 1] = np.tensordot(mps.sites[imps - 1], back.reshape(su),
                                                   axes=(-1, 0))
            elif mps.to_right and imps != n - 1:
                mps.sites[imps] = u.reshape(shape[:-1] + [-1])
                mps.qn[imps + 1] = np.array(qnlset)
                mps.qnidx = imps + 1
                l_array = environ.get_lr("L", imps, mps, mpo, "System")
                sv = vt.shape
                back, j = expm_krylov(
                    lambda y: hop_apply(l_array, r_array, [], y.reshape(sv)).ravel(),
                    1j * dt / 2, vt.ravel())
                if stats is not None:
                    stats.append(j)
                mps.sites[imps + 1] = np.tensordot(back.reshape(sv), mps.sites[imps + 1],
                                                   axes=(1, 0))
            else:
                mps.sites[imps] = mps_t
        mps.switch_direction()
    if normalize:
        # mps.py:657-661: "mps_and_coeff" for imaginary time, "mps_only" for real time; the site
        # tensors are scaled by 1/norm either way (coeff / |coeff| is not tracked here)
        mps.normalize_mps_only()
    return mps


def evolve_adaptive_tdvp_ps(mps_in, mpo, target_t, guess_dt, rtol=5e-4):
    """Step-size control around the one-site integrator by step doubling.  Reference:
    renormalizer/mps/mps.py:46-115 (adaptive_tdvp) + mps.py:644-662 (normalisation in evolve).
    Returns (new Mps, new guess_dt)."""
    p_restart, p_min, p_max = 0.5, 0.1, 2.0
    cur = mps_in
    evolved = 0
    while True:
        rest = target_t - evolved
        dt = guess_dt if abs(guess_dt) < abs(rest) else rest
        half1 = evolve_tdvp_ps(cur, mpo, dt / 2, normalize=False)
        half2 = evolve_tdvp_ps(half1, mpo, dt / 2, normalize=False)
        full = evolve_tdvp_ps(cur, mpo, dt, normalize=False)
        l1, l2, l12 = full.dot_conj(full), half2.dot_conj(half2), full.dot_conj(half2)
        dis = np.sqrt(max((l1 + l2 - l12 - np.conj(l12)).real, 0.0))      # mp.py:1009-1023
        p = (0.75 * rtol / (dis / half2.mp_norm + 1e-30)) ** (1.0 / 3)
        p = min(max(p, p_min), p_max)
        if p < p_restart:
            guess_dt = dt * p
            continue
        evolved += dt
        if np.allclose(evolved, target_t):
            half2.normalize_mps_only()
            return half2, guess_dt
        guess_dt *= p
        cur = half2


def evolve_tdvp_ps2(mps_in, mpo, dt, m_max, normalize=True, stats=None):
    """One time step of two-site projector-splitting TDVP with a fixed maximal bond dimension.

    Reference: renormalizer/mps/mps.py:1407-1517 (_evolve_tdvp_ps2, ivp_solver == "krylov"),
    mp.py:651-888 (_update_mps) and mp.py:890-908 (_push_cano).
    """
    mps = mps_in.to_complex() if not np.iscomplex(dt) else mps_in.copy()
    n = len(mps)
    environ = Environ(mps, mpo)
    for _ in range(2):
        for imps in mps.iter_idx_list(full=False):
            if mps.to_right:
                lidx, cidx0, cidx1, ridx = range(imps - 1, imps + 3)
                cidx2, last_idx = cidx1, n - 2
            else:
                lidx, cidx0, cidx1, ridx = range(imps - 2, imps + 2)
                cidx2, last_idx = cidx0, 1
            l_array = environ.read("L", lidx)
            r_array = environ.read("R", ridx)
            ms2 = np.tensordot(mps.sites[cidx0], mps.sites[cidx1], axes=1)
            shape2 = ms2.shape
            w0, w1 = mpo[cidx0], mpo[cidx1]
            mps_t, j = expm_krylov(
                lambda y: hop_apply(l_array, r_array, [w0, w1], y.reshape(shape2)).ravel(),
                -1j * dt / 2, ms2.ravel())
            if stats is not None:
                stats.append(j)
            qnbigl, qnbigr, _ = mps.big_qn([cidx0, cidx1])
            update_mps(mps, mps_t.reshape(shape2), [cidx0, cidx1], qnbigl, qnbigr, m_max)
            if imps == last_idx:
                continue
            if mps.to_right:
                l_array = environ.get_lr("L", lidx + 1, mps, mpo, "System")
            else:
                r_array = environ.get_lr("R", ridx - 1, mps, mpo, "System")
            ms1 = mps.sites[cidx2]
            shape1 = ms1.shape
            w2 = mpo[cidx2]
            back, j = expm_krylov(
                lambda y: hop_apply(l_array, r_array, [w2], y.reshape(shape1)).ravel(),
                1j * dt / 2, ms1.ravel())
            if stats is not None:
                stats.append(j)
            mps.sites[cidx2] = back.reshape(shape1)
            mps.push_cano(cidx2)
        mps.switch_direction()
    if normalize:
        mps.normalize_mps_only()
    return mps


# ------------------------------------------------------------------------------------------------
# Propagate-and-compress, the default integrator of Mps.evolve
# ------------------------------------------------------------------------------------------------
class Mpo:
    """MPO site tensors with the quantum numbers Mpo.apply needs.  Reference: mps/mpo.py:221-329."""

    def __init__(self, sites, qn, qntot, qnidx):
        self.sites = [np.asarray(s) for s in sites]
        self.qn = [np.asarray(q) for q in qn]
        self.qntot = np.asarray(qntot)
        self.qnidx = int(qnidx)

    def __len__(self):
        return len(self.sites)

    def __getitem__(self, i):
        return self.sites[i]


class CompressSpec:
    """The part of CompressConfig the truncation reads.  Reference: utils/configs.py:128-220."""

    def __init__(self, criteria="threshold", threshold=1e-3, max_bonddim=32):
        self.criteria, self.threshold, self.max_bonddim = criteria, threshold, max_bonddim

    def both(self):
        """mps.py:808-811: a threshold criterion is tightened to `both` while contracting."""
        return CompressSpec("both", self.threshold, self.max_bonddim) if self.criteria == "threshold" else self

    def m_trunc(self, sigma):
        thr = m_trunc_threshold(sigma, self.threshold)
        fix = m_trunc_fixed(sigma, self.max_bonddim)
        return {"threshold": thr, "fixed": fix, "both": min(thr, fix)}[self.criteria]


def mps_scale(mps, val, inplace=False):
    """mp.py:983-994: multiply the site at the quantum-number centre."""
    new = mps if inplace else mps.copy()
    if np.iscomplex(val):
        new.sites = [s.astype(np.complex128) for s in new.sites]
    else:
        val = np.real(val)
    new.sites[new.qnidx] = new.sites[new.qnidx] * val
    return new


def mpo_apply(mpo, mps):
    """mpo @ mps without compression.  Reference: mps/mpo.py:331-389."""
    new = mps.copy()
    for i, (w, a) in enumerate(zip(mpo.sites, mps.sites)):
        if a.ndim == 3:      # "apqb,cqd->acpbd"
            mt = np.moveaxis(np.tensordot(w, a, axes=([2], [1])), 3, 1)
            new.sites[i] = mt.reshape(w.shape[0] * a.shape[0], w.shape[1], w.shape[-1] * a.shape[-1])
        else:                # "apqb,cqrd->acprbd"
            mt = np.moveaxis(np.tensordot(w, a, axes=([2], [1])), [-3, -2], [1, 3])
            new.sites[i] = mt.reshape(w.shape[0] * a.shape[0], w.shape[1], a.shape[2],
                                      w.shape[-1] * a.shape[-1])
    orig = new.qnidx
    new.move_qnidx(mpo.qnidx)
    nq = len(new.qntot)
    new.qn = [add_outer(np.array(qo), np.array(qm)).reshape(-1, nq) for qo, qm in zip(mpo.qn, new.qn)]
    new.qntot = new.qntot + mpo.qntot
    new.move_qnidx(orig)
    return new


def mps_add(a, b):
    """Direct sum of two MPS / MPDM.  Reference: mp.py:374-436 and the coeff rule of
    mps.py:1802-1808 (unequal coefficients are first multiplied into the states, in place)."""
    assert np.all(a.qntot == b.qntot) and len(a) == len(b)
    if not np.allclose(a.coeff, b.coeff):
        mps_scale(a, a.coeff, inplace=True)
        mps_scale(b, b.coeff, inplace=True)
        a.coeff = 1
        b.coeff = 1
    cplx = any(np.iscomplexobj(s) for s in a.sites + b.sites)
    dt = np.complex128 if cplx else np.float64
    n = len(a)
    new = a.copy()
    for i, (x, y) in enumerate(zip(a.sites, b.sites)):
        if i == 0:
            new.sites[i] = np.concatenate([x, y], axis=-1).astype(dt)
        elif i == n - 1:
            new.sites[i] = np.concatenate([x, y], axis=0).astype(dt)
        else:
            t = np.zeros((x.shape[0] + y.shape[0],) + x.shape[1:-1] + (x.shape[-1] + y.shape[-1],), dtype=dt)
            t[:x.shape[0], ..., :x.shape[-1]] = x
            t[x.shape[0]:, ..., x.shape[-1]:] = y
            new.sites[i] = t
    new.move_qnidx(b.qnidx)
    new.to_right = b.to_right
    new.qn = [np.concatenate([q1, q2]) for q1, q2 in zip(new.qn, b.qn)]
    new.qn[0] = np.zeros((1, new.qn[0].shape[1]), dtype=int)
    new.qn[-1] = np.zeros((1, new.qn[0].shape[1]), dtype=int)
    return new


def compress(mps, spec, temp_m_trunc=None):
    """SVD truncation sweep of a canonicalised MPS, in place.  Reference: mp.py:437-511 with
    _update_ms (mp.py:245-295).  `temp_m_trunc` (int or one entry per bond) overrides `spec`."""
    assert mps.qnidx == (0 if mps.to_right else len(mps) - 1)
    system = "L" if mps.to_right else "R"
    for idx in mps.iter_idx_list(full=False):
        shape = mps.sites[idx].shape
        qnbigl, qnbigr, _ = mps.big_qn([idx])
        u, sigma, qnlset, v, sigma, qnrset = svd_qn(mps.sites[idx], qnbigl, qnbigr, mps.qntot,
                                                    system=system, full_matrices=False)
        vt = v.T
        if temp_m_trunc is None:
            m = min(spec.m_trunc(sigma), len(sigma))
        elif np.ndim(temp_m_trunc) > 0:
            m = min(int(temp_m_trunc[idx + 1 if mps.to_right else idx]), len(sigma))
        else:
            m = min(int(temp_m_trunc), len(sigma))
        u, vt, sigma = u[:, :m], vt[:m, :], sigma[:m]
        # mp.py:270-275: the singular values go with the centre -- for an operator, the other way
        sigma_right = mps.to_right != mps.is_mpo
        if mps.to_right:
            if sigma_right:
                vt = sigma[:, None] * vt
            else:
                u = u * sigma[None, :]
            mps.sites[idx + 1] = np.tensordot(vt, mps.sites[idx + 1], axes=1)
            mps.sites[idx] = u.reshape(shape[:-1] + (m,))
            mps.qn[idx + 1] = np.array(qnlset[:m])
            mps.qnidx = idx + 1
        else:
            if sigma_right:
                vt = sigma[:, None] * vt
            else:
                u = u * sigma[None, :]
            mps.sites[idx - 1] = np.tensordot(mps.sites[idx - 1], u, axes=1)
            mps.sites[idx] = vt.reshape((m,) + shape[1:])
            mps.qn[idx] = np.array(qnrset[:m])
            mps.qnidx = idx - 1
    mps.switch_direction()
    return mps


def compressed_sum(terms, spec, batchsize=5, temp_m_trunc=None):
    """lib.py:417-439.  A single term is canonicalised and compressed IN PLACE (and returned)."""
    queue = list(terms)
    if len(queue) == 1:
        return compress(queue[0].canonicalise(), spec, temp_m_trunc)
    while len(queue) != 1:
        batch, queue = queue[:batchsize], queue[batchsize:]
        s = batch[0]
        for t in batch[1:]:
            s = mps_add(s, t)
        queue.append(compress(s.canonicalise(), spec, temp_m_trunc))
    return queue[0]


def evolve_prop_and_compress(mps, mpo, dt, spec, order=4, normalize=True):
    """One step of the propagate-and-compress integrator, fixed step.  Reference: mps.py:796-884
    (Taylor expansion of exp(-i H dt) to `order`, utils/rk.py:28-34; every H^k psi through
    Mpo.contract = apply + canonicalise + compress, mpo.py:415-419) and mps.py:657-661."""
    from math import factorial
    terms = [mps.copy()]
    while len(terms) < order + 1:
        terms.append(compress(mpo_apply(mpo, terms[-1]).canonicalise(), spec.both()))
    for k, t in enumerate(terms):
        mps_scale(t, (-1.0j * dt) ** k / factorial(k), inplace=True)
    new = compressed_sum(terms, spec)
    if normalize:
        new.normalize_mps_only()
    return new


def _contract(mpo, mps, spec):
    """Mpo.contract, mpo.py:415-419: apply + canonicalise + compress with the state's configuration."""
    return compress(mpo_apply(mpo, mps).canonicalise(), spec)


def evolve_pc_tdrk4(mps, mpo_t, dt, spec):
    """Classical RK4 propagate-and-compress step, mps.py:664-698; mpo_t(t) -> Mpo."""
    def stage(state, t):
        return mps_scale(_contract(mpo_t(t), state, spec), -1j)

    def shifted(k, f):
        return compress(mps_add(mps, mps_scale(k, f)).canonicalise(), spec)
    k1 = stage(mps, 0)
    k2 = stage(shifted(k1, 0.5 * dt), 0.5 * dt)
    k3 = stage(shifted(k2, 0.5 * dt), 0.5 * dt)
    k4 = stage(shifted(k3, dt), dt)
    return compressed_sum([mps, mps_scale(k1, dt / 6), mps_scale(k2, 2 * dt / 6), mps_scale(k3, 2 * dt / 6),
                           mps_scale(k4, dt / 6)], spec)


def evolve_pc_tdrk(mps, mpo_t, dt, spec, tableau, order, adaptive=False, guess_dt=None, rtol=5e-4):
    """General explicit Runge-Kutta propagate-and-compress step with the embedded-pair step control,
    mps.py:700-793.  tableau = (a, b, c) of utils/rk.py.  Returns (new state, guess_dt)."""
    a, b, c = tableau
    nstage = len(c)

    def norm(m):
        return abs(m.coeff) * m.mp_norm

    def sub_step(y, tau, t0):
        ks = []
        for i in range(nstage):
            k = compressed_sum([y] + [mps_scale(ks[j], a[i, j] * tau) for j in range(i) if a[i, j] != 0], spec,
                               batchsize=6)
            ks.append(mps_scale(_contract(mpo_t(c[i] * tau + t0), k, spec), -1j))
        new = compressed_sum([y] + [mps_scale(ks[i], b[0, i] * tau) for i in range(nstage) if b[0, i] != 0], spec,
                             batchsize=6)
        if not adaptive:
            return new, 0
        err = None
        for i in range(nstage):
            if not np.allclose(b[0, i], b[1, i]):
                term = mps_scale(ks[i], (b[0, i] - b[1, i]) * tau)
                err = term if err is None else mps_add(err, term)
        return new, norm(err) / norm(new)

    if not adaptive:
        return sub_step(mps, dt, 0)[0], guess_dt
    p_restart, p_min, p_max = 0.5, 0.1, 2.0
    evolved, new = 0, mps

    def min_abs(x, y):
        return x if abs(x) < abs(y) else y
    while True:
        tau = min_abs(guess_dt, dt - evolved)
        new, error = sub_step(new, tau, evolved)       # mps.py:757-759: kept even when judged inaccurate
        p = (rtol / (error + 1e-30)) ** (1 / order[0])
        if p < p_restart:
            guess_dt = tau * max(p_min, p)
        elif np.allclose(tau + evolved, dt):
            return new, min_abs(tau * p, guess_dt)
        else:
            guess_dt *= min(p, p_max)
            evolved += tau


def normalize(mps, kind):
    """mps.py:619-642 / 2025-2058, in place."""
    nrm = mps.mp_norm
    if kind == "mps_only":
        new_coeff = mps.coeff
    elif kind == "mps_and_coeff":
        new_coeff = mps.coeff / np.linalg.norm(mps.coeff)
    elif kind == "mps_norm_to_coeff":
        new_coeff = mps.coeff * nrm
    else:
        raise ValueError(f"kind={kind} is not valid.")
    mps_scale(mps, 1.0 / nrm, inplace=True)
    mps.coeff = new_coeff
    return mps


def bond_dims_exact(mps):
    """mp.py:130-142: bond dimensions of an exact factorisation."""
    p = np.array([float(np.prod(s.shape[1:-1])) for s in mps.sites])
    with np.errstate(over="ignore"):
        d1 = [1] + list(np.cumprod(p))
        d2 = ([1] + list(np.cumprod(p[::-1])))[::-1]
    return np.minimum(d1, d2)


def expand_bond_dimension(mps, hint_mpo, max_bonddim, coef=1e-10):
    """Fill the bond dimension up to `max_bonddim` with states reached through `hint_mpo`.
    Reference: mps.py:1934-2023 with include_ex=False (expand_bond_dimension_general, ex_mps=None).
    The aliasing of the reference is kept: the first `expander` IS `lastone`, compressed in place."""
    spec = CompressSpec("fixed", max_bonddim=max_bonddim)          # unused: every compress is explicit
    max_dims = np.full(len(mps.bond_dims), max_bonddim, dtype=int)
    m_target = np.minimum(max_dims - np.array(mps.bond_dims), bond_dims_exact(mps)).astype(int)
    lastone = mps
    expander_list = []
    expander_dims = np.zeros_like(m_target)
    hint_max = max(hint_mpo.sites[0].shape[0], *(w.shape[-1] for w in hint_mpo.sites))
    while True:
        lastone = normalize(mpo_apply(hint_mpo, lastone), "mps_and_coeff")
        lastone = compress(lastone.canonicalise(), spec, int(np.max(m_target)))
        expander_list.append(lastone)
        expander = compressed_sum(expander_list, spec, temp_m_trunc=m_target)
        if np.all(np.array(expander.bond_dims) >= m_target):
            break
        if np.all(np.array(expander.bond_dims) == expander_dims):
            m_target2 = np.max(m_target - np.array(expander_dims))
            expander2 = compress(mpo_apply(hint_mpo, lastone).canonicalise(), spec, int(np.maximum(m_target2, 1)))
            expander = mps_add(expander, expander2)
            break
        expander_dims = np.array(expander.bond_dims)
        lastone = compress(lastone.canonicalise(), spec, int(np.max(m_target) / hint_max) + 1)
    norm = abs(mps.coeff) * mps.mp_norm
    new = mps_add(mps, mps_scale(expander, coef * norm, inplace=True))
    return normalize(compress(new.canonicalise(), spec, max_dims), "mps_norm_to_coeff")


def mps_distance(a, b):
    """mp.py:1009-1023 (equal coefficients)."""
    l1, l2, l12 = a.dot_conj(a), b.dot_conj(b), a.dot_conj(b)
    return float(np.sqrt(max((l1 + l2 - l12 - np.conj(l12)).real, 0.0)))


def evolve_prop_and_compress_adaptive(mps, mpo, dt, spec, guess_dt, rtol=5e-4, order=5, normalize=True):
    """Propagate-and-compress with adaptive step control.  Reference: mps.py:796-880 (adaptive
    branch: error = distance between the sums to order-1 and to order, p = (rtol / error)^(1/order),
    recursion over the remaining time) and mps.py:657-661.  Returns (new Mps, new guess_dt)."""
    from math import factorial
    p_restart, p_min, p_max = 0.5, 0.1, 2.0

    def step(psi, evolve_dt, guess):
        terms = [psi.copy()]
        while len(terms) < order + 1:
            terms.append(compress(mpo_apply(mpo, terms[-1]).canonicalise(), spec.both()))
        while True:
            h = guess if abs(guess) < abs(evolve_dt) else evolve_dt
            scaled = [mps_scale(t, (-1.0j * h) ** k / factorial(k)) for k, t in enumerate(terms)]
            new1 = compressed_sum(scaled[:-1], spec)
            new2 = compressed_sum([new1, scaled[-1]], spec)
            dis = mps_distance(new1, new2)
            p = (rtol / (dis / new2.mp_norm + 1e-30)) ** (1.0 / order)
            if np.allclose(h, evolve_dt):
                if p < p_restart:
                    guess = h * max(p_min, p)
                else:
                    return new2, (h * p if abs(h * p) < abs(guess) else guess)
            else:
                if p < p_restart:
                    guess = guess * max(p_min, p)
                else:
                    guess = guess * min(p, p_max)
                    return step(new2, evolve_dt - h, guess)

    new, guess = step(mps, dt, guess_dt)
    if normalize:
        new.normalize_mps_only()
    return new, guess


def calc_bond_singular_values(mps_in):
    """Singular values at every bond, padded to a rectangle.  Reference: mps.py:1759-1773
    (copy, ensure_right_canonical, compress(temp_m_trunc=inf, ret_s=True); mp.py:497-511)."""
    mps = mps_in.copy()
    mps.ensure_right_canonical()
    system = "L" if mps.to_right else "R"
    s_list = []
    for idx in mps.iter_idx_list(full=False):
        shape = mps.sites[idx].shape
        qnbigl, qnbigr, _ = mps.big_qn([idx])
        u, sigma, qnlset, v, sigma, qnrset = svd_qn(mps.sites[idx], qnbigl, qnbigr, mps.qntot,
                                                    system=system, full_matrices=False)
        s_list.append(sigma)
        m = len(sigma)
        if mps.to_right:
            mps.sites[idx + 1] = np.tensordot(sigma[:, None] * v.T, mps.sites[idx + 1], axes=1)
            mps.sites[idx] = u.reshape(shape[:-1] + (m,))
            mps.qn[idx + 1] = np.array(qnlset)
            mps.qnidx = idx + 1
        else:
            mps.sites[idx - 1] = np.tensordot(mps.sites[idx - 1], u * sigma[None, :], axes=1)
            mps.sites[idx] = v.T.reshape((m,) + shape[1:])
            mps.qn[idx] = np.array(qnrset)
            mps.qnidx = idx - 1
    width = max(len(x) for x in s_list)
    return np.array([np.pad(x, (0, width - len(x))) for x in s_list])


def calc_bond_entropy(mps):
    """Von Neumann entropy of every bond.  Reference: mps.py:1775-1793, utils/utils.py:41-48."""
    out = []
    for sigma in calc_bond_singular_values(mps):
        p = sigma ** 2
        p = p / p.sum()
        p = p[0 < p]
        out.append(-(p * np.log(p)).sum())
    return np.array(out)


def mps_conj(mps):
    new = mps.copy()
    new.sites = [s.conj() for s in new.sites]
    return new


def variational_compress(state, mpo, sigmaqn_mpo, mpo_to_right, max_bonddim, method="2site",
                         vguess_m=(5, 5), vrtol=1e-5, vprocedure=None):
    """A compressed approximation of mpo @ state by sweeps.  Reference: mp.py:513-650
    (Mpo.contract(algo="variational"), guess=None): the guess is the product of the SVD-compressed
    operator and state; every site update applies H_eff built from environments with the guess as bra
    and `state` as ket to the centre tensor of `state` and decomposes the result into the guess
    (_update_mps); convergence on the relative distance between successive sweeps."""
    spec = CompressSpec("fixed", max_bonddim=max_bonddim)
    op = Mps(mpo.sites, mpo.qn, sigmaqn_mpo, mpo.qntot, mpo.qnidx, mpo_to_right, is_mpo=True)
    c_op = compress(op.copy().canonicalise(), spec, vguess_m[0])
    c_state = compress(state.copy().canonicalise(), spec, vguess_m[1])
    mps = mpo_apply(Mpo(c_op.sites, c_op.qn, c_op.qntot, c_op.qnidx), c_state)
    mps.ensure_left_canonical()
    if vprocedure is None:                                      # configs.py:159-166
        vprocedure = ([[max_bonddim, 1.0], [max_bonddim, 0.7]] if method == "1site" else []) + \
            [[max_bonddim, 0.5], [max_bonddim, 0.3], [max_bonddim, 0.1]] + [[max_bonddim, 0]] * 10
    environ = Environ(state, mpo.sites, "L", mps_conj=mps_conj(mps))
    n = len(mps)
    mps_old = None
    for isweep, (m_max, percent) in enumerate(vprocedure):
        for imps in mps.iter_idx_list(full=True):
            if method == "2site" and ((mps.to_right and imps == n - 1) or (not mps.to_right and imps == 0)):
                break
            lmethod, rmethod = ("System", "Enviro") if mps.to_right else ("Enviro", "System")
            if method == "1site":
                lidx, cidx, ridx = imps - 1, [imps], imps + 1
            elif mps.to_right:
                lidx, cidx, ridx = imps - 1, [imps, imps + 1], imps + 2
            else:
                lidx, cidx, ridx = imps - 2, [imps - 1, imps], imps + 1
            bra = mps_conj(mps)
            ltensor = environ.get_lr("L", lidx, state, mpo.sites, lmethod, mps_conj=bra)
            rtensor = environ.get_lr("R", ridx, state, mpo.sites, rmethod, mps_conj=bra)
            qnbigl, qnbigr, qnmat = mps.big_qn(cidx)
            mask = get_qn_mask(qnmat, mps.qntot)
            cmo = [mpo.sites[i] for i in cidx]
            if method == "1site":
                cms = state.sites[cidx[0]]
            else:
                cms = np.tensordot(state.sites[cidx[0]], state.sites[cidx[1]], axes=1)
            cout = hop_apply(ltensor, rtensor, cmo, cms)
            cout[~mask] = 0
            update_mps(mps, cout, cidx, qnbigl, qnbigr, int(m_max), percent)
        mps.switch_direction()
        if isweep > 0 and percent == 0:
            error = mps_distance(mps, mps_old) / np.sqrt(mps.dot_conj(mps).real)
            if error < vrtol:
                break
        mps_old = mps.copy()
    mps.canonicalise()
    return mps

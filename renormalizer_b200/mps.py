"""Matrix product state resident in HBM, with the sweep-path methods of the reference.

Mirrors the parts of renormalizer/mps/mp.py (MatrixProduct) and renormalizer/mps/mps.py (Mps)
that the DMRG / TDVP-PS sweeps touch: quantum-number bookkeeping (_get_big_qn, move_qnidx),
canonicalisation (_push_cano, canonicalise, ensure_*_canonical), the centre update
(_update_mps), evolve() with the projector-splitting TDVP integrator, expectation, dot, norm.
Site tensors are CUDA tensors (float64 or complex128); quantum numbers are small host arrays.
"""
from typing import List

import numpy as np
import torch

from . import ops
from .backend import backend, asxp, asnumpy
from .configs import CompressConfig, CompressCriteria, EvolveConfig, OptimizeConfig, EvolveMethod
from .hop_expr import hop_expr_dtype
from .krylov import expm_krylov
from .lib import Environ, contract_one_site
from .svd_qn import add_outer, svd_qn, select_basis, eigh_qn, economic_rank, qn_mask_outer


class Mps:
    def __init__(self, sites, qn, sigmaqn, qntot, qnidx, to_right, coeff=1.0):
        self._mp = [asxp(s) for s in sites]
        self.qn = [np.asarray(q) for q in qn]
        self.sigmaqn = [np.asarray(s) for s in sigmaqn]
        self.qntot = np.asarray(qntot)
        self.qnidx = int(qnidx)
        self.to_right = bool(to_right)
        self.coeff = coeff
        self.compress_config = CompressConfig()
        self.optimize_config = OptimizeConfig()
        self.evolve_config = EvolveConfig()
        # True for an operator handled as a matrix product (Mpo.as_matrix_product): canonicalisation
        # balances the norm between the factors and the singular values go the other way
        self.is_mpo = False

    # ------------------------------------------------------------------ construction / transfer
    @classmethod
    def from_numpy(cls, sites, qn, sigmaqn, qntot, qnidx, to_right, coeff=1.0):
        return cls(sites, qn, sigmaqn, qntot, qnidx, to_right, coeff)

    @classmethod
    def without_qn(cls, sites, qnidx=None, to_right=False):
        """MPS of a model with no conserved quantum number (all qn zero)."""
        n = len(sites)
        dims = [s.shape[0] for s in sites] + [sites[-1].shape[-1]]
        qn = [np.zeros((d, 1), dtype=int) for d in dims]
        sigmaqn = [np.zeros((s.shape[1], 1), dtype=int) for s in sites]
        return cls(sites, qn, sigmaqn, np.array([0]), n - 1 if qnidx is None else qnidx, to_right)

    def to_numpy(self):
        return [asnumpy(s) for s in self._mp]

    def load_sites_from_host(self, host_sites, non_blocking=True):
        """Host (pinned) buffers -> the existing device site tensors (same shapes)."""
        for dst, src in zip(self._mp, host_sites):
            dst.copy_(src, non_blocking=non_blocking)

    def store_sites_to_host(self, host_sites, non_blocking=True):
        for src, dst in zip(self._mp, host_sites):
            dst.copy_(src, non_blocking=non_blocking)

    def metacopy(self):
        new = self.__class__.__new__(self.__class__)
        new._mp = [None] * len(self)
        new.qn = [q.copy() for q in self.qn]
        new.sigmaqn = self.sigmaqn
        new.qntot = self.qntot.copy()
        new.qnidx = self.qnidx
        new.to_right = self.to_right
        new.coeff = self.coeff
        new.is_mpo = getattr(self, "is_mpo", False)
        new.compress_config = self.compress_config.copy()
        new.optimize_config = self.optimize_config.copy()
        new.evolve_config = self.evolve_config.copy()
        return new

    def copy(self):
        new = self.metacopy()
        new._mp = [s.clone() for s in self._mp]
        return new

    def to_complex(self, inplace=False):
        new = self if inplace else self.metacopy()
        new._mp = [s.to(torch.complex128) if not s.is_complex() else (s if inplace else s.clone())
                   for s in self._mp]
        return new

    def conj(self):
        new = self.metacopy()
        new._mp = [s.conj().resolve_conj() for s in self._mp]
        return new

    # ------------------------------------------------------------------ container protocol
    def __len__(self):
        return len(self._mp)

    def __getitem__(self, i):
        return self._mp[i]

    def __setitem__(self, i, t):
        self._mp[i] = t if isinstance(t, torch.Tensor) else asxp(t)

    def __iter__(self):
        return iter(self._mp)

    @property
    def site_num(self):
        return len(self._mp)

    @property
    def is_complex(self):
        return any(s.is_complex() for s in self._mp)

    @property
    def dtype(self):
        return torch.complex128 if self.is_complex else torch.float64

    @property
    def bond_dims(self):
        return [s.shape[0] for s in self._mp] + [self._mp[-1].shape[-1]]

    @property
    def pbond_list(self):
        return [s.shape[1] for s in self._mp]

    @property
    def bond_dims_exact(self):
        """mp.py:130-142: bond dimensions of an exact factorisation (physical x ancilla dimension
        per site for a density operator)."""
        p = np.array([float(np.prod(s.shape[1:-1])) for s in self._mp])
        with np.errstate(over="ignore"):
            d1 = [1] + list(np.cumprod(p))
            d2 = ([1] + list(np.cumprod(p[::-1])))[::-1]
        return np.minimum(d1, d2)

    @property
    def total_bytes(self):
        return sum(s.numel() * s.element_size() for s in self._mp)

    def _get_sigmaqn(self, idx):
        return self.sigmaqn[idx]

    # ------------------------------------------------------------------ qn bookkeeping
    def move_qnidx(self, dstidx: int):
        """mp.py:159-172."""
        for idx in range(self.qnidx + 1, self.site_num + 1):
            self.qn[idx] = self.qntot - self.qn[idx]
        for idx in range(self.site_num, dstidx, -1):
            self.qn[idx] = self.qntot - self.qn[idx]
        self.qnidx = dstidx

    def iter_idx_list(self, full: bool, stop_idx: int = None):
        """mp.py:230-243."""
        if self.to_right:
            last = stop_idx if stop_idx is not None else (self.site_num if full else self.site_num - 1)
            return range(self.qnidx, last)
        last = stop_idx if stop_idx is not None else (-1 if full else 0)
        return range(self.qnidx, last, -1)

    def _switch_direction(self):
        """mp.py:297-306."""
        assert self.to_right is not None
        if self.to_right:
            self.qnidx = self.site_num - 1
            self.to_right = False
        else:
            self.qnidx = 0
            self.to_right = True

    def _get_big_qn(self, cidx: List[int], need_mat: bool = True):
        """mp.py:308-352: quantum numbers of the super-L / super-R blocks and of the centre.
        `need_mat=False` skips the (left x right) outer sum, which only the eigensolver's mask
        needs (16.7 M entries for a two-site centre at M = 512); None is returned in its place."""
        if len(cidx) == 2:
            cidx = sorted(cidx)
            assert cidx[0] + 1 == cidx[1]
        elif len(cidx) > 2:
            assert False
        assert self.qnidx in cidx
        sigmaqn = [np.array(self._get_sigmaqn(idx)) for idx in cidx]
        qnl = np.array(self.qn[cidx[0]])
        qnr = np.array(self.qn[cidx[-1] + 1])
        if len(cidx) == 1:
            if self.to_right:
                qnbigl, qnbigr = add_outer(qnl, sigmaqn[0]), qnr
            else:
                qnbigl, qnbigr = qnl, add_outer(sigmaqn[0], qnr)
        else:
            qnbigl, qnbigr = add_outer(qnl, sigmaqn[0]), add_outer(sigmaqn[1], qnr)
        return qnbigl, qnbigr, (add_outer(qnbigl, qnbigr) if need_mat else None)

    # ------------------------------------------------------------------ canonical form
    def check_left_canonical(self, rtol=None, atol=None):
        """mp.py:174-181 / matrix.py:93-103."""
        atol = backend.canonical_atol if atol is None else atol
        rtol = backend.canonical_rtol if rtol is None else rtol
        for s in self._mp[:-1]:
            m = s.reshape(-1, s.shape[-1])
            g = ops.matmul(m.conj().transpose(0, 1).contiguous(), m)
            if not torch.allclose(g, torch.eye(g.shape[0], dtype=g.dtype, device=g.device), rtol=rtol, atol=atol):
                return False
        return True

    def check_right_canonical(self, rtol=None, atol=None):
        atol = backend.canonical_atol if atol is None else atol
        rtol = backend.canonical_rtol if rtol is None else rtol
        for s in self._mp[1:]:
            m = s.reshape(s.shape[0], -1)
            g = ops.matmul(m, m.conj().transpose(0, 1).contiguous())
            if not torch.allclose(g, torch.eye(g.shape[0], dtype=g.dtype, device=g.device), rtol=rtol, atol=atol):
                return False
        return True

    @property
    def is_left_canonical(self):
        return self.qnidx == self.site_num - 1

    @property
    def is_right_canonical(self):
        return self.qnidx == 0

    def ensure_left_canonical(self, rtol=None, atol=None):
        """mp.py:206-216."""
        if self.to_right or self.qnidx != self.site_num - 1 or not self.check_left_canonical(rtol, atol):
            self.move_qnidx(0)
            self.to_right = True
            return self.canonicalise()
        return self

    def ensure_right_canonical(self, rtol=None, atol=None):
        """mp.py:218-228."""
        if (not self.to_right) or self.qnidx != 0 or not self.check_right_canonical(rtol, atol):
            self.move_qnidx(self.site_num - 1)
            self.to_right = False
            return self.canonicalise()
        return self

    def _update_ms(self, idx, u, vt, sigma=None, qnlset=None, qnrset=None, m_trunc=None):
        """mp.py:245-295."""
        if m_trunc is None:
            m_trunc = u.shape[1]
        u = u[:, :m_trunc]
        vt = vt[:m_trunc, :]
        is_mpo = getattr(self, "is_mpo", False)
        if sigma is None:
            if is_mpo:                              # canonicalise an operator: balance the norm
                if self.to_right:
                    nrm = torch.linalg.vector_norm(vt)
                    u, vt = u * nrm, vt / nrm
                else:
                    nrm = torch.linalg.vector_norm(u)
                    u, vt = u / nrm, vt * nrm
        else:
            sig = torch.from_numpy(np.ascontiguousarray(sigma[:m_trunc])).to(u.device).to(u.dtype)
            if self.to_right != is_mpo:             # (mps, to_right) or (mpo, to_left)
                vt = vt * sig[:, None]
            else:
                u = u * sig[None, :]
        shape = self._mp[idx].shape
        pdim = tuple(shape[1:-1])
        if self.to_right:
            self._mp[idx + 1] = ops.tensordot1(vt.contiguous(), self._mp[idx + 1])
            self._mp[idx] = u.contiguous().reshape((shape[0],) + pdim + (m_trunc,))
            if qnlset is not None:
                self.qn[idx + 1] = np.array(qnlset[:m_trunc])
                self.qnidx = idx + 1
        else:
            self._mp[idx - 1] = ops.tensordot1(self._mp[idx - 1], u.contiguous())
            self._mp[idx] = vt.contiguous().reshape((m_trunc,) + pdim + (shape[-1],))
            if qnrset is not None:
                self.qn[idx] = np.array(qnrset[:m_trunc])
                self.qnidx = idx - 1

    def _push_cano(self, idx):
        """mp.py:890-908: move the canonical centre one site on with a QR."""
        qnbigl, qnbigr, _ = self._get_big_qn([idx], need_mat=False)
        system = "L" if self.to_right else "R"
        u, qnlset, v, qnrset = svd_qn(self._mp[idx], qnbigl, qnbigr, self.qntot, QR=True,
                                      system=system, full_matrices=False)
        self._update_ms(idx, u, v.transpose(0, 1), sigma=None, qnlset=qnlset, qnrset=qnrset)

    def canonicalise(self, stop_idx: int = None):
        """mp.py:910-922."""
        if self.to_right:
            assert self.qnidx == 0
        else:
            assert self.qnidx == self.site_num - 1
        idx = None
        for idx in self.iter_idx_list(full=False, stop_idx=stop_idx):
            self._push_cano(idx)
        if (not self.to_right and idx == 1) or (self.to_right and idx == self.site_num - 2):
            self._switch_direction()
        return self

    def compress(self, temp_m_trunc=None, ret_s=False):
        """mp.py:437-511: SVD compression sweep of a canonicalised MPS."""
        if self.to_right:
            assert self.qnidx == 0
        else:
            assert self.qnidx == self.site_num - 1
        if self.compress_config.bonddim_should_set:
            self.compress_config.set_bonddim(len(self) + 1)
        system = "L" if self.to_right else "R"
        s_list = []
        for idx in self.iter_idx_list(full=False):
            qnbigl, qnbigr, _ = self._get_big_qn([idx], need_mat=False)
            u, sigma, qnlset, v, sigma, qnrset = svd_qn(self._mp[idx], qnbigl, qnbigr, self.qntot,
                                                        system=system, full_matrices=False)
            s_list.append(sigma)
            if temp_m_trunc is None:
                m_trunc = self.compress_config.compute_m_trunc(sigma, idx, self.to_right)
            else:
                if isinstance(temp_m_trunc, (list, tuple, np.ndarray)):
                    m_trunc = temp_m_trunc[idx + 1 if self.to_right else idx]
                else:
                    m_trunc = temp_m_trunc
                m_trunc = min(m_trunc, len(sigma))
            self._update_ms(idx, u, v.transpose(0, 1), sigma, qnlset, qnrset, m_trunc)
        self._switch_direction()
        if not ret_s:
            return self
        mx = max(len(s) for s in s_list)
        return self, np.array([np.pad(s, (0, mx - len(s))) for s in s_list])

    # ------------------------------------------------------------------ centre update (DMRG)
    def _update_mps(self, cstruct, cidx, qnbigl, qnbigr, percent=0):
        """mp.py:651-888 without on-the-fly swapping: single-state SVD branch, or -- when `cstruct`
        is a list -- the state-averaged branch (basis from the averaged reduced density matrix,
        mp.py:780-838), which returns the rotated centre tensors of every state."""
        if self.compress_config.ofs is not None:
            raise NotImplementedError("on-the-fly swapping is outside the accelerated path")
        system = "L" if self.to_right else "R"
        if self.compress_config.bonddim_should_set:
            self.compress_config.set_bonddim(len(self) + 1)
        multi = isinstance(cstruct, list)
        rotated_c, averaged_ms = [], []
        if not multi:
            # The reference always asks for full matrices (mp.py:741) and lets select_basis ignore the
            # null-space columns it does not need.  They can only be selected when percent != 0 or
            # when the bond limit exceeds the number of economic singular vectors, so in every other
            # case the completion (random vectors, two projections and a QR per block) is skipped:
            # select_basis returns the same vectors.
            bond = cidx[0] + 1 if self.to_right else cidx[-1]
            need_full = percent != 0 or (
                self.compress_config.criteria is not CompressCriteria.threshold
                and int(self.compress_config.max_dims[bond]) > economic_rank(qnbigl, qnbigr, self.qntot))
            keep = None
            if percent == 0 and self.compress_config.criteria is CompressCriteria.fixed:
                keep = int(self.compress_config.max_dims[bond])
            Uset, SUset, qnlnew, Vset, SVset, qnrnew = svd_qn(cstruct, qnbigl, qnbigr, self.qntot, system=system,
                                                              full_matrices=need_full, keep_hint=keep)
            if self.to_right:
                m_trunc = self.compress_config.compute_m_trunc(SUset, cidx[0], self.to_right)
                ms, msdim, msqn, compms = select_basis(Uset, SUset, qnlnew, Vset, m_trunc, percent=percent)
                ms = ms.contiguous().reshape(list(qnbigl.shape[:-1]) + [msdim])
                compms = compms.transpose(0, 1).contiguous().reshape([msdim] + list(qnbigr.shape[:-1]))
            else:
                m_trunc = self.compress_config.compute_m_trunc(SVset, cidx[-1], self.to_right)
                ms, msdim, msqn, compms = select_basis(Vset, SVset, qnrnew, Uset, m_trunc, percent=percent)
                ms = ms.transpose(0, 1).contiguous().reshape([msdim] + list(qnbigr.shape[:-1]))
                compms = compms.contiguous().reshape(list(qnbigl.shape[:-1]) + [msdim])
        else:
            nl = int(np.prod(qnbigl.shape[:-1]))
            nr = int(np.prod(qnbigr.shape[:-1]))
            mats = [asxp(c).reshape(nl, nr) for c in cstruct]
            ddm = None
            for m2 in mats:
                if self.to_right:
                    term = ops.matmul(m2, m2.transpose(0, 1).contiguous())       # sum over the right indices
                else:
                    term = ops.matmul(m2.transpose(0, 1).contiguous(), m2)       # sum over the left indices
                ddm = term if ddm is None else ddm + term
            ddm = ddm / len(mats)
            Uset, Sset, qnnew = eigh_qn(ddm, qnbigl, qnbigr, self.qntot, system)
            m_trunc = self.compress_config.compute_m_trunc(Sset, cidx[0] if self.to_right else cidx[-1],
                                                           self.to_right)
            ms, msdim, msqn, _ = select_basis(Uset, Sset, qnnew, None, m_trunc, percent=percent)
            ms = ms.contiguous()
            if self.to_right:
                for m2 in mats:      # tensordot(ms, c) over the left indices: (msdim, right...)
                    rotated_c.append(ops.matmul(ms.transpose(0, 1).contiguous(), m2)
                                     .reshape([msdim] + list(qnbigr.shape[:-1])))
                compms = rotated_c[0]
                ms = ms.reshape(list(qnbigl.shape[:-1]) + [msdim])
            else:
                for m2 in mats:      # tensordot(c, ms) over the right indices: (left..., msdim)
                    rotated_c.append(ops.matmul(m2, ms).reshape(list(qnbigl.shape[:-1]) + [msdim]))
                compms = rotated_c[0]
                ms = ms.transpose(0, 1).contiguous().reshape([msdim] + list(qnbigr.shape[:-1]))
        n = self.site_num
        if len(cidx) == 1:
            i = cidx[0]
            self._mp[i] = ms
            if self.to_right:
                if i != n - 1:
                    averaged_ms = [ops.tensordot1(c, self._mp[i + 1]) for c in rotated_c]
                    self._mp[i + 1] = ops.tensordot1(compms, self._mp[i + 1])
                    self.qn[i + 1] = msqn
                    self.qnidx = i + 1
                else:
                    averaged_ms = [ops.tensordot1(self._mp[i], c) for c in rotated_c]
                    self._mp[i] = ops.tensordot1(self._mp[i], compms)
                    self.qnidx = n - 1
            else:
                if i != 0:
                    averaged_ms = [ops.tensordot1(self._mp[i - 1], c) for c in rotated_c]
                    self._mp[i - 1] = ops.tensordot1(self._mp[i - 1], compms)
                    self.qn[i] = msqn
                    self.qnidx = i - 1
                else:
                    averaged_ms = [ops.tensordot1(c, self._mp[i]) for c in rotated_c]
                    self._mp[i] = ops.tensordot1(compms, self._mp[i])
                    self.qnidx = 0
        else:
            if self.to_right:
                self._mp[cidx[0]], self._mp[cidx[1]] = ms, compms
                self.qnidx = cidx[1]
            else:
                self._mp[cidx[1]], self._mp[cidx[0]] = ms, compms
                self.qnidx = cidx[0]
            averaged_ms = rotated_c
            self.qn[cidx[1]] = msqn
        return averaged_ms if multi else None

    # ------------------------------------------------------------------ scalars
    def dot(self, other) -> complex:
        """<conj(self) ... > as in mp.py:933-958: sum_i self_i * other_i (no conjugation)."""
        assert len(self) == len(other)
        e0 = torch.ones((1, 1), dtype=torch.float64, device=self._mp[0].device)
        for mt1, mt2 in zip(self._mp, other):
            t = ops.tensordot1(e0, mt2)                       # (a, d.., r2)
            a = t.reshape(-1, t.shape[-1])                     # ((a,d..), r2)
            b = mt1.reshape(-1, mt1.shape[-1])                 # ((a,d..), r1)
            e0 = ops.matmul(b.transpose(0, 1).contiguous(), a)  # (r1, r2)
        return complex(e0[0, 0].item())

    @property
    def mp_norm(self) -> float:
        """mp.py:355-372."""
        res = self.conj().dot(self).real
        if res < 0:
            assert abs(res) < 1e-8
            res = 0
        return float(np.sqrt(res))

    @property
    def norm(self):
        return abs(self.coeff) * self.mp_norm

    def scale(self, val, inplace=False):
        """mp.py:984-994: scale the tensor at the qn centre."""
        new = self if inplace else self.copy()
        if np.iscomplex(val):
            new.to_complex(inplace=True)
        else:
            val = val.real if hasattr(val, "real") else val
        new._mp[self.qnidx] = new._mp[self.qnidx] * val
        return new

    def normalize(self, kind):
        """mps.py:2025-2058."""
        nrm = self.mp_norm
        if kind == "mps_only":
            new_coeff = self.coeff
        elif kind == "mps_and_coeff":
            new_coeff = self.coeff / abs(self.coeff)
        elif kind == "mps_norm_to_coeff":
            new_coeff = self.coeff * nrm
        else:
            raise ValueError(f"kind={kind} is not valid.")
        self.scale(1.0 / nrm, inplace=True)
        self.coeff = new_coeff
        return self

    def expectation(self, mpo, self_conj=None):
        """mps.py:471-525: <self| mpo |self> through a right environment."""
        if self_conj is None:
            self_conj = self.conj()
        r = torch.ones((1, 1, 1), dtype=torch.float64, device=self._mp[0].device)
        for i in range(len(self) - 1, -1, -1):
            r = contract_one_site(r, self._mp[i], mpo[i], "R", ms_conj=self_conj[i])
        val = complex(r[0, 0, 0].item())
        if np.isclose(val.imag, 0):
            return float(val.real)
        return val

    def calc_bond_singular_values(self) -> np.ndarray:
        """mps.py:1759-1773: the singular values at every bond (rows padded with zeros), from a
        compression sweep that truncates nothing on a right-canonical copy."""
        mps = self.copy()
        mps.ensure_right_canonical()
        _, s_array = mps.compress(temp_m_trunc=np.inf, ret_s=True)
        return s_array

    def calc_bond_entropy(self, s_array=None) -> np.ndarray:
        """mps.py:1775-1793: von Neumann entropy -Tr(rho ln rho) of either block at every bond."""
        if s_array is None:
            s_array = self.calc_bond_singular_values()
        out = []
        for sigma in s_array:
            p = np.asarray(sigma) ** 2
            p = p / p.sum()
            p = p[0 < p]
            out.append(-(p * np.log(p)).sum())
        return np.array(out)

    def calc_entropy(self, entropy_type):
        """mps.py:1689-1732; the reduced-density-matrix entropies need the model's site layout."""
        if entropy_type != "bond":
            raise NotImplementedError(f"entropy type {entropy_type} (mps.py:1716-1727) is outside the sweep path")
        return self.calc_bond_entropy()

    def distance(self, other) -> float:
        """mp.py:1009-1023 with the coeff rule of mps.py:1810-1816."""
        if not np.allclose(self.coeff, other.coeff):
            self.scale(self.coeff, inplace=True)
            other.scale(other.coeff, inplace=True)
            self.coeff = 1
            other.coeff = 1
        l1 = self.conj().dot(self)
        l2 = other.conj().dot(other)
        l12 = self.conj().dot(other)
        d2 = (l1 + l2 - l12 - l12.conjugate()).real
        return float(np.sqrt(d2)) if d2 > 0 else 0.0

    # ------------------------------------------------------------------ time evolution
    def add(self, other):
        """Direct sum of the bond spaces, mp.py:374-436 with the coeff rule of mps.py:1802-1808."""
        if not np.allclose(self.coeff, other.coeff):
            self.scale(self.coeff, inplace=True)
            other.scale(other.coeff, inplace=True)
            self.coeff = 1
            other.coeff = 1
        assert np.all(self.qntot == other.qntot) and self.site_num == other.site_num
        new = self.metacopy()
        new.compress_config.update(self.compress_config)
        dtype = torch.complex128 if (self.is_complex or other.is_complex) else torch.float64
        n = self.site_num
        for i, (a, b) in enumerate(zip(self._mp, other._mp)):
            if i == 0:
                new._mp[i] = torch.cat([a.to(dtype), b.to(dtype)], dim=-1)
            elif i == n - 1:
                new._mp[i] = torch.cat([a.to(dtype), b.to(dtype)], dim=0)
            else:
                t = torch.zeros((a.shape[0] + b.shape[0],) + tuple(a.shape[1:-1]) + (a.shape[-1] + b.shape[-1],),
                                dtype=dtype, device=a.device)
                t[:a.shape[0], ..., :a.shape[-1]] = a
                t[a.shape[0]:, ..., a.shape[-1]:] = b
                new._mp[i] = t
        new.move_qnidx(other.qnidx)
        new.to_right = other.to_right
        new.qn = [np.concatenate([q1, q2]) for q1, q2 in zip(new.qn, other.qn)]
        new.qn[0] = np.zeros((1, new.qn[0].shape[1]), dtype=int)
        new.qn[-1] = np.zeros((1, new.qn[0].shape[1]), dtype=int)
        return new

    def __add__(self, other):
        return self.add(other)

    def variational_compress(self, mpo=None, guess=None):
        """A compressed approximation of mpo @ self by sweeps, mp.py:513-650: environments with the
        guess as bra and `self` as ket, H_eff applied to the centre tensor of `self`, the result
        decomposed into the guess (_update_mps); converged when successive sweeps agree to vrtol.
        `self` is not overwritten, `guess` is.  The bra and ket bonds differ; every tensor is
        zero-padded to one common bond dimension per bond, so the sweeps run on the same (square)
        environment and H_eff kernels as DMRG and the results are sliced back."""
        if mpo is None:
            raise NotImplementedError("Recommend to use svd to compress a single mps/mpo/mpdm.")
        if guess is None:
            nq = len(self.qntot)
            op = mpo.as_matrix_product(nq)
            op.compress_config = self.compress_config.copy()
            compressed_op = op.canonicalise().compress(temp_m_trunc=self.compress_config.vguess_m[0])
            compressed_mps = self.copy().canonicalise().compress(temp_m_trunc=self.compress_config.vguess_m[1])
            from .mpo import Mpo
            guess = Mpo([asnumpy(t) for t in compressed_op], qn=compressed_op.qn, qntot=compressed_op.qntot,
                        qnidx=compressed_op.qnidx).apply(compressed_mps)
        mps = guess
        mps.ensure_left_canonical()
        procedure = mps.compress_config.vprocedure
        method = mps.compress_config.vmethod
        n = self.site_num
        cplx = self.is_complex or mps.is_complex
        dtype = torch.complex128 if cplx else torch.float64
        limit = max(int(c.bond_dim_max_value) if isinstance(c, CompressConfig) else int(c) for c, _ in procedure)
        big = [max(a, b, limit) if 0 < i < n else 1
               for i, (a, b) in enumerate(zip(self.bond_dims, mps.bond_dims))]

        def pad(t, i):
            out = torch.zeros((big[i],) + tuple(t.shape[1:-1]) + (big[i + 1],), dtype=dtype, device=t.device)
            out[:t.shape[0], ..., :t.shape[-1]] = t
            return out

        ket = [pad(t, i) for i, t in enumerate(self._mp)]

        def bra():                                  # the conjugated guess, padded (mp.py:600-607)
            return [pad(t.conj(), i) for i, t in enumerate(mps._mp)]

        environ = Environ(ket, mpo, "L", mps_conj=bra())
        mps_old = None
        for isweep, (compress_config, percent) in enumerate(procedure):
            if isinstance(compress_config, CompressConfig):
                mps.compress_config = compress_config
            elif isinstance(compress_config, (int, np.integer)):
                mps.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=int(compress_config))
            else:
                assert False
            for imps in mps.iter_idx_list(full=True):
                if method == "2site" and ((mps.to_right and imps == n - 1) or ((not mps.to_right) and imps == 0)):
                    break
                lmethod, rmethod = ("System", "Enviro") if mps.to_right else ("Enviro", "System")
                if method == "1site":
                    lidx, cidx, ridx = imps - 1, [imps], imps + 1
                elif method == "2site":
                    if mps.to_right:
                        lidx, cidx, ridx = imps - 1, [imps, imps + 1], imps + 2
                    else:
                        lidx, cidx, ridx = imps - 2, [imps - 1, imps], imps + 1
                else:
                    assert False
                conj = bra()
                ltensor = environ.GetLR("L", lidx, ket, mpo, itensor=None, method=lmethod, mps_conj=conj)
                rtensor = environ.GetLR("R", ridx, ket, mpo, itensor=None, method=rmethod, mps_conj=conj)
                qnbigl, qnbigr, _ = mps._get_big_qn(cidx, need_mat=False)
                qn_mask = qn_mask_outer(qnbigl, qnbigr, mps.qntot)
                cmo = [mpo[idx] for idx in cidx]
                cms = ket[cidx[0]] if method == "1site" else ops.tensordot1(ket[cidx[0]], ket[cidx[1]])
                hop = hop_expr_dtype(ltensor, rtensor, cmo, list(cms.shape), dtype)
                cout = hop(cms)
                hop.close()
                # back to the guess's own bond dimensions; drop what violates the quantum numbers
                cout = cout[:mps.bond_dims[cidx[0]], ..., :mps.bond_dims[cidx[-1] + 1]].contiguous()
                cout = cout * torch.from_numpy(qn_mask).to(cout.device).to(cout.dtype)
                mps._update_mps(cout, cidx, qnbigl, qnbigr, percent)
            mps._switch_direction()
            if isweep > 0 and percent == 0:
                error = mps.distance(mps_old) / np.sqrt(mps.dot(mps.conj()).real)
                if error < mps.compress_config.vrtol:
                    break
            mps_old = mps.copy()
        mps.canonicalise()
        return mps

    def expand_bond_dimension(self, hint_mpo=None, coef=1e-10, include_ex=True, ex_mps=None):
        """Fill the bond dimensions up to compress_config's limit with states reached through
        `hint_mpo` (mps.py:636-640, 1934-2023), the preparation step of a TDVP-PS run.  The
        reference's `include_ex=True` admixes a model-specific excited state (ground state +
        creation operators, or the maximally entangled state): pass it as `ex_mps`; a random expander
        (`hint_mpo=None`) needs the model classes as well.  Both are outside the sweep path."""
        from .lib import compressed_sum
        if hint_mpo is None:
            raise NotImplementedError("a random expander needs the reference's Model (mps.py:1983-1984); "
                                      "pass a hint MPO")
        if include_ex and ex_mps is None:
            raise NotImplementedError("include_ex=True builds a model-specific state (mps.py:1944-1957): "
                                      "pass it as ex_mps, or use include_ex=False")
        mps = self
        mps.compress_config.set_bonddim(len(mps.bond_dims))
        m_target = np.minimum(np.array(mps.compress_config.max_dims) - np.array(mps.bond_dims),
                              mps.bond_dims_exact).astype(int)
        if ex_mps is not None:
            ex_mps.compress_config = mps.compress_config
            ex_mps.move_qnidx(mps.qnidx)
            ex_mps.to_right = mps.to_right
            lastone = mps + ex_mps
        else:
            lastone = mps
        expander_list = []
        expander_dims = np.zeros_like(m_target)
        while True:
            lastone = (hint_mpo @ lastone).normalize("mps_and_coeff")
            # more bond dimension for `lastone` for quick increase
            lastone = lastone.canonicalise().compress(int(np.max(m_target)))
            expander_list.append(lastone)
            expander = compressed_sum(expander_list, temp_m_trunc=m_target)
            if np.all(np.array(expander.bond_dims) >= m_target):
                break
            if np.all(np.array(expander.bond_dims) == expander_dims):
                # the expander does not grow any more: the target is too high
                m_target2 = np.max(m_target - np.array(expander_dims))
                expander2 = (hint_mpo @ lastone).canonicalise().compress(int(np.maximum(m_target2, 1)))
                expander = expander + expander2
                break
            expander_dims = np.array(expander.bond_dims)
            temp_m_trunc = int(np.max(m_target) / np.max(hint_mpo.bond_dims)) + 1
            lastone = lastone.canonicalise().compress(temp_m_trunc)
        return ((mps + expander.scale(coef * mps.norm, inplace=True)).canonicalise()
                .compress(mps.compress_config.max_dims).normalize("mps_norm_to_coeff"))

    def _evolve_prop_and_compress(self, mpo, evolve_dt):
        """Propagate and compress, mps.py:796-884: the Taylor expansion of exp(-i H dt) (4th order
        with a fixed step, 5th with adaptive step control); every H^k psi is Mpo.contract (apply +
        canonicalise + compress, with a threshold criterion tightened to `both` while contracting),
        the scaled terms are summed and compressed once more.  The truncations are the SVD path of
        this library."""
        from math import factorial
        from .lib import compressed_sum
        config = self.evolve_config
        order = config.taylor_order
        termlist = [self]
        orig = self.compress_config
        contract_config = self.compress_config.copy()
        if contract_config.criteria is CompressCriteria.threshold:
            contract_config.criteria = CompressCriteria.both
        self.compress_config = contract_config
        while len(termlist) < order + 1:
            termlist.append(mpo.contract(termlist[-1]))
        for t in termlist:
            t.compress_config = orig
        if not config.adaptive:
            scaled = [term.scale((-1.0j * evolve_dt) ** idx / factorial(idx), inplace=idx > 0)
                      for idx, term in enumerate(termlist)]
            return compressed_sum(scaled)
        # adaptive step control (mps.py:826-880): the error estimate is the distance between the sums
        # to order-1 and to order; the remaining time is evolved by recursion
        config.check_valid_dt(evolve_dt)
        p_restart, p_min, p_max = 0.5, 0.1, 2.0
        while True:
            dt = config.guess_dt if abs(config.guess_dt) < abs(evolve_dt) else evolve_dt
            scaled = [term.scale((-1.0j * dt) ** idx / factorial(idx)) for idx, term in enumerate(termlist)]
            new_mps1 = compressed_sum(scaled[:-1])
            new_mps2 = compressed_sum([new_mps1, scaled[-1]])
            dis = new_mps1.distance(new_mps2)
            p = (config.adaptive_rtol / (dis / new_mps2.mp_norm + 1e-30)) ** (1 / order)
            if np.allclose(dt, evolve_dt):
                if p < p_restart:                       # not accurate in this final sub-step: restart
                    config.guess_dt = dt * max(p_min, p)
                else:
                    new_mps2.evolve_config.guess_dt = dt * p if abs(dt * p) < abs(config.guess_dt) \
                        else config.guess_dt
                    return new_mps2
            else:
                if p < p_restart:
                    config.guess_dt *= max(p_min, p)
                else:
                    config.guess_dt *= min(p, p_max)
                    new_mps2.evolve_config.guess_dt = config.guess_dt
                    del new_mps1, termlist, scaled
                    return new_mps2._evolve_prop_and_compress(mpo, evolve_dt - dt)

    def evolve(self, mpo, evolve_dt, normalize=True):
        """mps.py:644-662: propagate-and-compress (the default) and the projector-splitting
        integrators TDVP-PS / TDVP-PS2; the variational mean-field TDVP variants are not sweeps over
        H_eff and stay outside the accelerated path."""
        method = self.evolve_config.method
        if method in (EvolveMethod.prop_and_compress, EvolveMethod.prop_and_compress_tdrk4,
                      EvolveMethod.prop_and_compress_tdrk):
            new_mps = {EvolveMethod.prop_and_compress: self._evolve_prop_and_compress,
                       EvolveMethod.prop_and_compress_tdrk4: self._evolve_prop_and_compress_tdrk4,
                       EvolveMethod.prop_and_compress_tdrk: self._evolve_prop_and_compress_tdrk}[method](mpo, evolve_dt)
            if normalize:
                new_mps.normalize("mps_and_coeff" if np.iscomplex(evolve_dt) else "mps_only")
            return new_mps
        if method not in (EvolveMethod.tdvp_ps, EvolveMethod.tdvp_ps2):
            raise NotImplementedError(
                f"evolve method {method} is outside the accelerated path "
                "(propagate-and-compress, TDVP-PS and TDVP-PS2, mps.py:796,1268,1407, are)")
        if self.evolve_config.ivp_solver != "krylov":
            raise NotImplementedError("only the Krylov local solver is accelerated")
        if method is EvolveMethod.tdvp_ps2:
            if self.evolve_config.adaptive:
                raise NotImplementedError("adaptive step control wraps the one-site integrator only (mps.py:46)")
            new_mps = self._evolve_tdvp_ps2(mpo, evolve_dt)
        elif self.evolve_config.adaptive:
            new_mps = self._evolve_adaptive(mpo, evolve_dt)
        else:
            new_mps = self._evolve_tdvp_ps(mpo, evolve_dt)
        if normalize:
            if np.iscomplex(evolve_dt):
                new_mps.normalize("mps_and_coeff")
            else:
                new_mps.normalize("mps_only")
        return new_mps

    @staticmethod
    def _mpo_of_time(mpo):
        """mps.py:669-676: a fixed MPO, or a callable t -> MPO over 0 .. evolve_dt."""
        if callable(mpo) and not hasattr(mpo, "contract"):
            return mpo
        if not hasattr(mpo, "contract"):
            raise TypeError(f"unsupported mpo type: {mpo}")
        return lambda t, *args, **kwargs: mpo

    def _evolve_prop_and_compress_tdrk4(self, mpo, evolve_dt):
        """Classical 4th-order Runge-Kutta step for a (possibly time-dependent) Hamiltonian,
        mps.py:664-698: every stage is Mpo.contract (apply + canonicalise + compress), the stage
        states and the final sum are canonicalised and compressed with the state's own configuration."""
        from .lib import compressed_sum
        mpo_t = self._mpo_of_time(mpo)
        k1 = mpo_t(0).contract(self).scale(-1j)
        tmp = self + k1.scale(0.5 * evolve_dt)
        tmp.canonicalise().compress()
        k2 = mpo_t(0.5 * evolve_dt).contract(tmp).scale(-1j)
        tmp = self + k2.scale(0.5 * evolve_dt)
        tmp.canonicalise().compress()
        k3 = mpo_t(0.5 * evolve_dt).contract(tmp).scale(-1j)
        tmp = self + k3.scale(evolve_dt)
        tmp.canonicalise().compress()
        k4 = mpo_t(evolve_dt).contract(tmp).scale(-1j)
        return compressed_sum([self, k1.scale(1 / 6 * evolve_dt), k2.scale(2 / 6 * evolve_dt),
                               k3.scale(2 / 6 * evolve_dt), k4.scale(1 / 6 * evolve_dt)])

    def _evolve_prop_and_compress_tdrk(self, mpo, evolve_dt):
        """General explicit Runge-Kutta step from the tableau of evolve_config.rk_config, with the
        embedded-pair step-size control when evolve_config.adaptive, mps.py:700-793."""
        from functools import reduce
        from .lib import compressed_sum
        mpo_t = self._mpo_of_time(mpo)
        rk = self.evolve_config.rk_config
        a, b, c = rk.tableau

        def sub_step(y, tau, t0):
            ks = []
            for istage in range(rk.stage):
                k = compressed_sum([y] + [ks[i].scale(a[istage, i] * tau) for i in range(istage) if a[istage, i] != 0],
                                   batchsize=6)
                k = mpo_t(c[istage] * tau + t0, mps=k).contract(k).scale(-1j)
                ks.append(k)
            new = compressed_sum([y] + [ks[i].scale(b[0, i] * tau) for i in range(rk.stage) if b[0, i] != 0],
                                 batchsize=6)
            if not self.evolve_config.adaptive:
                assert len(rk.order) == 1
                return new, 0
            assert len(rk.order) == 2 and rk.order[0] - rk.order[1] == 1
            err = reduce(lambda m1, m2: m1.add(m2),
                         [ks[i].scale((b[0, i] - b[1, i]) * tau) for i in range(rk.stage)
                          if not np.allclose(b[0, i], b[1, i])])
            return new, err.norm / new.norm

        self.evolve_config.check_valid_dt(evolve_dt)
        if not self.evolve_config.adaptive:
            return sub_step(self, evolve_dt, 0)[0]
        p_restart, p_min, p_max = 0.5, 0.1, 2.0
        evolved = 0
        new = self

        def min_abs(x, y):
            return x if abs(x) < abs(y) else y
        while True:
            dt = min_abs(new.evolve_config.guess_dt, evolve_dt - evolved)
            # as in the reference (mps.py:757-759) the trial state replaces the current one even when the
            # step is then judged inaccurate and repeated with a smaller guess
            new, error = sub_step(new, dt, evolved)
            p = (new.evolve_config.adaptive_rtol / (error + 1e-30)) ** (1 / rk.order[0])
            if p < p_restart:
                new.evolve_config.guess_dt = dt * max(p_min, p)
            elif np.allclose(dt + evolved, evolve_dt):
                new.evolve_config.guess_dt = min_abs(dt * p, new.evolve_config.guess_dt)
                return new
            else:
                new.evolve_config.guess_dt *= min(p, p_max)
                evolved += dt

    def _evolve_adaptive(self, mpo, evolve_target_t):
        """mps.py:46-115 (adaptive_tdvp): step-doubling error control around _evolve_tdvp_ps."""
        config = self.evolve_config.copy()
        cur = self
        p_restart, p_min, p_max = 0.5, 0.1, 2.0
        evolved_t = 0
        while True:
            rest = evolve_target_t - evolved_t
            dt = config.guess_dt if abs(config.guess_dt) < abs(rest) else rest
            half1 = cur._evolve_tdvp_ps(mpo, dt / 2)
            half2 = half1._evolve_tdvp_ps(mpo, dt / 2)
            full = cur._evolve_tdvp_ps(mpo, dt)
            dis = full.distance(half2)
            del half1, full
            p = (0.75 * config.adaptive_rtol / (dis / half2.mp_norm + 1e-30)) ** (1.0 / 3)
            p = min(max(p, p_min), p_max)
            if p < p_restart:
                config.guess_dt = dt * p
                continue
            evolved_t += dt
            if np.allclose(evolved_t, evolve_target_t):
                half2.evolve_config.guess_dt = config.guess_dt
                return half2
            config.guess_dt *= p
            cur = half2

    def _evolve_tdvp_ps(self, mpo, evolve_dt, site_filter=None, site_probe=None, half_sweeps=2, site_hook=None):
        """One-site projector-splitting TDVP step (mps.py:1268-1404, Krylov local solver):
        forward half sweep and backward half sweep, each site evolved by dt/2 with H_eff and each
        bond matrix evolved backwards with the zero-site H_eff.

        Measurement aids (not in the reference, used by bench.py only): `site_filter`, a set of site
        indices -- the other sites are passed over with the QR and the environment update alone;
        `site_probe(imps, energy)` receives Re <C|H_eff|C> of every evolved centre tensor, a gauge
        invariant the CPU oracle is compared with; `half_sweeps` stops after the first half;
        `site_hook(stage, imps, info)` is called before ("pre": l_array, r_array, mps) and after
        ("done") every evolved site."""
        # mps.py:1272-1279: imaginary time keeps the state's dtype.  A step with both a real and an
        # imaginary part makes exp(-i dt H_eff) complex, so a real state is promoted (the reference's
        # NumPy arithmetic promotes implicitly).
        if np.iscomplex(evolve_dt) and complex(evolve_dt).real == 0:
            mps = self.copy()
        else:
            mps = self.to_complex()
        cdtype = mps.dtype
        # The reference builds both environment chains here and notes that "almost half is not
        # used" (mps.py:1282-1284): a sweep that starts at the right end only ever reads the L
        # chain of the initial state (the R environments are rebuilt site by site before they are
        # read), and vice versa.  Only the chain that is read is constructed.
        environ = Environ(mps, mpo, "R" if mps.to_right else "L")
        local_steps = []
        n = len(mps)
        for _ in range(half_sweeps):
            for imps in mps.iter_idx_list(full=True):
                system = "L" if mps.to_right else "R"
                l_array = environ.read("L", imps - 1)
                r_array = environ.read("R", imps + 1)
                shape = list(mps[imps].shape)
                evolve_site = site_filter is None or imps in site_filter
                if evolve_site and site_hook is not None:
                    site_hook("pre", imps, dict(l_array=l_array, r_array=r_array, mps=mps))
                if evolve_site:
                    hop = hop_expr_dtype(l_array, r_array, [mpo[imps]], shape, cdtype)
                    mps_t, j = expm_krylov(hop, -1j * evolve_dt / 2, mps[imps].reshape(-1))
                    if site_probe is not None:
                        hc = hop(mps_t.reshape(shape)).reshape(-1)
                        site_probe(imps, float(torch.vdot(mps_t.reshape(-1), hc).real))
                    hop.close()
                    local_steps.append(j)
                else:
                    mps_t = mps[imps]
                mps_t = mps_t.reshape(shape)
                qnbigl, qnbigr, _ = mps._get_big_qn([imps], need_mat=False)
                u, qnlset, v, qnrset = svd_qn(mps_t, qnbigl, qnbigr, mps.qntot, QR=True,
                                              system=system, full_matrices=False)
                vt = v.transpose(0, 1).contiguous()
                if not mps.to_right and imps != 0:
                    mps[imps] = vt.reshape([-1] + shape[1:])
                    mps.qn[imps] = np.array(qnrset)
                    mps.qnidx = imps - 1
                    r_array = environ.GetLR("R", imps, mps, mpo, itensor=r_array, method="System")
                    u = u.contiguous()
                    shape_u = list(u.shape)
                    if evolve_site:
                        hop_u = hop_expr_dtype(l_array, r_array, [], shape_u, cdtype)
                        back, j = expm_krylov(hop_u, 1j * evolve_dt / 2, u.reshape(-1))
                        hop_u.close()
                        local_steps.append(j)
                    else:
                        back = u
                    mps[imps - 1] = ops.tensordot1(mps[imps - 1], back.reshape(shape_u))
                elif mps.to_right and imps != n - 1:
                    mps[imps] = u.contiguous().reshape(shape[:-1] + [-1])
                    mps.qn[imps + 1] = np.array(qnlset)
                    mps.qnidx = imps + 1
                    l_array = environ.GetLR("L", imps, mps, mpo, itensor=l_array, method="System")
                    shape_svt = list(vt.shape)
                    if evolve_site:
                        hop_svt = hop_expr_dtype(l_array, r_array, [], shape_svt, cdtype)
                        back, j = expm_krylov(hop_svt, 1j * evolve_dt / 2, vt.reshape(-1))
                        hop_svt.close()
                        local_steps.append(j)
                    else:
                        back = vt
                    mps[imps + 1] = ops.tensordot1(back.reshape(shape_svt), mps[imps + 1])
                else:
                    mps[imps] = mps_t
                if evolve_site and site_hook is not None:
                    site_hook("done", imps, {})
            mps._switch_direction()
        mps.evolve_config.stat = local_steps
        return mps

    def _evolve_tdvp_ps2(self, mpo, evolve_dt):
        """Two-site projector-splitting TDVP step (mps.py:1407-1517, Krylov local solver): every
        pair of neighbouring sites is evolved forward by dt/2 with the two-site H_eff and split by
        the truncating SVD of `_update_mps`; the site that moves on with the sweep is evolved
        backwards with the one-site H_eff and re-canonicalised."""
        # mps.py:1272-1279: imaginary time keeps the state's dtype.  A step with both a real and an
        # imaginary part makes exp(-i dt H_eff) complex, so a real state is promoted (the reference's
        # NumPy arithmetic promotes implicitly).
        if np.iscomplex(evolve_dt) and complex(evolve_dt).real == 0:
            mps = self.copy()
        else:
            mps = self.to_complex()
        cdtype = mps.dtype
        # mps.py:1422-1424 builds both environment chains; only the one this sweep reads is needed
        environ = Environ(mps, mpo, "R" if mps.to_right else "L")
        local_steps = []
        n = len(mps)
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
                ms2 = ops.tensordot1(mps[cidx0], mps[cidx1])
                shape2 = list(ms2.shape)
                hop = hop_expr_dtype(l_array, r_array, [mpo[cidx0], mpo[cidx1]], shape2, cdtype)
                mps_t, j = expm_krylov(hop, -1j * evolve_dt / 2, ms2.reshape(-1))
                hop.close()
                local_steps.append(j)
                qnbigl, qnbigr, _ = mps._get_big_qn([cidx0, cidx1], need_mat=False)
                mps._update_mps(mps_t.reshape(shape2), [cidx0, cidx1], qnbigl, qnbigr)
                if imps == last_idx:
                    continue
                if mps.to_right:
                    l_array = environ.GetLR("L", lidx + 1, mps, mpo, itensor=l_array, method="System")
                else:
                    r_array = environ.GetLR("R", ridx - 1, mps, mpo, itensor=r_array, method="System")
                ms1 = mps[cidx2]
                shape1 = list(ms1.shape)
                hop1 = hop_expr_dtype(l_array, r_array, [mpo[cidx2]], shape1, cdtype)
                back, j = expm_krylov(hop1, 1j * evolve_dt / 2, ms1.reshape(-1))
                hop1.close()
                local_steps.append(j)
                mps[cidx2] = back.reshape(shape1)
                mps._push_cano(cidx2)
            mps._switch_direction()
        mps.evolve_config.stat = local_steps
        return mps

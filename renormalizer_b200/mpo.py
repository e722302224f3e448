"""MPO container for the sweep path: a list of site tensors W[b, up, down, f] kept both on the
host (sparse structure) and on the device.  The symbolic MPO construction of the reference
(renormalizer/mps/mpo.py, symbolic_mpo.py) is outside the accelerated path; an Mpo here is built
from site tensors (e.g. `Mpo([np.asarray(m.array) for m in reference_mpo])`)."""
import numpy as np

from .ops import MpoSite


class Mpo:
    def __init__(self, site_tensors, offset=0.0, qn=None, qntot=None, qnidx=None, sigmaqn=None,
                 to_right=False):
        """`qn` (one array per bond), `qntot` and `qnidx` are the operator's quantum numbers
        (mp.py:34-80); only `apply` / `contract` read them, and an operator that conserves every
        quantum number (all zero, the default) needs none.  `sigmaqn` (per site, shape (d, d, nq):
        Mpo._get_sigmaqn, mpo.py:293-295) and `to_right` are needed only to compress the operator
        itself (the default guess of the variational compression)."""
        self._sites = [w if isinstance(w, MpoSite) else MpoSite(w) for w in site_tensors]
        self.offset = offset
        self.qn = None if qn is None else [np.asarray(q) for q in qn]
        self.qntot = None if qntot is None else np.asarray(qntot)
        self.qnidx = len(self._sites) - 1 if qnidx is None else int(qnidx)
        self.sigmaqn = None if sigmaqn is None else [np.asarray(q) for q in sigmaqn]
        self.to_right = bool(to_right)

    def as_matrix_product(self, nq=1):
        """The operator as a matrix product with two physical indices per site (an `Mps` object
        with is_mpo set): what MatrixProduct.canonicalise / compress act on for an Mpo."""
        from .mps import Mps
        qn = self.qn if self.qn is not None else [np.zeros((d, nq), dtype=int) for d in self.bond_dims]
        qntot = self.qntot if self.qntot is not None else np.zeros(nq, dtype=int)
        sq = self.sigmaqn if self.sigmaqn is not None else \
            [np.zeros((d, d, nq), dtype=int) for d in self.pbond_list]
        mp = Mps([s.dense for s in self._sites], qn, sq, qntot, self.qnidx, self.to_right)
        mp.is_mpo = True
        return mp

    def __len__(self):
        return len(self._sites)

    def __getitem__(self, i):
        return self._sites[i]

    def __iter__(self):
        return iter(self._sites)

    @property
    def site_num(self):
        return len(self._sites)

    @property
    def bond_dims(self):
        return [s.shape[0] for s in self._sites] + [self._sites[-1].shape[-1]]

    @property
    def pbond_list(self):
        return [s.shape[1] for s in self._sites]

    def to_numpy(self):
        return [s.array.copy() for s in self._sites]

    # ---- the little MPO algebra the omega-targeting sweep needs (gs.py:106-111) -----------------
    @classmethod
    def identity_like(cls, mpo):
        """Bond-dimension-1 identity with the physical dimensions of `mpo` (Mpo.identity)."""
        return cls([np.eye(d).reshape(1, d, d, 1) for d in mpo.pbond_list])

    def scale(self, val):
        """Multiply the operator by a real scalar (carried by the first site)."""
        sites = self.to_numpy()
        sites[0] = sites[0] * float(val)
        return Mpo(sites, self.offset)

    def add(self, other):
        """Operator sum as the direct sum of the bond spaces (MatrixProduct.add, mp.py:374-409)."""
        assert len(self) == len(other) and self.pbond_list == other.pbond_list
        n = len(self)
        out = []
        for i, (a, b) in enumerate(zip(self.to_numpy(), other.to_numpy())):
            if n == 1:
                out.append(a + b)
            elif i == 0:
                out.append(np.concatenate([a, b], axis=3))
            elif i == n - 1:
                out.append(np.concatenate([a, b], axis=0))
            else:
                w = np.zeros((a.shape[0] + b.shape[0], a.shape[1], a.shape[2], a.shape[3] + b.shape[3]))
                w[:a.shape[0], :, :, :a.shape[3]] = a
                w[a.shape[0]:, :, :, a.shape[3]:] = b
                out.append(w)
        return Mpo(out, self.offset)

    def squared(self):
        """The operator product O.O as ONE MPO with merged bonds,
        W2[(b,c), up, down, (g,i)] = sum_f W[b, up, f, g] W[c, f, down, i]:
        the reference's two-layer environments `Environ(mps, [mpo, mpo])` with the pair of MPO
        bonds (b, c) read as one index."""
        return Mpo([two_layer_site(w, w) for w in self.to_numpy()], self.offset)

    @property
    def nbytes(self):
        return sum(s.array.nbytes for s in self._sites)

    # ---- operator application (propagate-and-compress, mps.py:796-884) ---------------------------
    def apply(self, mps, canonicalise: bool = False):
        """mpo @ mps (or @ mpdm) without compression, mpo.py:331-389:
        new[(a,c), p, (b,d)] = sum_q W[a,p,q,b] A[c,q,d] as one device GEMM per site."""
        from . import ops
        from .svd_qn import add_outer
        assert self.site_num == mps.site_num
        new = mps.metacopy()
        for i, (site, a) in enumerate(zip(self._sites, mps)):
            w = site.dense                                         # (a, p, q, b) on the device
            wa, wp, wq, wb = w.shape
            assert wq == a.shape[1]
            w2 = w.permute(0, 1, 3, 2).reshape(wa * wp * wb, wq)
            rest = tuple(a.shape[2:-1])                            # () for an MPS, (r,) for an MPDM
            nrest = int(np.prod(rest)) if rest else 1
            a2 = a.movedim(1, 0).reshape(wq, a.shape[0] * nrest * a.shape[-1])
            m = ops.matmul(w2.to(a.dtype) if a.is_complex() else w2, a2)
            m = m.reshape(wa, wp, wb, a.shape[0], nrest, a.shape[-1]).permute(0, 3, 1, 4, 2, 5)
            new[i] = m.reshape((wa * a.shape[0], wp) + rest + (wb * a.shape[-1],)).contiguous()
        nq = len(new.qntot)
        qn = self.qn if self.qn is not None else [np.zeros((d, nq), dtype=int) for d in self.bond_dims]
        qntot = self.qntot if self.qntot is not None else np.zeros(nq, dtype=int)
        orig_idx = new.qnidx
        new.move_qnidx(self.qnidx)
        new.qn = [add_outer(np.array(qo), np.array(qm)).reshape(-1, nq) for qo, qm in zip(qn, new.qn)]
        new.qntot = new.qntot + qntot
        new.move_qnidx(orig_idx)
        if canonicalise:
            new.canonicalise()
        return new

    def __matmul__(self, other):
        return self.apply(other)

    @property
    def bond_dims_mean(self):
        return int(round(np.mean(self.bond_dims)))

    def contract(self, mps, algo="svd"):
        """An approximation of mpo @ mps: apply -> canonicalise -> compress (mpo.py:391-425)."""
        if algo == "variational":
            return mps.variational_compress(self)
        if algo != "svd":
            raise AssertionError(f"unknown compression algorithm {algo}")
        new = self.apply(mps)
        new.canonicalise()
        new.compress()
        return new


class StackedMpo:
    """Sum of Hamiltonians kept as separate MPOs (block-diagonal sparse form, mpo.py:483-494):
    `optimize_mps(mps, StackedMpo([mpo1, mpo2, ...]))` sums the effective Hamiltonians of the
    members at every site."""

    def __init__(self, mpos):
        self.mpos = list(mpos)


def two_layer_site(upper, lower):
    """W2[(b,c), up, down, (g,i)] = sum_f upper[b, up, f, g] lower[c, f, down, i]."""
    b, d, _, g = upper.shape
    c, _, _, i = lower.shape
    w = np.einsum("befg,cfhi->bcehgi", upper, lower)
    return np.ascontiguousarray(w.reshape(b * c, d, d, g * i))

"""MPO container for the sweep path: a list of site tensors W[b, up, down, f] kept both on the
host (sparse structure) and on the device.  The symbolic MPO construction of the reference
(renormalizer/mps/mpo.py, symbolic_mpo.py) is outside the accelerated path; an Mpo here is built
from site tensors (e.g. `Mpo([np.asarray(m.array) for m in reference_mpo])`)."""
import numpy as np

from .ops import MpoSite


class Mpo:
    def __init__(self, site_tensors, offset=0.0):
        self._sites = [w if isinstance(w, MpoSite) else MpoSite(w) for w in site_tensors]
        self.offset = offset

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

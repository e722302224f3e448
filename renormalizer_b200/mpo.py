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

    @property
    def nbytes(self):
        return sum(s.array.nbytes for s in self._sites)

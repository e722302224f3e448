"""Environments -- mirror of renormalizer/mps/lib.py:12-262 (Environ, contract_one_site)."""
import torch

from . import ops
from .backend import asxp, backend


def contract_one_site(environ, ms, mo, domain, ms_conj=None):
    """Absorb one MPS/MPDM site and its MPO site into a left ("L") or right ("R") environment.

    Reference: lib.py:172-262.  `ms_conj` (optional) is the already-conjugated bra-side tensor,
    as in the reference; by default the bra is conj(ms).
    """
    if domain not in ("L", "R"):
        raise AssertionError("domain must be 'L' or 'R'")
    environ, ms = asxp(environ), asxp(ms)
    if ms.ndim not in (3, 4):
        raise ValueError(f"MPS ndim is not 3 or 4, got {ms.ndim}")
    bra = ms if ms_conj is None else asxp(ms_conj).conj().resolve_conj()
    return ops.env_update(environ, bra, ms, ops.as_mpo_site(mo), domain)


class Environ:
    """Left/right environment store; L(idx-1) - mpo(idx) - R(idx+1).  Reference: lib.py:12-129.
    Environments stay resident in HBM (the reference's `_virtual_disk` moves them to the host)."""

    def __init__(self, mps, mpo, domain=None, mps_conj=None):
        self._virtual_disk = {}
        self.sentinel = torch.ones((1, 1, 1), dtype=backend.real_dtype, device=backend.device)
        self._construct(mps, mpo, domain, mps_conj)

    def _construct(self, mps, mpo, domain=None, mps_conj=None):
        assert domain in ["L", "R", None]
        if domain is None:
            self._construct(mps, mpo, "L", mps_conj)
            self._construct(mps, mpo, "R", mps_conj)
            return
        n = len(mps)
        if domain == "L":
            start, end, inc = 0, n - 1, 1
        else:
            start, end, inc = n - 1, 0, -1
        self.write("L", -1, self.sentinel)
        self.write("R", n, self.sentinel)
        tensor = self.sentinel
        for idx in range(start, end, inc):
            conj = None if mps_conj is None else mps_conj[idx]
            tensor = contract_one_site(tensor, mps[idx], mpo[idx], domain, ms_conj=conj)
            self.write(domain, idx, tensor)

    def GetLR(self, domain, siteidx, mps, mpo, itensor=None, method="Scratch", mps_conj=None):
        assert domain in ["L", "R"]
        assert method in ["Enviro", "System", "Scratch"]
        if mps_conj is None:
            mps_conj = [None] * len(mps)
        if siteidx not in range(len(mps)):
            return self.sentinel
        if method == "Scratch":
            itensor = self.sentinel
            sitelist = range(siteidx + 1) if domain == "L" else range(len(mps) - 1, siteidx - 1, -1)
            for imps in sitelist:
                itensor = contract_one_site(itensor, mps[imps], mpo[imps], domain, ms_conj=mps_conj[imps])
        elif method == "Enviro":
            itensor = self.read(domain, siteidx)
        else:
            if itensor is None:
                offset = -1 if domain == "L" else 1
                itensor = self.read(domain, siteidx + offset)
            itensor = contract_one_site(itensor, mps[siteidx], mpo[siteidx], domain, mps_conj[siteidx])
            self.write(domain, siteidx, itensor)
        return itensor

    def write(self, domain, siteidx, tensor):
        self._virtual_disk[(domain, siteidx)] = tensor

    def read(self, domain, siteidx):
        return self._virtual_disk[(domain, siteidx)]


def compressed_sum(mps_list, batchsize=5, temp_m_trunc=None):
    """lib.py:417-439: sum in batches, canonicalise and compress after every batch."""
    assert len(mps_list) != 0
    queue = list(mps_list)
    if len(queue) == 1:
        new = queue[0].canonicalise()
        new.compress(temp_m_trunc=temp_m_trunc)
        return new
    while len(queue) != 1:
        batch, queue = queue[:batchsize], queue[batchsize:]
        s = batch[0]
        for t in batch[1:]:
            s = s.add(t)
        s.canonicalise()
        s.compress(temp_m_trunc=temp_m_trunc)
        queue.append(s)
    return queue[0]

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
    site = ops.as_mpo_site(mo)
    from . import parallel
    group = parallel.heff_group()
    if group is not None:
        out = _contract_one_site_sharded(environ, bra, ms, site, domain, group)
        if out is not None:
            return out
    return ops.env_update(environ, bra, ms, site, domain)


def _contract_one_site_sharded(environ, bra, ket, site, domain, group):
    """One sweep on several GPUs (parallel.enable_sharded_heff): the environment update costs as much as
    one H_eff application, so it is split the same way -- over the ket's NEW bond h of
    "abc, adf, bdeg, ceh -> fgh" (lib.py:214-258), which every step of the contraction chain carries:
    rank r absorbs ket[..., h_r] and the slices out[:, :, h_r] are all-gathered.  Returns None when the
    update is too small to be worth a collective."""
    import torch.distributed as dist
    from . import parallel
    grp = None if group is True else group
    rank, world = dist.get_rank(grp), dist.get_world_size(grp)
    mh = ket.shape[-1] if domain == "L" else ket.shape[0]
    ea, eb, ec = environ.shape
    inner = 1
    for x in ket.shape[1:-1]:
        inner *= int(x)
    mf = bra.shape[-1] if domain == "L" else bra.shape[0]
    flops = 2.0 * ea * eb * ec * inner * mh + 2.0 * ea * inner * mf * site.shape[0 if domain == "R" else 3] * mh
    if flops < parallel._heff["min_work"] or mh % world != 0:
        return None
    w = mh // world
    if domain == "L":
        part = ket[..., rank * w:(rank + 1) * w].contiguous()
    else:
        part = ket[rank * w:(rank + 1) * w].contiguous()
    local = ops.env_update(environ, bra, part, site, domain)            # (Mf, F, w)
    # gathered along the first axis: (world * Mf, F, w), read below as (world, Mf, F, w)
    buf = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    local = local.contiguous()
    if local.is_complex():                      # collectives see the interleaved real view
        dist.all_gather_into_tensor(torch.view_as_real(buf), torch.view_as_real(local), group=grp)
    else:
        dist.all_gather_into_tensor(buf, local, group=grp)
    parallel._heff["gathered_bytes"] += buf.numel() * buf.element_size()
    buf = buf.reshape((world,) + tuple(local.shape))
    return buf.permute(1, 2, 0, 3).reshape(local.shape[0], local.shape[1], world * w).contiguous()


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

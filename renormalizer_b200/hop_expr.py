"""H_eff . C -- mirror of renormalizer/mps/hop_expr.py:7-117 on the B200 path."""
import torch

from . import ops
from .backend import asxp


class _HopCallable:
    def __init__(self, plan):
        self.plan = plan

    def __call__(self, cstruct):
        c = asxp(cstruct)
        return self.plan.apply(c.reshape(self.plan.in_shape))

    def close(self):
        self.plan.close()


def hop_expr(ltensor, rtensor, cmo, cshape, twolayer: bool = False):
    """Return `expr` with expr(cstruct) = H_eff . cstruct for the centre tensor of shape cshape.

    Same arguments as the reference (hop_expr.py:7): ltensor (a,b,c), rtensor (l,f,k), cmo a list
    of 0, 1 or 2 MPO site tensors, cshape with or without ancilla indices.

    twolayer=True (hop_expr.py:24-52, the (H - omega)^2 expressions): ltensor (a,b,c,d) and
    rtensor (j,g,i,k) carry two MPO bonds and the result is
        out[d,h,k]   = L[a,b,c,d] W[b,e,f,g] W[c,f,h,i] R[j,g,i,k] C[a,e,j]        (one site)
        out[d,h,m,p] = L W1 W1 W2 W2 R C                                         (two sites).
    It runs on the same kernels as the one-layer case: the pair of MPO bonds is merged into one
    index, the MPO site becomes the product site of `mpo.two_layer_site` and the environments
    are read with their outer bonds exchanged (input on the a side, output on the d side).
    """
    ltensor, rtensor = asxp(ltensor), asxp(rtensor)
    if twolayer:
        import numpy as np
        from .mpo import two_layer_site
        nsite = len(cmo)
        if nsite not in (1, 2) or len(cshape) != nsite + 2:
            raise AssertionError("two-layer expressions exist for 1 or 2 sites without ancilla")
        a, b, c, d = ltensor.shape
        j, g, i, k = rtensor.shape
        lt = ltensor.permute(3, 1, 2, 0).reshape(d, b * c, a).contiguous()
        rt = rtensor.permute(3, 1, 2, 0).reshape(k, g * i, j).contiguous()
        sites = []
        for m in cmo:
            w = np.asarray(ops.as_mpo_site(m).array)
            # out index h is the DOWN index of the lower layer, the contracted e the UP index of
            # the upper layer: as a one-layer site W'[(b,c), h, e, (g,i)]
            sites.append(ops.MpoSite(np.ascontiguousarray(two_layer_site(w, w).transpose(0, 2, 1, 3))))
        cplx = lt.is_complex() or rt.is_complex()
        dtype = torch.complex128 if cplx else torch.float64
        return _HopCallable(ops.HopPlan(lt, rt, sites, cshape, dtype))
    sites = [ops.as_mpo_site(m) for m in cmo]
    cplx = ltensor.is_complex() or rtensor.is_complex()
    dtype = torch.complex128 if cplx else torch.float64
    return _HopCallable(ops.HopPlan(ltensor, rtensor, sites, cshape, dtype))


def hop_expr_dtype(ltensor, rtensor, cmo, cshape, dtype):
    """hop_expr with an explicit compute dtype (complex centre tensor on real environments).
    With parallel.enable_sharded_heff the application is split over the ranks of a process group
    (rows of L) and its result all-gathered; see parallel.ShardedHop."""
    from . import parallel
    ltensor, rtensor = asxp(ltensor), asxp(rtensor)
    sites = [ops.as_mpo_site(m) for m in cmo]
    group = parallel.heff_group()
    if (group is not None and ltensor.ndim == 3
            and parallel.heff_flops(ltensor.shape, rtensor.shape, cshape) >= parallel._heff["min_work"]):
        cshape = tuple(int(x) for x in cshape)
        return parallel.ShardedHop(
            ltensor, lambda l_slice: _HopCallable(ops.HopPlan(l_slice, rtensor, sites, cshape, dtype)),
            cshape[1:-1] + (int(rtensor.shape[0]),), group=group)
    return _HopCallable(ops.HopPlan(ltensor, rtensor, sites, cshape, dtype))

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
    of 0, 1 or 2 MPO site tensors, cshape with or without ancilla indices.  The two-layer
    (H - omega)^2 expressions (hop_expr.py:24-52) are outside the accelerated path.
    """
    if twolayer:
        raise NotImplementedError("two-layer (omega-targeting) H_eff is outside the accelerated path")
    ltensor, rtensor = asxp(ltensor), asxp(rtensor)
    sites = [ops.as_mpo_site(m) for m in cmo]
    cplx = ltensor.is_complex() or rtensor.is_complex()
    dtype = torch.complex128 if cplx else torch.float64
    return _HopCallable(ops.HopPlan(ltensor, rtensor, sites, cshape, dtype))


def hop_expr_dtype(ltensor, rtensor, cmo, cshape, dtype):
    """hop_expr with an explicit compute dtype (complex centre tensor on real environments)."""
    ltensor, rtensor = asxp(ltensor), asxp(rtensor)
    sites = [ops.as_mpo_site(m) for m in cmo]
    return _HopCallable(ops.HopPlan(ltensor, rtensor, sites, cshape, dtype))

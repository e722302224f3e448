"""Minimal stand-in for the third-party `opt_einsum` package (absent from this image).

TEST TOOLING ONLY: lets tests/golden/make_golden.py import the unmodified reference from
/root/reference in the build container so that golden vectors can be generated.  It maps
`contract` / `contract_expression` onto numpy.einsum (optimize="optimal"), which evaluates the
same pairwise tensordot chain opt_einsum would pick.  Never imported by the product.
"""
import numpy as np
from . import parser  # noqa: F401


_PATHS = {}


def _path(subscripts, ops):
    """Optimal pairwise contraction order, searched once per (expression, shapes)."""
    key = (subscripts, tuple(np.shape(o) for o in ops))
    if key not in _PATHS:
        _PATHS[key] = np.einsum_path(subscripts, *ops, optimize=("optimal" if len(ops) <= 7 else "greedy", 2 ** 40))[0]
    return _PATHS[key]


def contract(*args, **kwargs):
    kwargs.pop("backend", None)
    opt = kwargs.pop("optimize", "optimal")
    if opt not in ("optimal", "greedy", True, False):
        opt = "optimal"
    if isinstance(args[0], str) and opt == "optimal":
        return np.einsum(*args, optimize=_path(args[0], args[1:]))
    return np.einsum(*args, optimize=opt)


def contract_expression(subscripts, *operands, constants=None, optimize="optimal", **kwargs):
    constants = list(constants or [])
    n = len(operands)
    const_ops = {i: operands[i] for i in constants}
    var_pos = [i for i in range(n) if i not in const_ops]

    cache = {}

    def expr(*arrays, backend=None, **kw):
        assert len(arrays) == len(var_pos)
        ops = [None] * n
        for i, a in const_ops.items():
            ops[i] = a
        for i, a in zip(var_pos, arrays):
            ops[i] = a
        # the contraction path is searched once per expression (the two-layer expressions have
        # up to seven operands: an "optimal" search on every call would dominate the run time)
        if "path" not in cache:
            cache["path"] = _path(subscripts, ops)
        return np.einsum(subscripts, *ops, optimize=cache["path"])

    return expr

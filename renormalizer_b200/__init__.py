"""renormalizer_b200: B200-native (sm_100a) DMRG / TDVP sweep-site engine.

Drop-in for the hot path behind renormalizer.mps.backend: H_eff*C (hop_expr), the environment
update (contract_one_site) and the SVD/QR bond truncation (svd_qn), with the Davidson / Krylov /
sweep drivers that call them.  Hand-written CUDA behind a C ABI (include/rn_b200.h); PyTorch is
used for device memory and streams only.  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .backend import backend  # noqa: F401

__all__ = ["backend"]

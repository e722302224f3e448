"""Mirror of renormalizer/mps/backend.py:97-216 for the B200 path: dtype policy and device.

There is exactly one backend here (CUDA, sm_100a); `xp` of the reference corresponds to torch
device tensors.  Only 64-bit precision is offered (the reference default, backend.py:111-114).
"""
import numpy as np
import torch


class Backend:
    def __init__(self):
        self.real_dtype = torch.float64
        self.complex_dtype = torch.complex128
        self._canonical_atol = 1e-8      # backend.py:177-187
        self._canonical_rtol = 1e-5      # backend.py:190-200
        # GEMM path of the contraction kernels: 1 = tcgen05 int8 split GEMM (FP64-accurate, default;
        # contractions too small to fill a tile stay on the DMMA kernel), 0 = FP64 DMMA everywhere
        self.gemm_path = 1

    @property
    def dtypes(self):
        return self.real_dtype, self.complex_dtype

    @property
    def is_32bits(self):
        return False

    @property
    def canonical_atol(self):
        return self._canonical_atol

    @canonical_atol.setter
    def canonical_atol(self, value):
        self._canonical_atol = self._tol_checker(value)

    @property
    def canonical_rtol(self):
        return self._canonical_rtol

    @canonical_rtol.setter
    def canonical_rtol(self, value):
        self._canonical_rtol = self._tol_checker(value)

    @staticmethod
    def _tol_checker(value):
        if not isinstance(value, (int, float)) or value < 0:
            raise ValueError("Tolerance must be a non-negative float number")
        return value

    @property
    def device(self):
        return torch.device("cuda", torch.cuda.current_device())

    def sync(self):
        torch.cuda.synchronize()

    def free_all_blocks(self):
        torch.cuda.empty_cache()


backend = Backend()


def asxp(array, dtype=None):
    """Host or device array -> contiguous device tensor (matrix.py:314-322)."""
    if array is None:
        return None
    if hasattr(array, "array") and not isinstance(array, (np.ndarray, torch.Tensor)):
        array = array.array
    if isinstance(array, torch.Tensor):
        t = array
        if not t.is_cuda:
            t = t.to(backend.device)
    else:
        t = torch.from_numpy(np.ascontiguousarray(array)).to(backend.device)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    elif t.dtype not in (torch.float64, torch.complex128):
        t = t.to(torch.complex128 if t.is_complex() else torch.float64)
    if t.is_conj():          # a lazy conj view shares storage with its source: kernels read raw memory
        t = t.resolve_conj()
    return t.contiguous()


def asnumpy(array):
    """Device tensor -> NumPy array (matrix.py:298-311)."""
    if array is None:
        return None
    if isinstance(array, torch.Tensor):
        return array.detach().cpu().numpy()
    return np.asarray(array)

"""ctypes binding of librn_b200.so (the C ABI declared in include/rn_b200.h).

The product path has NO fallback: if the shared library is missing or the device is not an
sm_100 part, every compute entry point raises.
"""
import ctypes
import os
from ctypes import c_int, c_long, c_void_p, c_char_p, POINTER, c_double

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librn_b200.so")

_lib = None
_inited_devices = {}

_vp, _i, _l = c_void_p, c_int, c_long

# name -> (restype, argtypes); must list every symbol declared in include/rn_b200.h
SIGNATURES = {
    "rn_init": (_i, [_i, POINTER(_i), POINTER(_i)]),
    "rn_version": (c_char_p, []),
    "rn_dgemm_tn": (_i, [_vp, _i, _i, _i, _vp, _l, _vp, _l, _vp, _l, _i, _i, _l, _l, _l]),
    "rn_matmul": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _i]),
    "rn_ozaki_gemm_tn": (_i, [_vp, _i, _i, _i, _vp, _l, _vp, _l, _vp, _l, _i]),
    "rn_set_ozaki": (_i, [_i, c_double]),
    "rn_pack": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _l, _l, _vp, _l]),
    "rn_wapply": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _l, _l, _l, _l, _i, _i, _i,
                       _l, _l, _l, _l, _l, _vp, _vp, _vp]),
    "rn_hop_plan_create": (_i, [POINTER(_vp), _vp, _i, _i, _vp, _i, _i, _i, _vp, _i, _i, _i,
                                _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i]),
    "rn_hop_apply": (_i, [_vp, _vp, _vp, _vp]),
    "rn_hop_plan_launches": (_l, [_vp]),
    "rn_hop_plan_destroy": (_i, [_vp, _vp]),
    "rn_env_update": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _vp, _vp, _i, _i, _i, _i, _i,
                           _vp, _vp, _vp, _vp, _i]),
    "rn_qr": (_i, [_vp, _i, _i, _i, _vp, _l, _vp, _l, _vp, _l]),
    "rn_lq": (_i, [_vp, _i, _i, _i, _vp, _l, _vp, _l, _vp, _l]),
    "rn_svd_jacobi": (_i, [_vp, _i, _i, _i, _vp, _l, _vp, _l, _vp, _vp, _l, _i, POINTER(_i)]),
    "rn_svd_host": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _i]),
    "rn_svd": (_i, [_vp, _i, _i, _i, _vp, _l, _vp, _l, _vp, _vp, _l, _i, _i, POINTER(_i)]),
    "rn_multi_dot": (_i, [_vp, _i, _l, _i, _vp, _l, _vp, _vp, _vp]),
    "rn_lanczos_update": (_i, [_vp, _l, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rn_scale_inv": (_i, [_vp, _l, _vp, _vp, _vp]),
    "rn_lincomb": (_i, [_vp, _i, _l, _i, _vp, _l, _vp, _vp]),
    "rn_allclose": (_i, [_vp, _i, _l, _vp, _vp, c_double, c_double, _vp]),
    "rn_lanczos_step": (_i, [_vp, _vp, _l, _vp, _i, _vp, _vp, _vp, _vp]),
    "rn_expm_krylov": (_i, [_vp, _vp, _i, _l, _vp, c_double, c_double, _vp, POINTER(_i)]),
    "rn_krylov_max_dim": (_i, []),
    "rn_hop_apply_host": (_i, [_i, _i, _vp, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _i,
                               _vp, _i, _vp, _i, _vp, _vp, _i]),
    "rn_launch_count": (_l, []),
    "rn_profile_begin": (_i, []),
    "rn_profile_end": (_i, [POINTER(c_double), POINTER(c_double), POINTER(c_long)]),
    "rn_davidson": (_i, [POINTER(_vp), _i, _vp, _i, _l, _vp, _vp, _vp, c_double, c_double, _i, _i, c_double,
                        _vp, POINTER(c_double), POINTER(_i), POINTER(_i)]),
    "rn_int8_peak": (_i, [_vp, _i, POINTER(c_double)]),
    "rn_env_update_host": (_i, [_i, _i, _vp, _i, _i, _i, _vp, _vp, _i, _i, _i, _i,
                                _vp, _i, _i, _vp, _i]),
}

REDUCE_BLOCKS = 296


class RnError(RuntimeError):
    pass


def load():
    """Load the shared library and bind every symbol (no GPU needed)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RnError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  renormalizer_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_fast = {"lib": None, "dev": -1, "raw_stream": None}


def get(device_index=None):
    """Library handle, initialised for the given (default: current) CUDA device."""
    import torch
    if device_index is None and _fast["lib"] is not None and _fast["dev"] == torch.cuda.current_device():
        return _fast["lib"]
    if not torch.cuda.is_available():
        raise RnError("renormalizer_b200 needs a CUDA device (B200, sm_100a); none is available "
                      "and there is no CPU fallback")
    lib = load()
    if device_index is None:
        device_index = torch.cuda.current_device()
    if device_index not in _inited_devices:
        sm, cc = c_int(0), c_int(0)
        err = lib.rn_init(device_index, ctypes.byref(sm), ctypes.byref(cc))
        if err:
            raise RnError(f"rn_init failed on device {device_index} (cuda error {err})")
        _inited_devices[device_index] = (sm.value, cc.value)
    _fast["lib"], _fast["dev"] = lib, device_index
    _fast["raw_stream"] = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    return lib


def check(err, what):
    if err:
        raise RnError(f"{what} failed with CUDA error {err}")


def stream_ptr():
    """cudaStream_t of torch's current stream on the initialised device (fast path: no Python-side
    device bookkeeping)."""
    import torch
    dev = torch.cuda.current_device()
    if dev != _fast["dev"]:
        get(dev)                    # a torch.cuda.set_device since the last call: re-key on it
    raw = _fast["raw_stream"]
    if raw is not None:
        return raw(dev)
    return torch.cuda.current_stream().cuda_stream


class LaunchCounter:
    """Kernels launched by librn_b200.so (counted inside the library at every launch site);
    bench.py reports the difference over the timed region as gpu_launches."""

    @classmethod
    def total(cls):
        return int(load().rn_launch_count())

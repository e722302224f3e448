"""Thin Python wrappers over the C ABI: device tensors in, device tensors out.

Everything here launches kernels from librn_b200.so; torch is used for allocation and streams.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check, stream_ptr, LaunchCounter
from .backend import backend, asxp


def _is_cplx(t):
    return 1 if t.dtype == torch.complex128 else 0


def _es(t):
    return 2 if t.dtype == torch.complex128 else 1


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _dense(t):
    """What the C ABI reads is t.data_ptr(): a lazily conjugated view (torch's conj bit, which
    `.contiguous()` keeps on an already contiguous tensor) must be materialised first."""
    if t.is_conj():
        t = t.resolve_conj()
    return t if t.is_contiguous() else t.contiguous()


def _check_dev(*ts):
    for t in ts:
        if t is None:
            continue
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous()):
            raise ValueError("expected contiguous CUDA tensors")
        if t.is_conj() or t.is_neg():
            raise ValueError("lazily conjugated / negated view passed to a kernel (use _dense)")
        if t.dtype not in (torch.float64, torch.complex128):
            raise ValueError(f"unsupported dtype {t.dtype}")


def promote(*ts):
    """Bring tensors to a common dtype (float64 unless any is complex128)."""
    cplx = any(t.dtype == torch.complex128 for t in ts)
    dt = torch.complex128 if cplx else torch.float64
    return [t if t.dtype == dt else t.to(dt) for t in ts]


def gemm_tn(a, b, m, n, k, lda, ldb, out=None, ldc=None, accumulate=False):
    """out[i, j] (+)= sum_k a[i*lda+k] b[j*ldb+k] on float64 storage (real views)."""
    lib = _lib.get()
    if out is None:
        out = torch.empty((m, n), dtype=torch.float64, device=a.device)
        ldc = n
    check(lib.rn_dgemm_tn(stream_ptr(), m, n, k, _ptr(a), lda, _ptr(b), ldb, _ptr(out), ldc,
                          1 if accumulate else 0, 1, 0, 0, 0), "rn_dgemm_tn")
    return out


def ozaki_gemm_tn(a, b, m, n, k, lda, ldb, nslices=7, out=None, ldc=None):
    """Same product as gemm_tn on the tcgen05 int8 split path (rn_ozaki_gemm_tn)."""
    lib = _lib.get()
    if out is None:
        out = torch.empty((m, n), dtype=torch.float64, device=a.device)
        ldc = n
    check(lib.rn_ozaki_gemm_tn(stream_ptr(), m, n, k, _ptr(a), lda, _ptr(b), ldb, _ptr(out), ldc,
                               nslices), "rn_ozaki_gemm_tn")
    return out


def pack(src, rows, cols, s_row, s_col, mode=0, conj=False):
    """Strided 2-D view of `src` -> K-major real operand (see rn_pack)."""
    lib = _lib.get()
    cplx = _is_cplx(src)
    es = 2 if cplx else 1
    out_rows = rows * (2 if (cplx and mode == 1) else 1)
    dst = torch.empty((out_rows, cols * es), dtype=torch.float64, device=src.device)
    check(lib.rn_pack(stream_ptr(), cplx, mode, 1 if conj else 0, rows, cols, _ptr(src),
                      s_row, s_col, _ptr(dst), cols * es), "rn_pack")
    return dst


def matmul(a, b):
    """a (M,K) @ b (K,N) for real / complex tensors through pack + the K-major GEMM
    (xp.tensordot with one contracted axis, matrix.py:210)."""
    a, b = promote(_dense(a), _dense(b))
    _check_dev(a, b)
    M, K = a.shape
    K2, N = b.shape
    assert K == K2
    cplx = _is_cplx(a)
    es = 2 if cplx else 1
    out = torch.empty((M, N), dtype=a.dtype, device=a.device)
    if M == 0 or N == 0:
        return out
    if K == 0:
        return out.zero_()
    check(_lib.get().rn_matmul(stream_ptr(), cplx, M, K, N, _ptr(a), _ptr(b), _ptr(out), backend.gemm_path),
          "rn_matmul")
    return out


def tensordot1(a, b):
    """tensordot(a, b, axes=1): contract the last axis of a with the first axis of b."""
    sa, sb = a.shape, b.shape
    out = matmul(a.reshape(-1, sa[-1]), b.reshape(sb[0], -1))
    return out.reshape(tuple(sa[:-1]) + tuple(sb[1:]))


class MpoSite:
    """One MPO site tensor W[b, up, down, f] (real) with its CSR forms on the device.

    orientation 0 (hop, left environment):  W'[p=b, D=up, q=down, F=f]
    orientation 1 (right environment)    :  W'[p=f, D=up, q=down, F=b]
    """

    def __init__(self, w):
        if isinstance(w, torch.Tensor):
            w = w.detach().cpu().numpy()
        w = np.asarray(w)
        if np.iscomplexobj(w):
            if np.abs(w.imag).max() > 0:
                raise NotImplementedError("complex MPO site tensors are outside the accelerated path")
            w = w.real
        self.array = np.ascontiguousarray(w, dtype=np.float64)
        assert self.array.ndim == 4 and self.array.shape[1] == self.array.shape[2]
        self.shape = self.array.shape
        self._csr = {}
        self._dense = None

    @property
    def dense(self):
        if self._dense is None:
            self._dense = asxp(self.array)
        return self._dense

    def csr(self, orientation):
        if orientation not in self._csr:
            w = self.array
            wb, d, _, wf = w.shape
            if orientation == 0:
                t = w.transpose(1, 3, 0, 2)      # up, f, b, down
            else:
                t = w.transpose(1, 0, 3, 2)      # up, b, f, down
            D, F, P, Q = t.shape
            flat = t.reshape(D * F, P * Q)
            rows, cols = np.nonzero(flat)
            rowptr = np.zeros(D * F + 1, dtype=np.int32)
            np.add.at(rowptr, rows + 1, 1)
            rowptr = np.cumsum(rowptr).astype(np.int32)
            pq = cols.astype(np.int32)
            val = flat[rows, cols].astype(np.float64)
            if len(pq) == 0:
                pq = np.zeros(1, dtype=np.int32)
                val = np.zeros(1, dtype=np.float64)
            dev = backend.device
            self._csr[orientation] = (F, torch.from_numpy(rowptr).to(dev),
                                      torch.from_numpy(pq).to(dev), torch.from_numpy(val).to(dev))
        return self._csr[orientation]


_site_cache = {}


def as_mpo_site(w):
    if isinstance(w, MpoSite):
        return w
    if hasattr(w, "array") and isinstance(getattr(w, "array"), MpoSite):
        return w.array
    key = id(w)
    hit = _site_cache.get(key)
    if hit is not None and hit[0] is w:
        return hit[1]
    site = MpoSite(w)
    if len(_site_cache) > 4096:
        _site_cache.clear()
    _site_cache[key] = (w, site)
    return site


class HopPlan:
    """Device plan for H_eff . C (rn_hop_plan_*)."""

    def __init__(self, ltensor, rtensor, sites, cshape, dtype, path=None):
        lib = _lib.get()
        self.lib = lib
        nsite = len(sites)
        cshape = tuple(int(x) for x in cshape)
        ancilla = nsite > 0 and (2 * nsite + 2 == len(cshape))
        if not ancilla and nsite + 2 != len(cshape):
            raise ValueError(f"cshape {cshape} does not match {nsite} centre sites")
        self.dtype = dtype
        self.L = _dense(ltensor if ltensor.dtype == dtype else ltensor.to(dtype))
        self.R = _dense(rtensor if rtensor.dtype == dtype else rtensor.to(dtype))
        _check_dev(self.L, self.R)
        La, Lb, Lc = self.L.shape
        Rl, Rf, Rk = self.R.shape
        d = [1, 1]
        g = [1, 1]
        for i in range(nsite):
            if ancilla:
                d[i], g[i] = cshape[1 + 2 * i], cshape[2 + 2 * i]
            else:
                d[i] = cshape[1 + i]
        if cshape[0] != Lc or cshape[-1] != Rk:
            raise ValueError("cshape bonds do not match the environments")
        self.sites = sites
        csr = [s.csr(0) for s in sites]
        for i, s in enumerate(sites):
            if s.shape[1] != d[i]:
                raise ValueError("MPO physical dimension does not match cshape")
        self._keep = csr
        null = (0, None, None, None)
        c1 = csr[0] if nsite >= 1 else null
        c2 = csr[1] if nsite == 2 else null
        self.out_shape = (La,) + cshape[1:-1] + (Rl,)
        self.in_shape = cshape
        handle = ctypes.c_void_p()
        path = backend.gemm_path if path is None else path
        check(lib.rn_hop_plan_create(ctypes.byref(handle), stream_ptr(), 1 if dtype == torch.complex128 else 0,
                                     nsite, _ptr(self.L), La, Lb, Lc, _ptr(self.R), Rl, Rf, Rk,
                                     d[0], g[0], d[1], g[1],
                                     c1[0], _ptr(c1[1]), _ptr(c1[2]), _ptr(c1[3]),
                                     c2[0], _ptr(c2[1]), _ptr(c2[2]), _ptr(c2[3]), path),
              "rn_hop_plan_create")
        self.handle = handle
        self.nlaunch = 2 + nsite + 1

    def apply(self, c, out=None):
        if c.dtype != self.dtype:
            if c.is_complex():
                raise ValueError("complex centre tensor on a real H_eff plan (the imaginary part would be dropped)")
            c = c.to(self.dtype)
        c = _dense(c)
        _check_dev(c)
        if out is None:
            out = torch.empty(self.out_shape, dtype=self.dtype, device=c.device)
        check(self.lib.rn_hop_apply(self.handle, stream_ptr(), _ptr(c), _ptr(out)), "rn_hop_apply")
        return out

    def close(self):
        if getattr(self, "handle", None):
            self.lib.rn_hop_plan_destroy(self.handle, stream_ptr())
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def env_update(environ, bra, ket, site, domain, path=None):
    """rn_env_update: absorb one site into a left/right environment (device tensors).
    `bra` is the un-conjugated bra-side site tensor."""
    lib = _lib.get()
    environ, bra, ket = promote(environ, bra, ket)
    environ, bra, ket = _dense(environ), _dense(bra), _dense(ket)
    _check_dev(environ, bra, ket)
    cplx = _is_cplx(ket)
    Ea, Eb, Ec = environ.shape
    if ket.ndim == 3:
        d, g = ket.shape[1], 1
    elif ket.ndim == 4:
        d, g = ket.shape[1], ket.shape[2]
    else:
        raise ValueError(f"MPS ndim is not 3 or 4, got {ket.ndim}")
    if domain == "L":
        dom = 0
        Mf, Mh = bra.shape[-1], ket.shape[-1]
        assert bra.shape[0] == Ea and ket.shape[0] == Ec and site.shape[0] == Eb
    elif domain == "R":
        dom = 1
        Mf, Mh = bra.shape[0], ket.shape[0]
        assert bra.shape[-1] == Ea and ket.shape[-1] == Ec and site.shape[-1] == Eb
    else:
        raise ValueError("domain must be 'L' or 'R'")
    F, rowptr, pq, val = site.csr(dom)
    out = torch.empty((Mf, F, Mh), dtype=ket.dtype, device=ket.device)
    path = backend.gemm_path if path is None else path
    check(lib.rn_env_update(stream_ptr(), cplx, dom, _ptr(environ), Ea, Eb, Ec, _ptr(bra), _ptr(ket),
                            d, g, Mf, Mh, F, _ptr(rowptr), _ptr(pq), _ptr(val), _ptr(out), path),
          "rn_env_update")
    return out


def qr(a, lq=False):
    """Householder QR (or LQ) of a 2-D device tensor; returns (Q, R) or (L, Q)."""
    lib = _lib.get()
    a = _dense(a)
    _check_dev(a)
    m, n = a.shape
    k = min(m, n)
    cplx = _is_cplx(a)
    if not lq:
        q = torch.empty((m, k), dtype=a.dtype, device=a.device)
        r = torch.empty((k, n), dtype=a.dtype, device=a.device)
        check(lib.rn_qr(stream_ptr(), cplx, m, n, _ptr(a), n, _ptr(q), k, _ptr(r), n), "rn_qr")
        return q, r
    l = torch.empty((m, k), dtype=a.dtype, device=a.device)
    q = torch.empty((k, n), dtype=a.dtype, device=a.device)
    check(lib.rn_lq(stream_ptr(), cplx, m, n, _ptr(a), n, _ptr(l), k, _ptr(q), n), "rn_lq")
    return l, q


# below this many columns the bare Jacobi iteration is cheaper than the two QR factorisations
SVD_PRECONDITION_MIN = 48


def svd(a, max_sweeps=40, sort=True, precondition=None):
    """SVD of a bond-matrix block; returns (U, S, Vh) with S sorted descending when sort.

    rn_svd: one-sided Jacobi, preconditioned as in Drmac & Veselic (SIAM J. Matrix Anal. Appl. 29,
    1322): the columns of the tall orientation T are ordered by decreasing norm, T P = Q1 R1,
    R1^H = Q2 R2, and the (block) Jacobi rotations run on X = R2^H (k x k) instead of T.  Bond
    matrices of a converged state are strongly graded (singular values decay exponentially); on
    those the bare iteration (rn_svd_jacobi, `precondition=False`) needs > 40 sweeps over the long
    columns of T while X needs 6-9 sweeps over columns of length k."""
    lib = _lib.get()
    a = _dense(a)
    _check_dev(a)
    m, n = a.shape
    k = min(m, n)
    cplx = _is_cplx(a)
    if precondition is None:
        precondition = k >= SVD_PRECONDITION_MIN
    u = torch.empty((m, k), dtype=a.dtype, device=a.device)
    s = torch.empty((k,), dtype=torch.float64, device=a.device)
    vh = torch.empty((k, n), dtype=a.dtype, device=a.device)
    sweeps = ctypes.c_int(0)
    if precondition:
        check(lib.rn_svd(stream_ptr(), cplx, m, n, _ptr(a), n, _ptr(u), k, _ptr(s), _ptr(vh), n,
                         max_sweeps, backend.gemm_path, ctypes.byref(sweeps)), "rn_svd")
    else:
        check(lib.rn_svd_jacobi(stream_ptr(), cplx, m, n, _ptr(a), n, _ptr(u), k, _ptr(s), _ptr(vh), n,
                                max_sweeps, ctypes.byref(sweeps)), "rn_svd_jacobi")
    svd.last_sweeps = abs(sweeps.value)
    if sweeps.value < 0:
        # the last sweep still rotated: the factors are not orthogonal to working precision
        if not precondition:
            return svd(a, max_sweeps=max_sweeps, sort=sort, precondition=True)
        if max_sweeps < 120:
            return svd(a, max_sweeps=120, sort=sort, precondition=True)
        raise _lib.RnError(f"rn_svd: Jacobi iteration on a {m} x {n} block did not converge in {max_sweeps} sweeps")
    if sort:
        order = torch.argsort(s, descending=True)
        u, s, vh = u.index_select(1, order), s.index_select(0, order), vh.index_select(0, order)
    return u, s, vh


class VecWorkspace:
    """Scratch for the reductions of the Krylov / Davidson kernels."""

    def __init__(self, device, nvec_max=64):
        self.nvec_max = nvec_max
        self.ws = torch.empty(2 * _lib.REDUCE_BLOCKS * nvec_max, dtype=torch.float64, device=device)


def multi_dot(V, x, nvec, n, cplx, ws, out=None):
    """out[i] = <V[i], x> for the first nvec rows of the 2-D stack V (complex -> 2 doubles)."""
    lib = _lib.get()
    assert nvec <= ws.nvec_max
    if out is None:
        out = torch.empty(2 * nvec, dtype=torch.float64, device=x.device)
    es = 2 if cplx else 1
    ld = V.stride(0) * es if V.ndim == 2 else n * es
    check(lib.rn_multi_dot(stream_ptr(), 1 if cplx else 0, n, nvec, _ptr(V), ld, _ptr(x), _ptr(ws.ws),
                           _ptr(out)), "rn_multi_dot")
    return out


def lincomb(V, coef, nvec, n, cplx, out):
    """out = sum_i coef[i] V[i]; coef is a device tensor (complex when cplx)."""
    lib = _lib.get()
    es = 2 if cplx else 1
    ld = V.stride(0) * es
    check(lib.rn_lincomb(stream_ptr(), 1 if cplx else 0, n, nvec, _ptr(V), ld, _ptr(coef), _ptr(out)),
          "rn_lincomb")
    return out


def lanczos_update(w, vj, vjm1, alpha, beta_prev, ws, beta_out):
    lib = _lib.get()
    nd = w.numel() * _es(w)
    check(lib.rn_lanczos_update(stream_ptr(), nd, _ptr(w), _ptr(vj), _ptr(vjm1), _ptr(alpha),
                                _ptr(beta_prev), _ptr(ws.ws), _ptr(beta_out)), "rn_lanczos_update")


def scale_inv(x, s, out):
    lib = _lib.get()
    nd = x.numel() * _es(x)
    check(lib.rn_scale_inv(stream_ptr(), nd, _ptr(x), _ptr(s), _ptr(out)), "rn_scale_inv")


def lanczos_step(plan, n, V, j, alpha, beta, w, ws):
    """One fused Lanczos iteration on the Krylov stack V through rn_lanczos_step."""
    check(plan.lib.rn_lanczos_step(plan.handle, stream_ptr(), n, _ptr(V), j, _ptr(alpha), _ptr(beta),
                                   _ptr(w), _ptr(ws.ws)), "rn_lanczos_step")


_CUDA_ERROR_NOT_SUPPORTED = 801


def expm_krylov_plan(plan, v, dt):
    """expm(dt * H_eff) v through rn_expm_krylov (whole Lanczos loop in one C call).
    Returns (result, H_eff applications) or None when the Krylov dimension outgrows the device
    eigen-solver (the caller then iterates step by step)."""
    out = torch.empty_like(v)
    nsteps = ctypes.c_int(0)
    dt = complex(dt)
    err = plan.lib.rn_expm_krylov(plan.handle, stream_ptr(), _is_cplx(v), v.numel(), _ptr(v),
                                  dt.real, dt.imag, _ptr(out), ctypes.byref(nsteps))
    if err == _CUDA_ERROR_NOT_SUPPORTED:
        return None
    check(err, "rn_expm_krylov")
    return out, nsteps.value


def allclose(a, b, rtol=1e-5, atol=1e-8, flag=None):
    """numpy.allclose(a, b) for two device vectors of the same dtype (one D2H int read)."""
    lib = _lib.get()
    if flag is None:
        flag = torch.empty(1, dtype=torch.int32, device=a.device)
    check(lib.rn_allclose(stream_ptr(), _is_cplx(a), a.numel(), _ptr(a), _ptr(b), rtol, atol, _ptr(flag)),
          "rn_allclose")
    return int(flag.item()) == 0


def davidson_plans(plans, x0, mask_u8, hdiag, inverse=1.0, tol=1e-12, max_cycle=100, max_space=12, lindep=1e-14):
    """Lowest eigenpair of inverse * sum_p H_eff[p] in the masked subspace through rn_davidson (the
    whole Davidson iteration in one C call).  Returns (e, c, H_eff applications, converged)."""
    lib = _lib.get()
    x0 = _dense(x0.reshape(-1))
    hdiag = _dense(hdiag.reshape(-1))
    _check_dev(x0, hdiag)
    if hdiag.dtype != torch.float64:
        raise ValueError("hdiag must be real")
    n = x0.numel()
    handles = (ctypes.c_void_p * len(plans))(*[p.handle for p in plans])
    out = torch.empty_like(x0)
    e, nhop, conv = ctypes.c_double(0), ctypes.c_int(0), ctypes.c_int(0)
    check(lib.rn_davidson(handles, len(plans), stream_ptr(), _is_cplx(x0), n, _ptr(x0),
                          _ptr(mask_u8) if mask_u8 is not None else None, _ptr(hdiag), float(inverse), tol,
                          max_cycle, max_space, lindep, _ptr(out), ctypes.byref(e), ctypes.byref(nhop),
                          ctypes.byref(conv)), "rn_davidson")
    return e.value, out, nhop.value, bool(conv.value)

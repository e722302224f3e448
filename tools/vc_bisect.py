"""Diagnostic: run the variational-compression golden case on the GPU with every C-ABI op wrapped by
a torch-on-CUDA cross-check (and optionally replaced by it), to find the op that deviates.
usage: python tools/vc_bisect.py [1site|2site] [replace=svd,qr,env,hop,matmul]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_mpo, load_oracle_mps
from renormalizer_b200 import ops
import renormalizer_b200.mps as mpsmod, renormalizer_b200.hop_expr as hopmod, renormalizer_b200.gs as gsmod
from renormalizer_b200.configs import CompressConfig, CompressCriteria, EvolveConfig, EvolveMethod
from renormalizer_b200.mpo import Mpo
from renormalizer_b200.mps import Mps

method = sys.argv[1] if len(sys.argv) > 1 else "1site"
replace = set()
for a in sys.argv[2:]:
    if a.startswith("replace="):
        replace = set(a.split("=", 1)[1].split(","))
LOG = []


def rel(a, b):
    d = (a - b).abs().max().item() if a.numel() else 0.0
    return d / max(b.abs().max().item() if b.numel() else 0.0, 1e-300)


_svd, _qr, _env, _matmul = ops.svd, ops.qr, ops.env_update, ops.matmul


def svd(a, **kw):
    tu, ts, tvh = torch.linalg.svd(a, full_matrices=False)
    if "svd" in replace:
        svd.last_sweeps = 1
        return tu, ts, tvh
    u, s, vh = _svd(a, **kw)
    k = s.numel()
    eye = torch.eye(k, dtype=u.dtype, device=u.device)
    ou = (u.conj().T @ u - eye).abs().max().item() if k else 0
    ov = (vh @ vh.conj().T - eye).abs().max().item() if k else 0
    rec = rel((u * s.to(u.dtype)) @ vh, a) if k and a.abs().max() > 0 else 0
    ds = (s - ts).abs().max().item() / max(ts.max().item(), 1e-300) if k else 0
    LOG.append(("svd", tuple(a.shape), ou, ov, rec, ds, (s / max(s.max().item(), 1e-300)).cpu().numpy()))
    return u, s, vh


def qr(a, lq=False):
    if "qr" in replace:
        if not lq:
            return torch.linalg.qr(a)
        q, r = torch.linalg.qr(a.conj().T)
        return r.conj().T.contiguous(), q.conj().T.contiguous()
    x, y = _qr(a, lq=lq)
    q = y if lq else x
    k = min(a.shape)
    eye = torch.eye(k, dtype=a.dtype, device=a.device)
    oq = ((q @ q.conj().T if lq else q.conj().T @ q) - eye).abs().max().item()
    LOG.append(("lq" if lq else "qr", tuple(a.shape), oq, rel(x @ y, a) if a.abs().max() > 0 else 0))
    return x, y


def env_ref(environ, bra, ket, site, domain):
    w = site.dense.to(ket.dtype) if not isinstance(site, torch.Tensor) else site
    e, b, k = environ.to(ket.dtype), bra.to(ket.dtype).conj(), ket
    if ket.ndim == 3:
        if domain == "L":
            return torch.einsum("abc,adf,bdeg,ceh->fgh", e, b, w, k)
        return torch.einsum("fda,gdeb,hec,abc->fgh", b, w, k, e)
    if domain == "L":
        return torch.einsum("abc,adxf,bdeg,cexh->fgh", e, b, w, k)
    return torch.einsum("fdxa,gdeb,hexc,abc->fgh", b, w, k, e)


def env_update(environ, bra, ket, site, domain, path=None):
    environ, bra, ket = ops.promote(environ, bra, ket)
    ref = env_ref(environ, bra, ket, site, domain)
    if "env" in replace:
        return ref.contiguous()
    got = _env(environ, bra, ket, site, domain, path=path)
    LOG.append(("env" + domain, tuple(environ.shape), tuple(bra.shape), tuple(ket.shape), rel(got, ref)))
    return got


def matmul(a, b):
    a2, b2 = ops.promote(a, b)
    ref = a2 @ b2
    if "matmul" in replace:
        return ref
    got = _matmul(a, b)
    if ref.numel() and ref.abs().max() > 0:
        LOG.append(("matmul", tuple(a.shape), tuple(b.shape), rel(got, ref)))
    return got


_hop = mpsmod.hop_expr_dtype


class Hop:
    def __init__(self, l, r, cmo, shape, dtype):
        self.args = (l, r, cmo, tuple(shape), dtype)
        self.real = None if "hop" in replace else _hop(l, r, cmo, shape, dtype)

    def ref(self, c):
        l, r, cmo, shape, dtype = self.args
        l, r, c = l.to(dtype), r.to(dtype), c.reshape(shape).to(dtype)
        ws = [ops.as_mpo_site(m).dense.to(dtype) for m in cmo]
        if len(ws) == 1:
            return torch.einsum("abc,bdef,lfk,cek->adl", l, ws[0], r, c)
        return torch.einsum("abc,bdef,fghi,lik,cegk->adhl", l, ws[0], ws[1], r, c)

    def __call__(self, c):
        ref = self.ref(c)
        if self.real is None:
            return ref.contiguous()
        got = self.real(c)
        LOG.append(("hop", tuple(self.args[3]), rel(got, ref)))
        return got

    def close(self):
        if self.real is not None:
            self.real.close()


ops.svd, ops.qr, ops.env_update, ops.matmul = svd, qr, env_update, matmul
mpsmod.hop_expr_dtype = lambda l, r, cmo, shape, dtype: Hop(l, r, cmo, shape, dtype)

g = np.load(os.path.join(ROOT, "tests", "golden", "vcompress.npz"))
n = int(g["mpo_n"])
mpo = Mpo(load_mpo(g), qn=[g[f"mpo_qn_{i}"] for i in range(n + 1)], qntot=g["mpo_qntot"],
          qnidx=int(g["mpo_qnidx"]), sigmaqn=[g[f"mpo_sigmaqn_{i}"] for i in range(n)],
          to_right=bool(g["mpo_to_right"]))
om = load_oracle_mps(g, "mps", meta="mps")
state = Mps(om.sites, om.qn, om.sigmaqn, om.qntot, om.qnidx, om.to_right)
state.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps)
state.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=8, vmethod=method)
np.random.seed(0)
new = mpo.contract(state, algo="variational")
print(f"method {method} replace {sorted(replace)}: bond_dims {new.bond_dims} ref {list(g[f'{method}_bond_dims'])}")
print(f"  norm {new.mp_norm:.10f}  ref {float(g[f'{method}_norm']):.10f}")
np.set_printoptions(precision=2, linewidth=200)
bad = 0
for e in LOG:
    errs = [x for x in e[1:] if isinstance(x, float)]
    if errs and max(errs) > 1e-9:
        bad += 1
        if bad <= 25:
            print("  DEVIATION", e)
print(f"  {len(LOG)} ops checked, {bad} deviating")

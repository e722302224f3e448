"""Diagnostic for the omega-targeting sweep (not part of the product)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
from helpers import load_mpo, load_oracle_mps
from renormalizer_b200.mpo import Mpo
from renormalizer_b200.mps import Mps
from renormalizer_b200 import gs as dgs
from renormalizer_b200.lib import Environ
from oracle import contract as oc
g = np.load("tests/golden/holstein.npz")
mpo = Mpo(load_mpo(g))
sq = mpo.add(Mpo.identity_like(mpo).scale(-float(g["omega"]))).squared()
om = load_oracle_mps(g, "mps0")
mps = Mps(om.sites, om.qn, om.sigmaqn, om.qntot, om.qnidx, om.to_right)
mps.optimize_config.procedure = [[int(a), float(b)] for a, b in g["procedure"]]
mps.optimize_config.method = "2site"
orig = dgs.eigh_iterative
def spy(mps_, qn_mask, l, r, cmo, guess):
    gn = float(torch.linalg.vector_norm(guess))
    mk = int(qn_mask.sum())
    # compare device hop with the oracle on the guess
    from renormalizer_b200.hop_expr import hop_expr_dtype
    hop = hop_expr_dtype(l, r, cmo, qn_mask.shape, torch.float64)
    got = hop(guess.reshape(qn_mask.shape)).cpu().numpy()
    ref = oc.hop_apply(l.cpu().numpy(), r.cpu().numpy(), [c.array for c in cmo], guess.cpu().numpy().reshape(qn_mask.shape))
    err = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)
    print(f"site shape {qn_mask.shape} allowed {mk} |guess| {gn:.3e} hop relerr {err:.2e} |L| {float(l.abs().max()):.3e} |R| {float(r.abs().max()):.3e}", flush=True)
    e, c, nh = orig(mps_, qn_mask, l, r, cmo, guess)
    print(f"   -> e {e:.6e} |c| {float(torch.linalg.vector_norm(c)):.3e} nhop {nh}", flush=True)
    return e, c, nh
dgs.eigh_iterative = spy
od = dgs.eigh_direct
def spyd(mps_, qn_mask, l, r, cmo):
    from renormalizer_b200.hop_expr import hop_expr_dtype
    rng = np.random.default_rng(0)
    x = rng.standard_normal(qn_mask.shape)
    hop = hop_expr_dtype(l, r, cmo, qn_mask.shape, torch.float64)
    got = hop(torch.from_numpy(x).cuda()).cpu().numpy()
    ref = oc.hop_apply(l.cpu().numpy(), r.cpu().numpy(), [c.array for c in cmo], x)
    u = np.zeros(qn_mask.size); iu = int(np.nonzero(qn_mask.reshape(-1))[0][3]); u[iu] = 1.0
    gotu = hop(torch.from_numpy(u.reshape(qn_mask.shape)).cuda()).cpu().numpy()
    refu = oc.hop_apply(l.cpu().numpy(), r.cpu().numpy(), [c.array for c in cmo], u.reshape(qn_mask.shape))
    xt = torch.zeros(qn_mask.size, dtype=torch.float64, device="cuda"); xt[np.int64(iu)] = 1
    gotu2 = hop(xt.reshape(qn_mask.shape)).cpu().numpy()
    print(f"   unit vector {iu}: |got| {np.abs(gotu).max():.3e} |got2| {np.abs(gotu2).max():.3e} |ref| {np.abs(refu).max():.3e} inverse {mps_.optimize_config.inverse}", flush=True)
    print(f"   direct: L {tuple(l.shape)} R {tuple(r.shape)} W {[c.shape for c in cmo]} |L| {float(l.abs().max()):.2e} |R| {float(r.abs().max()):.2e} |ref| {np.abs(ref).max():.2e} relerr {np.abs(got-ref).max()/max(np.abs(ref).max(),1e-300):.2e}", flush=True)
    e, c = od(mps_, qn_mask, l, r, cmo)
    print(f"direct site shape {qn_mask.shape} e {e:.6e} |c| {float(torch.linalg.vector_norm(c)):.3e}", flush=True)
    return e, c
dgs.eigh_direct = spyd
from renormalizer_b200.backend import backend
backend.gemm_path = int(sys.argv[1]) if len(sys.argv) > 1 else 1
np.random.seed(99)
try:
    energies, opt = dgs.optimize_mps(mps, mpo, omega=float(g["omega"]))
    print(energies, opt.expectation(mpo), float(g["omega_expectation"]))
except Exception as ex:
    print("FAILED", repr(ex))

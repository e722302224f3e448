"""Diagnostic: host-side profile + per-phase CUDA time of one 2-site DMRG sweep (not part of the product).
    python tools/pyprof_dmrg.py [M] [nmol]"""
import cProfile, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from renormalizer_b200 import _lib, ops
from renormalizer_b200.backend import backend
from renormalizer_b200.configs import CompressConfig, CompressCriteria
from renormalizer_b200.gs import single_sweep
from renormalizer_b200.lib import Environ
from renormalizer_b200.mpo import Mpo
from renormalizer_b200.mps import Mps
M = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nmol = int(sys.argv[2]) if len(sys.argv) > 2 else 20
args = bench.argparse.Namespace(workload="holstein_dmrg", modes=20, levels=8, bond=M, dt=0.05, mols=nmol, orbitals=12, fmo_modes=2)
backend.gemm_path = 1
_lib.get()
w = bench.make_workload(sys.argv[3] if len(sys.argv) > 3 else "holstein_dmrg", M, args, 1234)
meta = w["meta"]
mps = Mps(w["sites"], meta["qn"], meta["sigmaqn"], meta["qntot"], meta["qnidx"], meta["to_right"])
mpo = Mpo(w["mpo"])
mps.optimize_config.method = "2site"
mps.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=M)
mps.ensure_right_canonical()
env = Environ(mps, mpo, "R")
for _ in range(int(sys.argv[4]) if len(sys.argv) > 4 else 2):
    single_sweep(mps, mpo, env, None, 0.0, None)
torch.cuda.synchronize()
# time the SVD calls of the next sweep with events
svd_ms, shapes = [], []
orig_svd = ops.svd
def timed_svd(a, *k, **kw):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); r = orig_svd(a, *k, **kw); e.record(); e.synchronize()
    svd_ms.append(s.elapsed_time(e)); shapes.append(tuple(a.shape)); return r
ops.svd = timed_svd
dav_ms = []
orig_dav = ops.davidson_plans
def timed_dav(*k, **kw):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); r = orig_dav(*k, **kw); e.record(); e.synchronize()
    dav_ms.append(s.elapsed_time(e)); return r
ops.davidson_plans = timed_dav
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
single_sweep(mps, mpo, env, None, 0.0, None)
torch.cuda.synchronize()
pr.disable()
wall = time.perf_counter() - t0
print(f"sweep wall {wall:.3f} s, {w['nsite'] - 1} site updates, svd total {sum(svd_ms) / 1e3:.3f} s over {len(svd_ms)} calls, "
      f"davidson total {sum(dav_ms) / 1e3:.3f} s over {len(dav_ms)} calls, hops {sum(mps.hop_counts)} ({mps.hop_counts})")
big = sorted(zip(svd_ms, shapes), reverse=True)[:8]
print("largest svd calls (ms, shape):", [(round(a, 1), s) for a, s in big])
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)

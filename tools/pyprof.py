"""Diagnostic: host-side profile of one TDVP-PS step (not part of the product)."""
import cProfile, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from renormalizer_b200 import _lib
from renormalizer_b200.backend import backend
from renormalizer_b200.mpo import Mpo
from renormalizer_b200.mps import Mps
args = bench.argparse.Namespace(workload="sbm_tdvp", modes=20, levels=8, bond=256, dt=0.05, mols=20)
backend.gemm_path = 1
_lib.get()
w = bench.make_workload(args, 1234)
meta = w["meta"]
mps = Mps(w["sites"], meta["qn"], meta["sigmaqn"], meta["qntot"], meta["qnidx"], meta["to_right"])
mpo = Mpo(w["mpo"])
from renormalizer_b200.configs import EvolveConfig, EvolveMethod
mps.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps)
for _ in range(2):
    mps = mps.evolve(mpo, 0.05)
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
mps = mps.evolve(mpo, 0.05)
torch.cuda.synchronize()
pr.disable()
print("step wall", time.perf_counter() - t0, "krylov steps", sum(mps.evolve_config.stat), len(mps.evolve_config.stat))
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)

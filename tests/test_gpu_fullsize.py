"""Sweep-level parity at the BENCHMARKED sizes (BASELINE.json configs[1] and [2]): full-size site updates on
the CUDA path against the CPU oracle on the same inputs -- energies to 1e-10, site tensors to 1e-8
(north_star's tolerances).  The oracle finishes one such update in seconds; the state the updates start
from is produced by the CUDA path (one warm-up step / sweep), its inputs are downloaded for the oracle."""
import argparse
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = [pytest.mark.gpu, pytest.mark.slow]

E_TOL = 1e-10
T_TOL = 1e-8


def _args(**kw):
    base = dict(modes=20, mols=20, levels=8, orbitals=12, fmo_modes=2, dt=0.05)
    base.update(kw)
    return argparse.Namespace(**base)


def _device_state(work):
    from renormalizer_b200.configs import CompressConfig, CompressCriteria, EvolveConfig, EvolveMethod
    from renormalizer_b200.mps import Mps
    meta = work["meta"]
    m = Mps(work["sites"], meta["qn"], meta["sigmaqn"], meta["qntot"], meta["qnidx"], meta["to_right"])
    m.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps)
    m.optimize_config.method = "2site"
    m.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=work["bond"])
    return m


def test_tdvp_ps_site_updates_m256_vs_oracle():
    """Spin-boson TDVP-PS, 20 modes x 8 levels, M = 256 (the headline bench workload): after one full step
    on the device, four site updates spread over the next half sweep are replayed by the oracle."""
    import bench
    from renormalizer_b200.backend import asnumpy
    from renormalizer_b200.hop_expr import hop_expr_dtype
    from renormalizer_b200.krylov import expm_krylov
    from renormalizer_b200.mpo import Mpo
    from oracle.contract import hop_apply
    from oracle.krylov import expm_krylov as oracle_expm
    work = bench.make_workload("sbm_tdvp", 256, _args(), seed=7)
    mpo = Mpo(work["mpo"])
    state = _device_state(work).evolve(mpo, work["dt"])
    e0, e1 = _device_state(work).expectation(mpo), state.expectation(mpo)
    assert abs(e1 - e0) < 1e-9 * max(1.0, abs(e0))                      # real-time TDVP conserves the energy
    assert abs(state.mp_norm - 1) < 1e-12
    chosen = bench.flop_quantile_sites(work, state.bond_dims, 4)
    caps = bench.gpu_capture(work, mpo, state, chosen)
    assert [c["imps"] for c in caps] == chosen and len(caps) == 4
    for cap in caps:
        t_cpu, scal, _ = bench.cpu_replay(work, cap)
        assert abs(scal[0] - cap["scalars"][0]) < E_TOL * max(1.0, abs(scal[0]))
        # the evolved centre tensor itself (same inputs, so the same gauge): element-wise and as an overlap
        shape = list(cap["site"].shape)
        w = work["mpo"][cap["imps"]]
        ref, nref = oracle_expm(lambda y: hop_apply(cap["lt"], cap["rt"], [w], y.reshape(shape)).ravel(),
                                -1j * work["dt"] / 2, cap["site"].astype(np.complex128).ravel())
        hop = hop_expr_dtype(torch.from_numpy(cap["lt"]).cuda(), torch.from_numpy(cap["rt"]).cuda(), [mpo[cap["imps"]]],
                             shape, torch.complex128)
        got, ngot = expm_krylov(hop, -1j * work["dt"] / 2, torch.from_numpy(cap["site"]).cuda().to(torch.complex128).reshape(-1))
        hop.close()
        got = asnumpy(got)
        assert ngot == nref                                             # same number of H_eff applications
        assert np.abs(got - ref).max() < T_TOL * np.abs(ref).max()
        assert abs(abs(np.vdot(ref, got)) / (np.linalg.norm(ref) * np.linalg.norm(got)) - 1) < T_TOL


def test_dmrg_site_update_m512_vs_oracle():
    """Holstein chain 2-site DMRG, 20 molecules x 8 levels, M = 512: after two sweeps on the device, one
    site update in the middle of the chain (Davidson + truncating SVD) against the oracle."""
    import bench
    from renormalizer_b200.gs import single_sweep
    from renormalizer_b200.lib import Environ
    from renormalizer_b200.mpo import Mpo
    work = bench.make_workload("holstein_dmrg", 512, _args(), seed=7)
    mpo = Mpo(work["mpo"])
    m = _device_state(work)
    m.ensure_right_canonical()
    env = Environ(m, mpo, "R")
    energies = []
    for _ in range(2):
        micro, _, _ = single_sweep(m, mpo, env, None, 0.0, None)
        energies.append(min(e for e, _ in micro))
    assert energies[1] <= energies[0] + 1e-10
    caps = bench.gpu_capture(work, mpo, m, [19])
    assert len(caps) == 1 and caps[0]["cidx"] == [19, 20]
    t_cpu, scal, _ = bench.cpu_replay(work, caps[0])
    assert abs(scal[0] - caps[0]["scalars"][0]) < E_TOL * max(1.0, abs(scal[0]))
    assert abs(scal[0] - energies[1]) < 1e-6                            # the converged sweep energy

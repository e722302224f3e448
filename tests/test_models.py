"""Host-side model builders (renormalizer_b200/models.py) and the sampling helpers of bench.py: the
numeric MPO construction against dense operators and against the reference's symbolic MPO, the
seeded / quantum-number-blocked initial states (CPU only)."""
import argparse
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from renormalizer_b200 import models
from helpers import load_mpo


def _dense_qc(h1e, h2e):
    """sum h1e a+_p a_q + sum h2e a+_p a+_q a_r a_s with the Jordan-Wigner matrices of h_qc.py:150-159."""
    n = h1e.shape[0]
    z, sp, sm, i2 = np.diag([1.0, -1.0]), np.diag([1.0], 1), np.diag([1.0], -1), np.eye(2)

    def kron(mats):
        out = np.eye(1)
        for m in mats:
            out = np.kron(out, m)
        return out
    a = [kron([z] * j + [sp] + [i2] * (n - j - 1)) for j in range(n)]
    ad = [kron([z] * j + [sm] + [i2] * (n - j - 1)) for j in range(n)]
    h = np.zeros((2 ** n, 2 ** n))
    for p, q in np.argwhere(h1e != 0):
        h += h1e[p, q] * ad[p] @ a[q]
    for p, q, r, s in np.argwhere(h2e != 0):
        h += h2e[p, q, r, s] * ad[p] @ ad[q] @ a[r] @ a[s]
    return h


def test_qc_mpo_equals_dense_jordan_wigner_operator():
    rng = np.random.default_rng(0)
    h1e, h2e = models.random_qc_integrals(3, rng)
    sites, bond_qn = models.qc_mpo(h1e, h2e)
    dense = _dense_qc(h1e, h2e)
    assert np.abs(dense - dense.T).max() < 1e-13                     # Hermitian by construction
    assert np.abs(models.mpo_to_dense(sites) - dense).max() < 1e-13
    # bond quantum numbers: one label per bond state, (0, 0) at both ends
    assert [len(q) for q in bond_qn] == [1] + [s.shape[-1] for s in sites]
    assert not bond_qn[0].any() and not bond_qn[-1].any()


def test_qc_mpo_bond_dimensions_are_minimal():
    """The numeric rank factorisation reaches the bond dimension of the complementary-operator
    construction, 2 (n/2)^2 + 3 (n/2) + 2 in the middle of n spin orbitals."""
    rng = np.random.default_rng(1)
    h1e, h2e = models.random_qc_integrals(4, rng)
    sites, _ = models.qc_mpo(h1e, h2e)
    n = 8
    assert max(s.shape[-1] for s in sites) == 2 * (n // 2) ** 2 + 3 * (n // 2) + 2


def test_qc_mpo_equals_reference_symbolic_mpo(golden):
    """H6 / STO-3G integrals of the reference's test (mps/tests/test_gs.py:103-145): the MPO built
    numerically here and the reference's symbolic MPO (golden) are the same operator."""
    g = golden("qc_h6")
    ours, _ = models.qc_mpo(g["h1e"], g["h2e"])
    ref = load_mpo(g)
    assert len(ours) == len(ref) == 12
    a, b = models.mpo_to_dense(ours), models.mpo_to_dense(ref)
    assert np.abs(a - b).max() < 1e-12 * max(1.0, np.abs(b).max())
    assert max(s.shape[-1] for s in ours) <= max(s.shape[-1] for s in ref)


def test_exciton_phonon_mpo_reduces_to_the_holstein_chain():
    nmol, d = 3, 3
    jm = np.zeros((nmol, nmol))
    for i in range(nmol - 1):
        jm[i, i + 1] = jm[i + 1, i] = -0.1
    a = models.holstein_mpo(nmol, d, e0=0.3, j=-0.1, omega=0.2, g=1.0)
    b, _ = models.exciton_phonon_mpo([0.3] * nmol, jm, [0.2], [1.0], d)
    assert np.abs(models.mpo_to_dense(a) - models.mpo_to_dense(b)).max() < 1e-13
    # long-range couplings: Hermitian, conserves the exciton number
    rng = np.random.default_rng(2)
    jm = rng.standard_normal((nmol, nmol))
    jm = jm + jm.T
    w, _ = models.exciton_phonon_mpo(np.arange(nmol), jm, [0.2, 0.5], [1.0, 0.3], 2)
    h = models.mpo_to_dense(w)
    assert np.abs(h - h.T).max() < 1e-13


def _check_state(sites, qn, sigmaqn, qntot):
    from oracle import sweep as osw
    m = osw.Mps(sites, qn, sigmaqn, np.asarray(qntot), len(sites) - 1, False)
    for i, s in enumerate(sites):                                     # every non-zero entry obeys the labels
        left, right = np.asarray(qn[i]), np.asarray(qn[i + 1])
        sq = np.asarray(sigmaqn[i])
        nz = np.argwhere(np.abs(s) > 0)
        for idx in nz[:: max(1, len(nz) // 200)]:
            l, r = idx[0], idx[-1]
            phys = tuple(idx[1:-1])
            acc = left[l] + sq[phys]
            if i < len(sites) - 1:
                assert np.array_equal(acc, right[r])
            else:
                assert np.array_equal(acc, np.asarray(qntot))
    return m


def test_random_and_seeded_states_respect_quantum_numbers():
    rng = np.random.default_rng(3)
    sq = models.holstein_sigmaqn(4, 3)
    sites, qn = models.random_mps_qn(sq, [1], 9, rng)
    m = _check_state(sites, qn, sq, [1])
    assert m.check_left_canonical() and abs(m.mp_norm - 1) < 1e-12
    assert max(m.bond_dims) <= 9
    sq = models.qc_sigmaqn(8)
    sites, qn = models.seeded_mps_qn(sq, [2, 2], 12, rng, [1] * 4 + [0] * 4)
    m = _check_state(sites, qn, sq, [2, 2])
    assert max(m.bond_dims) <= 12
    # dominated by the Hartree-Fock determinant
    hf = np.ones(1)
    for i, s in enumerate(sites):
        hf = hf @ s[:, 1 if i < 4 else 0, :]
    assert abs(abs(hf[0]) - 1) < 1e-3 and abs(m.mp_norm - 1) < 1e-3
    # density operator: the ancilla index carries no quantum number
    sq1 = models.exciton_phonon_sigmaqn(2, 1, 3)
    sites, qn, pair = models.random_mpdm_qn(sq1, [1], 6, rng)
    assert all(s.ndim == 4 and s.shape[1] == s.shape[2] for s in sites)
    _check_state(sites, qn, pair, [1])


def test_bench_sampling_helpers():
    import bench
    args = argparse.Namespace(modes=4, mols=3, levels=3, orbitals=3, fmo_modes=1, dt=0.05)
    work = bench.make_workload("sbm_tdvp", 8, args, seed=1)
    assert bench.sample_sites(work, 99) == list(range(work["nsite"] - 1, -1, -1))      # TDVP: right to left
    assert bench.sample_sites(work, 2) == [3, 1]
    work = bench.make_workload("holstein_dmrg", 8, args, seed=1)
    assert bench.sample_sites(work, 99) == list(range(work["nsite"] - 1))
    dims = [s.shape[0] for s in work["sites"]] + [1]
    picks = bench.flop_quantile_sites(work, dims, 2)
    assert len(picks) == 2 and all(0 <= p < work["nsite"] - 1 for p in picks)
    # the picks sit where the bond dimension is largest, not at the cheap ends
    assert all(min(dims[p], dims[p + 2]) >= 3 for p in picks)
    for name in ("qc_dmrg", "fmo_thermal"):
        w = bench.make_workload(name, 6, args, seed=1)
        assert w["nsite"] == len(w["sites"]) == len(w["mpo"])

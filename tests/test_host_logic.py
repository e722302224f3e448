"""CPU tests: hand-built MPOs, host-side bookkeeping, and the C-ABI symbol table."""
import os
import re

import numpy as np
import pytest

from renormalizer_b200 import models


def kron_all(ops):
    out = np.eye(1)
    for o in ops:
        out = np.kron(out, o)
    return out


def test_spin_boson_mpo_dense():
    omegas, g, d = [0.7, 1.3, 2.1], [0.2, -0.4, 0.1], 3
    sites = models.spin_boson_mpo(0.3, 1.1, omegas, g, d)
    idn, num, x = models._boson_ops(d)
    sz, sx, i2 = np.diag([1.0, -1.0]), np.array([[0, 1.0], [1.0, 0]]), np.eye(2)
    h = kron_all([0.3 * sz + 1.1 * sx] + [idn] * 3)
    for k in range(3):
        ops = [i2] + [idn] * 3
        ops[k + 1] = omegas[k] * num
        h += kron_all(ops)
        ops = [sz] + [idn] * 3
        ops[k + 1] = g[k] * x
        h += kron_all(ops)
    assert np.abs(models.mpo_to_dense(sites) - h).max() < 1e-13
    assert [s.shape[0] for s in sites] == [1, 3, 3, 3]


def test_holstein_mpo_dense():
    nmol, d = 3, 3
    e0, j, omega, g = 0.5, -0.1, 0.2, 0.9
    sites = models.holstein_mpo(nmol, d, e0=e0, j=j, omega=omega, g=g)
    idn, num_b, x = models._boson_ops(d)
    adag = np.array([[0.0, 0.0], [1.0, 0.0]])
    a, i2 = adag.T, np.eye(2)
    n_e = adag @ a
    base = []
    for _ in range(nmol):
        base += [i2, idn]

    def term(repl):
        ops = list(base)
        for k, o in repl.items():
            ops[k] = o
        return kron_all(ops)
    h = np.zeros((2 ** nmol * d ** nmol,) * 2)
    for m in range(nmol):
        h += term({2 * m: e0 * n_e}) + term({2 * m + 1: omega * num_b})
        h += term({2 * m: g * omega * n_e, 2 * m + 1: x})
        if m < nmol - 1:
            h += term({2 * m: j * adag, 2 * m + 2: a}) + term({2 * m: j * a, 2 * m + 2: adag})
    dense = models.mpo_to_dense(sites)
    assert np.abs(dense - h).max() < 1e-13
    assert np.abs(dense - dense.T).max() < 1e-13


def test_random_mps_qn_is_canonical_and_in_sector():
    rng = np.random.default_rng(0)
    sq = models.holstein_sigmaqn(3, 3)
    sites, qn = models.random_mps_qn(sq, [1], 6, rng)
    for s in sites[:-1]:
        m = s.reshape(-1, s.shape[-1])
        assert np.abs(m.T @ m - np.eye(m.shape[1])).max() < 1e-12
    # every nonzero element respects qn_left + sigma = qn_right (left part)
    for i, s in enumerate(sites[:-1]):
        big = qn[i][:, None, None, :] + sq[i][None, :, None, :] - qn[i + 1][None, None, :, :]
        assert np.abs(s[np.any(big != 0, axis=-1)]).max() == 0
    psi = sites[0]
    for s in sites[1:]:
        psi = np.tensordot(psi, s, axes=(-1, 0))
    assert abs(np.linalg.norm(psi) - 1) < 1e-12


def test_random_mps_sites_canonical():
    rng = np.random.default_rng(1)
    sites = models.random_mps_sites([2, 4, 4, 4], 5, rng, dtype=np.complex128)
    assert [s.shape[0] for s in sites] == [1, 2, 5, 4]
    for s in sites[:-1]:
        m = s.reshape(-1, s.shape[-1])
        assert np.abs(m.conj().T @ m - np.eye(m.shape[1])).max() < 1e-12


def test_add_outer_and_mask_match_oracle():
    from renormalizer_b200.svd_qn import add_outer, get_qn_mask
    from oracle import svdqn
    rng = np.random.default_rng(2)
    a, b = rng.integers(0, 3, (4, 2)), rng.integers(0, 3, (3, 5, 2))
    assert np.array_equal(add_outer(a, b), svdqn.add_outer(a, b))
    assert np.array_equal(get_qn_mask(add_outer(a, b), [2, 1]), svdqn.get_qn_mask(svdqn.add_outer(a, b), [2, 1]))


def test_compress_config_m_trunc():
    from renormalizer_b200.configs import CompressConfig, CompressCriteria
    sigma = np.array([1.0, 0.5, 1e-2, 1e-5])
    c = CompressConfig(CompressCriteria.threshold, threshold=1e-3)
    assert c.compute_m_trunc(sigma, 0, True) == 3
    c = CompressConfig(CompressCriteria.fixed, max_bonddim=2)
    c.set_bonddim(5)
    assert c.compute_m_trunc(sigma, 0, True) == 2
    c = CompressConfig(CompressCriteria.both, threshold=1e-1, max_bonddim=3)
    c.set_bonddim(5)
    assert c.compute_m_trunc(sigma, 1, False) == 2
    with pytest.raises(ValueError):
        CompressConfig(threshold=1.0)


def test_c_abi_exports_every_declared_symbol():
    """include/rn_b200.h is the contract: every function it declares must be exported by the
    built library and bound by the loader (no compute call is made here)."""
    from renormalizer_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "rn_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(rn_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.rn_version().startswith(b"rn_b200")


def test_compute_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from renormalizer_b200 import _lib
    with pytest.raises(_lib.RnError):
        _lib.get()


def _dense_from_mpo(sites):
    """Contract MPO site tensors W[b, up, down, f] into the dense operator (rows = up indices)."""
    cur = sites[0][0]                                   # (up, down, f)
    cur = cur.transpose(2, 0, 1)[None]                  # dummy: (1, f, up, down)
    op = sites[0][0].transpose(2, 0, 1)                 # (f, U, D)
    for w in sites[1:]:
        op = np.einsum("fUD,fudg->gUuDd", op, w)
        g, U, u, D, d = op.shape
        op = op.reshape(g, U * u, D * d)
    return op[0]


def test_mpo_algebra_for_omega_targeting():
    """Mpo.add / scale / identity_like / squared (gs.py:106-111 needs (H - omega)^2 as an MPO):
    checked against dense operator algebra."""
    from renormalizer_b200.mpo import Mpo
    rng = np.random.default_rng(5)
    pd = [2, 3, 2, 2]
    bonds = [1, 3, 4, 2, 1]
    sites = [rng.standard_normal((bonds[i], pd[i], pd[i], bonds[i + 1])) for i in range(4)]
    mpo = Mpo(sites)
    H = _dense_from_mpo(mpo.to_numpy())
    n = H.shape[0]
    ident = Mpo.identity_like(mpo)
    assert np.allclose(_dense_from_mpo(ident.to_numpy()), np.eye(n))
    shifted = mpo.add(ident.scale(-0.37))
    Hs = _dense_from_mpo(shifted.to_numpy())
    assert np.allclose(Hs, H - 0.37 * np.eye(n), atol=1e-13)
    sq = shifted.squared()
    assert sq.bond_dims == [b * b for b in shifted.bond_dims]
    assert np.allclose(_dense_from_mpo(sq.to_numpy()), Hs @ Hs, atol=1e-11)


def test_sweep_host_logic_with_stubbed_kernels():
    """The host side of the sweeps (bookkeeping, truncation, drivers) against the reference's goldens
    with the C-ABI kernels replaced by the NumPy/torch test double `tests/_host_logic_stub.py`, in a
    subprocess so that the patched modules do not leak into this session."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, PYTHONPATH=here + os.pathsep + os.environ.get("PYTHONPATH", ""))
    select = ("expand or adaptive_golden or entropy or prop_and_compress or quickstart or thermal or exciton "
              "or density or variational or stacked or dmrg_qc_h6_golden or dmrg_holstein_golden")
    cmd = [sys.executable, "-m", "pytest", "-p", "_host_logic_stub", os.path.join(here, "test_gpu_sweeps.py"),
           "-q", "-x", "-p", "no:cacheprovider", "-k", f"({select}) and not tensor_path"]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=1500, cwd=os.path.dirname(here))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert " passed" in out.stdout and "failed" not in out.stdout

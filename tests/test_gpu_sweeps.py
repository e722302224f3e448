"""GPU parity tests of the sweep drivers (svd_qn, Krylov, Davidson, DMRG, TDVP-PS) against the
reference's golden vectors and the CPU oracle.  Tolerances follow BASELINE.json's north_star:
energies / observables to 1e-10, site tensors (gauge-invariant form) to 1e-8."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from helpers import load_mpo, load_oracle_mps, relerr
from renormalizer_b200.configs import EvolveConfig, EvolveMethod

pytestmark = pytest.mark.gpu

E_TOL = 1e-10
T_TOL = 1e-8


def dev(a):
    from renormalizer_b200.backend import asxp
    return asxp(a)


def host(t):
    return t.detach().cpu().numpy()


def to_device_mps(om):
    from renormalizer_b200.configs import EvolveConfig, EvolveMethod
    from renormalizer_b200.mps import Mps
    m = Mps(om.sites, om.qn, om.sigmaqn, om.qntot, om.qnidx, om.to_right)
    m.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps)     # the sweep integrator these tests exercise
    return m


@pytest.mark.parametrize("t", ["r", "c"])
@pytest.mark.parametrize("system", ["L", "R"])
def test_svd_qn_golden(golden, t, system):
    from renormalizer_b200.svd_qn import svd_qn
    g = golden("svdqn")
    k = f"{t}_{system}"
    c, ql, qr, qntot = g[k + "_c"], g[k + "_qnbigl"], g[k + "_qnbigr"], g["qntot"]
    mat = c.reshape(int(np.prod(ql.shape[:-1])), -1)
    # economic SVD: singular values, quantum numbers and the decomposition itself
    u, su, qnl, v, sv, qnr = svd_qn(dev(c), ql, qr, qntot, system=system, full_matrices=False)
    u, v = host(u), host(v)
    assert np.abs(su - g[k + "_svd_s"]).max() < 1e-13
    assert np.array_equal(np.array(qnl), g[k + "_svd_qnl"])
    assert np.array_equal(np.array(qnr), g[k + "_svd_qnr"])
    assert relerr((u * su) @ v.T, mat) < 1e-12
    nz = su > 1e-12
    ref_u = g[k + "_svd_u"]
    # gauge invariant comparison of the singular subspaces (site tensors to 1e-8)
    assert np.abs(np.abs(np.sum(u[:, nz].conj() * ref_u[:, nz], axis=0)) - 1).max() < T_TOL
    # full matrices: orthonormal completion, same sizes and quantum numbers as the reference
    np.random.seed(11)
    u, su, qnl, v, sv, qnr = svd_qn(dev(c), ql, qr, qntot, system=system, full_matrices=True)
    u, v = host(u), host(v)
    assert u.shape == g[k + "_fsvd_u"].shape and v.shape == g[k + "_fsvd_v"].shape
    assert np.abs(u.conj().T @ u - np.eye(u.shape[1])).max() < 1e-12
    assert np.abs(v.conj().T @ v - np.eye(v.shape[1])).max() < 1e-12
    assert sorted(map(tuple, qnl)) == sorted(map(tuple, g[k + "_fsvd_qnl"]))
    assert np.abs(np.sort(su) - np.sort(g[k + "_fsvd_su"])).max() < 1e-13
    # QR / LQ: orthonormal factor spans the same space, product reproduces the tensor
    u, qnl, v, qnr = svd_qn(dev(c), ql, qr, qntot, QR=True, system=system, full_matrices=False)
    u, v = host(u), host(v)
    assert relerr(u @ v.T, mat) < 1e-12
    assert np.array_equal(np.array(qnl), g[k + "_qr_qnl"])
    assert np.array_equal(np.array(qnr), g[k + "_qr_qnr"])
    q = u if system == "L" else v
    qref = g[k + "_qr_u"] if system == "L" else g[k + "_qr_v"]
    assert np.abs(q.conj().T @ q - np.eye(q.shape[1])).max() < 1e-12
    assert np.abs(q @ (q.conj().T @ qref) - qref).max() < T_TOL


def test_svd_qn_invalid_qn_raises():
    from renormalizer_b200.svd_qn import svd_qn, add_outer
    qnl = np.zeros((2, 1), dtype=int)
    with pytest.raises(ValueError):
        svd_qn(dev(np.ones((2, 2, 2))), add_outer(qnl, qnl), qnl, np.array([5]), system="L")


def test_expm_krylov_golden(golden):
    from renormalizer_b200.krylov import expm_krylov
    from renormalizer_b200 import ops
    g = golden("krylov")
    h, v = dev(g["h"]), dev(g["v"])
    for i in range(3):
        res, j = expm_krylov(lambda y: ops.matmul(h, y.reshape(-1, 1)).reshape(-1), complex(g[f"dt{i}"]), v)
        assert j == int(g[f"j{i}"])
        assert relerr(host(res), g[f"res{i}"]) < 1e-11


def test_expm_krylov_small_spaces():
    """Krylov space exhausted (n <= number of steps) and breakdown on an eigenvector."""
    from renormalizer_b200.krylov import expm_krylov
    from renormalizer_b200 import ops
    import scipy.linalg
    rng = np.random.default_rng(1)
    for n in (1, 2, 3, 6):
        h = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        h = h + h.conj().T
        v = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        hd = dev(h)
        res, j = expm_krylov(lambda y: ops.matmul(hd, y.reshape(-1, 1)).reshape(-1), -0.1j, dev(v))
        assert relerr(host(res), scipy.linalg.expm(-0.1j * h) @ v) < 1e-10
    n = 40
    h = np.diag(np.arange(n, dtype=float)).astype(complex)
    v = np.zeros(n, dtype=complex)
    v[3] = 2.0
    hd = dev(h)
    res, j = expm_krylov(lambda y: ops.matmul(hd, y.reshape(-1, 1)).reshape(-1), -0.5j, dev(v))
    assert relerr(host(res), np.exp(-0.5j * 3) * v) < 1e-12


@pytest.mark.parametrize("nroots", [1, 3])
def test_davidson_golden(golden, nroots):
    from renormalizer_b200.davidson import davidson
    from renormalizer_b200 import ops
    g = golden("davidson")
    a = dev(g["a"])
    hd = dev(np.diag(g["a"]).copy())
    count = [0]

    def aop(x):
        count[0] += 1
        return ops.matmul(a, x.reshape(-1, 1)).reshape(-1)
    e, c = davidson(aop, [dev(x) for x in g[f"x0_{nroots}"]],
                    lambda x, e, *args: x / (hd - e + 1e-4), max_cycle=100, nroots=nroots)
    assert np.abs(np.atleast_1d(e) - np.atleast_1d(g[f"e_{nroots}"])).max() < E_TOL
    cs = [c] if nroots == 1 else c
    ref = np.atleast_2d(g[f"c_{nroots}"])
    for ci, ri in zip(cs, ref):
        assert abs(abs(np.vdot(host(ci), ri)) - 1) < T_TOL
    assert abs(count[0] - int(g[f"nhop_{nroots}"])) <= 2


@pytest.mark.parametrize("method", ["1site", "2site"])
def test_dmrg_holstein_golden(golden, method):
    """optimize_mps on the reference's Holstein test model, same MPO and initial MPS."""
    from renormalizer_b200.gs import optimize_mps
    from renormalizer_b200.mpo import Mpo
    g = golden("holstein")
    mpo = Mpo(load_mpo(g))
    mps = to_device_mps(load_oracle_mps(g, "mps0"))
    mps.optimize_config.procedure = [[int(a), float(b)] for a, b in g["procedure"]]
    mps.optimize_config.method = method
    np.random.seed(99)
    energies, opt = optimize_mps(mps, mpo)
    ref = g[f"{method}_energies"]
    assert len(energies) == len(ref)
    # converged sweeps agree to the energy tolerance; early sweeps (random null-space completion
    # differs from LAPACK's) to the reference's own convergence criterion
    assert abs(energies[-1] - ref[-1]) < 1e-9
    assert np.abs(np.array(energies) - ref).max() < 1e-6
    assert abs(opt.expectation(mpo) - float(g[f"{method}_expectation"])) < 1e-9
    assert energies[-1] == pytest.approx(0.08401412 + float(g["gs_zpe"]), rel=1e-5)
    # the reference's final ensure_left_canonical().canonicalise() leaves a right-canonical MPS
    assert opt.check_right_canonical() and opt.qnidx == 0


def test_dmrg_first_site_energy_matches_oracle(golden):
    """One Davidson solve on identical inputs: energy to 1e-10, centre tensor to 1e-8."""
    from renormalizer_b200.gs import eigh_iterative
    from renormalizer_b200.lib import Environ
    from renormalizer_b200.mpo import Mpo
    from renormalizer_b200.svd_qn import get_qn_mask
    from oracle import sweep as osw, contract as oc
    from oracle.davidson import davidson as odav
    g = golden("holstein")
    mpo_np = load_mpo(g)
    om = load_oracle_mps(g, "mps0")
    om.ensure_right_canonical()
    # oracle: first two-site problem of a right-moving sweep
    env = osw.Environ(om, mpo_np, "R")
    cidx = [0, 1]
    ltensor, rtensor = env.sentinel, env.read("R", 2)
    qnbigl, qnbigr, qnmat = om.big_qn(cidx)
    mask = osw.get_qn_mask(qnmat, om.qntot)
    guess = np.tensordot(om.sites[0], om.sites[1], axes=1)
    cmo = [mpo_np[0], mpo_np[1]]
    hdiag = oc.hop_diag(ltensor, rtensor, cmo)[mask]

    def hop(x):
        full = np.zeros(mask.shape)
        full[mask] = x
        return oc.hop_apply(ltensor, rtensor, cmo, full)[mask]
    if mask.sum() < 4:
        pytest.skip("degenerate sector")
    e_ref, c_ref = odav(hop, [guess[mask]], lambda x, e, *a: x / (hdiag - e + 1e-4), max_cycle=100)
    dm = to_device_mps(om)
    mpo = Mpo(mpo_np)
    e, c, nhop = eigh_iterative(dm, mask, dev(ltensor), dev(rtensor), [mpo[0], mpo[1]], dev(guess))
    assert abs(e - e_ref) < E_TOL
    c_ref = c_ref / np.sign(c_ref[np.abs(c_ref).argmax()])
    assert np.abs(host(c)[mask] - c_ref).max() < T_TOL


def test_tdvp_ps_golden(golden):
    """Mps.evolve (TDVP-PS, Krylov) on the reference's spin-boson run."""
    from renormalizer_b200.mpo import Mpo
    g = golden("sbm")
    mpo = Mpo(load_mpo(g))
    sz = Mpo(load_mpo(g, "sigma_z"))
    mps = to_device_mps(load_oracle_mps(g, "mps0"))
    dt = float(g["dt"])
    szs, es = [mps.expectation(sz)], [mps.expectation(mpo)]
    for i in range(int(g["nsteps"])):
        mps = mps.evolve(mpo, dt)
        szs.append(mps.expectation(sz))
        es.append(mps.expectation(mpo))
        if i == 0:
            ref1 = to_device_mps(load_oracle_mps(g, "mps1"))
            assert abs(abs(ref1.conj().dot(mps)) - 1) < T_TOL
    assert np.abs(np.array(szs) - g["sigma_z_t"]).max() < E_TOL
    assert np.abs(np.array(es) - g["energy_t"]).max() < E_TOL
    refT = to_device_mps(load_oracle_mps(g, "mpsT"))
    assert abs(abs(refT.conj().dot(mps)) - 1) < T_TOL
    assert abs(mps.mp_norm - 1) < 1e-12


def test_tdvp_ps_vs_oracle_midsize():
    """Random full-rank complex MPS, M=24: one step against the oracle, site tensors compared
    through the overlap and observables."""
    from renormalizer_b200 import models
    from renormalizer_b200.mpo import Mpo
    from renormalizer_b200.mps import Mps
    from oracle import sweep as osw
    rng = np.random.default_rng(8)
    nmodes, d, M = 7, 4, 24
    omega, gcoup = models.ohmic_modes(nmodes, alpha=0.3, omega_c=5.0)
    w = models.spin_boson_mpo(0.2, 1.0, omega, gcoup, d)
    sites = models.random_mps_sites([2] + [d] * nmodes, M, rng, dtype=np.complex128)
    n = len(sites)
    qn = [np.zeros((s.shape[0], 1), dtype=int) for s in sites] + [np.zeros((1, 1), dtype=int)]
    sq = [np.zeros((s.shape[1], 1), dtype=int) for s in sites]
    om = osw.Mps(sites, qn, sq, [0], n - 1, False)
    dm = Mps(sites, qn, sq, [0], n - 1, False)
    dm.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps)
    o1 = osw.evolve_tdvp_ps(om, w, 0.05)
    d1 = dm.evolve(Mpo(w), 0.05)
    ref = Mps(o1.sites, o1.qn, o1.sigmaqn, o1.qntot, o1.qnidx, o1.to_right)
    assert abs(abs(ref.conj().dot(d1)) - 1) < T_TOL
    szm = models.spin_boson_sigma_z_mpo(nmodes, d)
    assert abs(d1.expectation(Mpo(szm)) - o1.expectation(szm)) < E_TOL
    assert abs(d1.expectation(Mpo(w)) - o1.expectation(w)) < E_TOL


def test_tdvp_roundtrip_time_reversal():
    """Size-independent property: evolving by +dt then -dt returns the initial state (TDVP-PS is
    symmetric), checked at a bond dimension the oracle would not finish quickly."""
    from renormalizer_b200 import models
    from renormalizer_b200.mpo import Mpo
    from renormalizer_b200.mps import Mps
    rng = np.random.default_rng(9)
    nmodes, d, M = 9, 6, 64
    omega, gcoup = models.ohmic_modes(nmodes, alpha=0.2, omega_c=5.0)
    mpo = Mpo(models.spin_boson_mpo(0.0, 1.0, omega, gcoup, d))
    sites = models.random_mps_sites([2] + [d] * nmodes, M, rng, dtype=np.complex128)
    m0 = Mps.without_qn(sites)
    m0.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps)
    m1 = m0.evolve(mpo, 0.02)
    m2 = m1.evolve(mpo, -0.02)
    assert abs(abs(m0.conj().dot(m2)) - 1) < 1e-9
    e0, e1 = m0.expectation(mpo), m1.expectation(mpo)
    assert abs(e0 - e1) < 1e-9     # energy conservation of TDVP


@pytest.fixture
def tensor_path():
    """Route EVERY contraction GEMM (however small) through the tcgen05 split path."""
    from renormalizer_b200 import _lib
    from renormalizer_b200.backend import backend
    lib = _lib.get()
    lib.rn_set_ozaki(7, 0.0)
    backend.gemm_path = 1
    yield
    backend.gemm_path = 1
    lib.rn_set_ozaki(7, 4.0e6)


def test_tdvp_ps_golden_tensor_path(golden, tensor_path):
    """Same golden TDVP-PS run with all GEMMs on tcgen05 (int8 split, 7 digits)."""
    from renormalizer_b200.mpo import Mpo
    g = golden("sbm")
    mpo = Mpo(load_mpo(g))
    sz = Mpo(load_mpo(g, "sigma_z"))
    mps = to_device_mps(load_oracle_mps(g, "mps0"))
    dt = float(g["dt"])
    szs, es = [mps.expectation(sz)], [mps.expectation(mpo)]
    for i in range(int(g["nsteps"])):
        mps = mps.evolve(mpo, dt)
        szs.append(mps.expectation(sz))
        es.append(mps.expectation(mpo))
    assert np.abs(np.array(szs) - g["sigma_z_t"]).max() < E_TOL
    assert np.abs(np.array(es) - g["energy_t"]).max() < E_TOL
    refT = to_device_mps(load_oracle_mps(g, "mpsT"))
    assert abs(abs(refT.conj().dot(mps)) - 1) < T_TOL


def test_dmrg_holstein_golden_tensor_path(golden, tensor_path):
    from renormalizer_b200.gs import optimize_mps
    from renormalizer_b200.mpo import Mpo
    g = golden("holstein")
    mpo = Mpo(load_mpo(g))
    mps = to_device_mps(load_oracle_mps(g, "mps0"))
    mps.optimize_config.procedure = [[int(a), float(b)] for a, b in g["procedure"]]
    mps.optimize_config.method = "2site"
    np.random.seed(99)
    energies, opt = optimize_mps(mps, mpo)
    ref = g["2site_energies"]
    assert abs(energies[-1] - ref[-1]) < 1e-9
    assert abs(opt.expectation(mpo) - float(g["2site_expectation"])) < 1e-9


def test_dmrg_omega_targeting_golden(golden):
    """optimize_mps(..., omega) -- the (H - omega)^2 variational function of gs.py:106-111 -- against
    the reference run of the same start state and the reference's own acceptance value
    (mps/tests/test_gs.py:66-86: 0.08401412 + zero point energy)."""
    from renormalizer_b200.gs import optimize_mps
    from renormalizer_b200.mpo import Mpo
    g = golden("holstein")
    mpo = Mpo(load_mpo(g))
    mps = to_device_mps(load_oracle_mps(g, "mps0"))
    mps.optimize_config.procedure = [[int(a), float(b)] for a, b in g["procedure"]]
    mps.optimize_config.method = "2site"
    mps.optimize_config.e_atol = 1e-6
    mps.optimize_config.e_rtol = 1e-6
    np.random.seed(99)
    energies, opt = optimize_mps(mps, mpo, omega=float(g["omega"]))
    e = opt.expectation(mpo)
    assert np.allclose(e, 0.08401412 + float(g["gs_zpe"]))          # the reference's own test
    assert abs(e - float(g["omega_expectation"])) < 1e-7
    assert abs(energies[-1] - g["omega_energies"][-1]) < 1e-9


@pytest.mark.parametrize("cplx", [False, True])
def test_svd_qn_rank_deficient_block_keeps_orthonormal_vectors(cplx):
    """A bond whose dimension exceeds the rank: the vectors with zero singular value that the sweep
    keeps must still be orthonormal (LAPACK returns them so; plain one-sided Jacobi returns zeros)."""
    from renormalizer_b200.svd_qn import svd_qn, add_outer
    rng = np.random.default_rng(12)
    def rnd(shape):
        a = rng.standard_normal(shape)
        return a + 1j * rng.standard_normal(shape) if cplx else a
    l, d, r, rank = 12, 4, 20, 5
    a = (rnd((l * d, rank)) @ rnd((rank, r))).reshape(l, d, r)
    qnl = np.zeros((l, 1), dtype=int)
    qnr = np.zeros((r, 1), dtype=int)
    sq = np.zeros((d, 1), dtype=int)
    u, su, _, v, sv, _ = svd_qn(dev(a), add_outer(qnl, sq), qnr, np.array([0]), system="L", full_matrices=False)
    u, v = host(u), host(v)
    k = min(l * d, r)
    assert u.shape == (l * d, k) and v.shape == (r, k)
    assert np.abs(u.conj().T @ u - np.eye(k)).max() < 1e-11
    assert np.abs(v.conj().T @ v - np.eye(k)).max() < 1e-11
    assert np.abs((u * su) @ v.T - a.reshape(l * d, r)).max() < 1e-12 * np.abs(a).max() * 100
    assert np.all(su[rank:] < 1e-12 * su[0])


def test_tdvp_ps2_golden(golden):
    """Mps.evolve with the two-site projector-splitting integrator (mps.py:1407-1517) on the
    reference's spin-boson run: observables, energy conservation, bond dimensions, final state."""
    from renormalizer_b200.configs import CompressConfig, CompressCriteria, EvolveConfig, EvolveMethod
    from renormalizer_b200.mpo import Mpo
    g = golden("sbm")
    mpo = Mpo(load_mpo(g))
    sz = Mpo(load_mpo(g, "sigma_z"))
    mps = to_device_mps(load_oracle_mps(g, "mps0"))
    mps.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=12)
    mps.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps2, adaptive=False)
    dt = float(g["dt"])
    szs, es = [mps.expectation(sz)], [mps.expectation(mpo)]
    for i in range(int(g["ps2_nsteps"])):
        mps = mps.evolve(mpo, dt)
        szs.append(mps.expectation(sz))
        es.append(mps.expectation(mpo))
    assert np.abs(np.array(szs) - g["ps2_sigma_z_t"]).max() < E_TOL
    assert np.abs(np.array(es) - g["ps2_energy_t"]).max() < E_TOL
    assert mps.bond_dims == list(g["ps2_bond_dims"])
    refT = to_device_mps(load_oracle_mps(g, "ps2_mpsT"))
    assert abs(abs(refT.conj().dot(mps)) - 1) < T_TOL


@pytest.mark.parametrize("method", ["1site", "2site"])
def test_dmrg_state_averaged_golden(golden, method):
    """State-averaged DMRG (nroots = 3; gs.py nroots > 1, mp.py:780-838, svd_qn.py:243 eigh_qn) against
    the reference run from the same start state and the reference's own acceptance values
    (mps/tests/test_gs.py:66-86)."""
    from renormalizer_b200.gs import optimize_mps
    from renormalizer_b200.mpo import Mpo
    g = golden("holstein")
    mpo = Mpo(load_mpo(g))
    mps = to_device_mps(load_oracle_mps(g, "mps0"))
    mps.optimize_config.procedure = [[int(a), float(b)] for a, b in g["procedure"]]
    mps.optimize_config.method = method
    mps.optimize_config.nroots = 3
    np.random.seed(99)
    energies, opts = optimize_mps(mps, mpo)
    assert len(opts) == 3
    got = np.array([o.expectation(mpo) for o in opts])
    assert np.abs(np.array(energies[-1]) - g[f"sa_{method}_energies"][-1]).max() < 1e-8
    assert np.abs(got - g[f"sa_{method}_expectations"]).max() < 1e-8
    std = np.array([0.08401412, 0.08449771, 0.08449801]) + float(g["gs_zpe"])
    assert np.allclose(got, std)
    # the three states are orthonormal
    for i in range(3):
        for j in range(3):
            ov = abs(opts[i].conj().dot(opts[j]))
            assert abs(ov - (1.0 if i == j else 0.0)) < 1e-6


# ---------------------------------------------------------------- BASELINE.json parity cases
def _device_dmrg(g, mpo, method, **cfg):
    from renormalizer_b200.gs import optimize_mps
    mps = to_device_mps(load_oracle_mps(g, "mps0"))
    mps.optimize_config.procedure = [[int(a), float(b)] for a, b in g["procedure"]]
    mps.optimize_config.method = method
    for k, v in cfg.items():
        setattr(mps.optimize_config, k, v)
    np.random.seed(99)
    return optimize_mps(mps, mpo)


@pytest.mark.parametrize("method", ["1site", "2site"])
def test_dmrg_stacked_mpo_golden(golden, method):
    """optimize_mps(mps, StackedMpo([H, H])) (mpo.py:483-494, mps/tests/test_gs.py:148-158): twice
    the single-MPO energies, and the reference's own stacked trajectory."""
    from renormalizer_b200.mpo import Mpo, StackedMpo
    g = golden("stacked")
    mpo = Mpo(load_mpo(g))
    e2, _ = _device_dmrg(g, StackedMpo([mpo, mpo]), method)
    ref = g[f"{method}_double_energies"]
    assert len(e2) == len(ref)
    assert abs(e2[-1] - ref[-1]) < 1e-9
    assert np.abs(np.array(e2) - ref).max() < 1e-6
    assert np.abs(np.array(e2) - 2 * g[f"{method}_single_energies"]).max() < 1e-6


def test_dmrg_stacked_split_hamiltonian_golden(golden):
    """A Hamiltonian split into its on-site and its coupling part, each member an MPO of its own."""
    from renormalizer_b200.mpo import Mpo, StackedMpo
    g = golden("stacked")
    stacked = StackedMpo([Mpo(load_mpo(g, "mpo_a")), Mpo(load_mpo(g, "mpo_b"))])
    e, opt = _device_dmrg(g, stacked, "2site")
    assert abs(e[-1] - g["split_energies"][-1]) < 1e-9
    assert np.abs(np.array(e) - g["split_energies"]).max() < 1e-6
    assert abs(opt.expectation(Mpo(load_mpo(g))) - float(g["split_expectation"])) < 1e-9


def test_dmrg_stacked_omega_raises(golden):
    from renormalizer_b200.mpo import Mpo, StackedMpo
    from renormalizer_b200.gs import optimize_mps
    g = golden("stacked")
    mpo = Mpo(load_mpo(g))
    mps = to_device_mps(load_oracle_mps(g, "mps0"))
    with pytest.raises(NotImplementedError):
        optimize_mps(mps, StackedMpo([mpo, mpo]), omega=0.1)


def _check_qc(g, mpo, e, opt):
    ref = g["energies"]
    assert len(e) == len(ref)
    # the start state is a Hartree-Fock product state plus 1e-8 noise: the first two sweeps grow
    # the bonds out of null-space completions (arbitrary in LAPACK and here), so they are compared
    # loosely; from the third sweep on the trajectory is the reference's
    e, ref = np.array(e), np.array(ref)
    assert np.abs(e[:2] - ref[:2]).max() < 1e-2
    assert np.abs(e[2:] - ref[2:]).max() < 1e-6
    assert abs(e[-1] - ref[-1]) < 1e-9
    assert abs(opt.expectation(mpo) - float(g["expectation"])) < 1e-8
    assert np.allclose(min(e), float(g["fci_e"]), atol=5e-3)           # test_gs.py:145
    assert max(opt.bond_dims) <= 30


def test_dmrg_qc_h6_golden(golden):
    """BASELINE configs[4] in miniature: ab initio DMRG on the reference's H6 FCIDUMP
    (mps/tests/test_gs.py:103-145) -- two conserved quantum numbers, M = 30, two-site sweeps."""
    from renormalizer_b200.mpo import Mpo
    g = golden("qc_h6")
    mpo = Mpo(load_mpo(g))
    e, opt = _device_dmrg(g, mpo, "2site")
    _check_qc(g, mpo, e, opt)


def test_dmrg_qc_h6_golden_tensor_path(golden, tensor_path):
    """The same run with every contraction on the tcgen05 digit path."""
    from renormalizer_b200.mpo import Mpo
    g = golden("qc_h6")
    mpo = Mpo(load_mpo(g))
    e, opt = _device_dmrg(g, mpo, "2site")
    _check_qc(g, mpo, e, opt)


def _device_mps_with_coeff(g, prefix):
    om = load_oracle_mps(g, prefix, meta=prefix)
    from renormalizer_b200.mps import Mps
    return Mps(om.sites, om.qn, om.sigmaqn, om.qntot, om.qnidx, om.to_right,
               coeff=complex(g[prefix + "_coeff"]))


def _exciton_run(g, mps, nsteps):
    from renormalizer_b200.configs import CompressConfig, CompressCriteria, EvolveConfig, EvolveMethod
    from renormalizer_b200.mpo import Mpo
    mpo = Mpo(load_mpo(g))
    occ = [Mpo(load_mpo(g, f"occ{i}")) for i in range(int(g["nmol"]))]
    mps.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=10)
    mps.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps, adaptive=False)
    occs, es = [[mps.expectation(o) for o in occ]], [mps.expectation(mpo)]
    for _ in range(nsteps):
        mps = mps.evolve(mpo, float(g["dt"]))
        occs.append([mps.expectation(o) for o in occ])
        es.append(mps.expectation(mpo))
    return mps, np.array(occs), np.array(es)


def test_tdvp_ps_exciton_golden(golden):
    """BASELINE configs[3] in miniature (example/fmo.py): long-range J matrix, one conserved
    exciton -- the quantum-number-blocked TDVP-PS case."""
    g = golden("exciton")
    mps, occs, es = _exciton_run(g, _device_mps_with_coeff(g, "mps0"), int(g["nsteps"]))
    assert np.abs(occs - g["occ_t"]).max() < E_TOL
    assert np.abs(es - g["energy_t"]).max() < E_TOL
    assert abs(occs[-1].sum() - 1) < E_TOL
    assert mps.bond_dims == list(g["bond_dims"])
    refT = to_device_mps(load_oracle_mps(g, "mpsT"))
    assert abs(abs(refT.conj().dot(mps)) - 1) < T_TOL


def test_tdvp_ps_density_operator_golden(golden):
    """The same evolution for a density operator (MpDm: every site carries an ancilla index,
    hop_expr.py:83-117, lib.py:213-262) started from the maximally entangled one-exciton state."""
    g = golden("exciton")
    dm0 = _device_mps_with_coeff(g, "dm0")
    assert dm0[0].ndim == 4
    dm, occs, es = _exciton_run(g, dm0, int(g["dm_nsteps"]))
    assert np.abs(occs - g["dm_occ_t"]).max() < E_TOL
    assert np.abs(es - g["dm_energy_t"]).max() < E_TOL
    ref = to_device_mps(load_oracle_mps(g, "dmT", meta="dm0"))
    assert abs(abs(ref.conj().dot(dm)) - 1) < T_TOL


def test_two_spin_quickstart_golden(golden):
    """BASELINE configs[0], the README quickstart (README.md:36-58) through Mps.evolve: the
    two-site sweep integrator reproduces the reference's tdvp_ps2 run, the values the README's
    own (propagate-and-compress) loop prints to that integrator's error, and -cos(2t)."""
    from renormalizer_b200.configs import EvolveConfig, EvolveMethod
    from renormalizer_b200.mpo import Mpo
    g = golden("two_spin")
    mpo, z = Mpo(load_mpo(g)), Mpo(load_mpo(g, "z"))
    mps = to_device_mps(load_oracle_mps(g, "mps0"))
    mps.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps2)
    zs = []
    for _ in range(10):
        mps = mps.evolve(mpo, 0.05)
        zs.append(mps.expectation(z))
    assert np.abs(np.array(zs) - g["ps2_z_t"]).max() < E_TOL
    assert np.abs(np.array(zs) - g["pc_z_t"]).max() < 1e-6
    assert np.abs(np.array(zs) + np.cos(2 * 0.05 * np.arange(1, 11))).max() < 1e-6
    assert mps.bond_dims == list(g["ps2_bond_dims"])


def test_thermal_imaginary_then_real_time_golden(golden):
    """Finite temperature (BASELINE configs[3]): imaginary-time TDVP-PS of a density operator
    (mps/thermalprop.py:96-98: MpDm.evolve(h_mpo, -i dbeta/2); real Krylov exponent, "mps_and_coeff"
    normalisation), then real-time steps of the thermal state.  Imaginary time amplifies rounding
    differences by exp(dbeta * spectral width) per step (the CPU restatement itself agrees with the
    reference to 1.4e-10 after four steps), hence 1e-9 on this leg."""
    from renormalizer_b200.configs import CompressConfig, CompressCriteria, EvolveConfig, EvolveMethod
    from renormalizer_b200.mpo import Mpo
    g = golden("thermal")
    mpo = Mpo(load_mpo(g))
    occ = [Mpo(load_mpo(g, f"occ{i}")) for i in range(int(g["nmol"]))]
    dm = _device_mps_with_coeff(g, "dm0")
    dm.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=8)
    dm.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps, adaptive=False)
    dbeta = float(g["beta"]) / int(g["nbeta"])
    occs, es = [[dm.expectation(o) for o in occ]], [dm.expectation(mpo)]
    for _ in range(int(g["nbeta"])):
        dm = dm.evolve(mpo, -0.5j * dbeta)
        assert not dm.is_complex                       # real MPDM stays real in imaginary time
        assert abs(dm.mp_norm - 1) < 1e-12 and dm.coeff == 1
        occs.append([dm.expectation(o) for o in occ])
        es.append(dm.expectation(mpo))
    assert np.abs(np.array(occs) - g["imag_occ"]).max() < 1e-9
    assert np.abs(np.array(es) - g["imag_energy"]).max() < 1e-9
    assert np.abs(np.array(occs[1]) - g["imag_occ"][1]).max() < E_TOL
    ref = to_device_mps(load_oracle_mps(g, "dm_beta", meta="dm0"))
    assert abs(abs(ref.conj().dot(dm)) - 1) < T_TOL
    rocc, ren = [], []
    for _ in range(2):
        dm = dm.evolve(mpo, 2.0)
        rocc.append([dm.expectation(o) for o in occ])
        ren.append(dm.expectation(mpo))
    assert np.abs(np.array(rocc) - g["real_occ"]).max() < 1e-9
    assert np.abs(np.array(ren) - g["real_energy"]).max() < 1e-9


def test_adaptive_tdvp_ps_golden(golden):
    """adaptive_tdvp (mps.py:46-115) through Mps.evolve: the controller accepts the same sub-steps
    and ends on the same guess_dt; observables agree to the local solver's own stopping tolerance
    once the step has doubled to dt = 4 (see the oracle test), to 1e-10 on the first call."""
    from renormalizer_b200.configs import CompressConfig, CompressCriteria, EvolveConfig, EvolveMethod
    from renormalizer_b200.mpo import Mpo
    g = golden("thermal")
    mpo = Mpo(load_mpo(g))
    occ = [Mpo(load_mpo(g, f"occ{i}")) for i in range(int(g["nmol"]))]
    mps = _device_mps_with_coeff(g, "mps0")
    mps.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=10)
    mps.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps, adaptive=True, guess_dt=1.0,
                                     adaptive_rtol=5e-4)
    occs, guesses = [], []
    for _ in range(3):
        mps = mps.evolve(mpo, 4.0)
        occs.append([mps.expectation(o) for o in occ])
        guesses.append(mps.evolve_config.guess_dt)
    assert np.allclose(guesses, g["adaptive_guess_dt"], rtol=1e-6)
    assert np.abs(np.array(occs[0]) - g["adaptive_occ"][0]).max() < E_TOL
    assert np.abs(np.array(occs) - g["adaptive_occ"]).max() < 1e-7
    refT = to_device_mps(load_oracle_mps(g, "adaptive_mpsT", meta="mps0"))
    assert abs(abs(refT.conj().dot(mps)) - 1) < 1e-6


def _device_mpo_with_qn(g, prefix="mpo"):
    from renormalizer_b200.mpo import Mpo
    n = int(g[prefix + "_n"])
    return Mpo(load_mpo(g, prefix), qn=[g[f"{prefix}_qn_{i}"] for i in range(n + 1)],
               qntot=g[prefix + "_qntot"], qnidx=int(g[prefix + "_qnidx"]))


@pytest.mark.parametrize("tag", ["thr", "fix"])
def test_prop_and_compress_exciton_golden(golden, tag):
    """Mps.evolve with its default configuration (propagate-and-compress, mps.py:796-884) on the
    exciton model with a conserved exciton: Mpo.apply, Mps.add, canonicalise and the SVD compression
    on the device -- occupations, energy and the bond dimensions the truncation chooses."""
    from renormalizer_b200.configs import CompressConfig, CompressCriteria
    from renormalizer_b200.mpo import Mpo
    g = golden("pc")
    mpo = _device_mpo_with_qn(g)
    occ = [Mpo(load_mpo(g, f"occ{i}")) for i in range(int(g["nmol"]))]
    mps = _device_mps_with_coeff(g, "mps0")
    assert mps.evolve_config.method is EvolveMethod.prop_and_compress          # the reference's default
    mps.compress_config = CompressConfig(CompressCriteria.threshold, threshold=1e-5) if tag == "thr" \
        else CompressConfig(CompressCriteria.fixed, max_bonddim=12)
    occs, es, dims = [], [], []
    for _ in range(4):
        mps = mps.evolve(mpo, 1.0)
        occs.append([mps.expectation(o) for o in occ])
        es.append(mps.expectation(mpo))
        dims.append(mps.bond_dims)
    assert np.array_equal(np.array(dims), g[f"{tag}_bond_dims"])
    assert np.abs(np.array(occs) - g[f"{tag}_occ"]).max() < E_TOL
    assert np.abs(np.array(es) - g[f"{tag}_energy"]).max() < E_TOL
    ref = to_device_mps(load_oracle_mps(g, f"{tag}_mpsT", meta="mps0"))
    assert abs(abs(ref.conj().dot(mps)) - 1) < T_TOL


def test_two_spin_quickstart_default_integrator_golden(golden):
    """The README quickstart exactly as printed (README.md:36-58): default Mps.evolve, ten steps of
    0.05, <Z_0> after each -- the values the reference prints."""
    from renormalizer_b200.mpo import Mpo
    from renormalizer_b200.mps import Mps
    g = golden("two_spin")
    mpo, z = _device_mpo_with_qn(g), Mpo(load_mpo(g, "z"))
    om = load_oracle_mps(g, "mps0")
    mps = Mps(om.sites, om.qn, om.sigmaqn, om.qntot, om.qnidx, om.to_right)
    zs = []
    for _ in range(10):
        mps = mps.evolve(mpo, 0.05)
        zs.append(mps.expectation(z))
    assert np.abs(np.array(zs) - g["pc_z_t"]).max() < E_TOL
    assert mps.bond_dims == list(g["pc_bond_dims"])


@pytest.mark.parametrize("tag", ["sbm", "ex", "dm"])
def test_expand_bond_dimension_golden(golden, tag):
    """Mps.expand_bond_dimension(hint_mpo, include_ex=False) (mps.py:1934-2023), the preparation of
    every TDVP-PS run: Mpo.apply, canonicalise, the SVD compressions and compressed_sum on the device.
    Bond dimensions, the norm carried to coeff, the state and its energy equal the reference's; the
    admixed expander (weight coef) is compared where the truncated spectra are non-degenerate (the
    spin-boson chain) -- with degenerate modes the kept multiplet members are arbitrary in LAPACK too.
    The expanded state then takes TDVP-PS steps at the full bond dimension, conserving the energy."""
    from renormalizer_b200.configs import CompressConfig, CompressCriteria
    from renormalizer_b200.mps import Mps
    g = golden("expand")
    mpo = _device_mpo_with_qn(g, f"{tag}_mpo")
    om = load_oracle_mps(g, f"{tag}_pre", meta=f"{tag}_pre")
    pre = Mps(om.sites, om.qn, om.sigmaqn, om.qntot, om.qnidx, om.to_right, coeff=complex(g[f"{tag}_pre_coeff"]))
    pre.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=int(g[f"{tag}_max_bonddim"]))
    new = pre.copy().expand_bond_dimension(mpo, coef=float(g[f"{tag}_coef"]), include_ex=False)
    ref = to_device_mps(load_oracle_mps(g, f"{tag}_post", meta=f"{tag}_post"))
    assert new.bond_dims == list(g[f"{tag}_post_bond_dims"])
    assert abs(new.coeff - complex(g[f"{tag}_post_coeff"])) < E_TOL
    assert abs(new.mp_norm - 1) < 1e-12
    assert abs(ref.conj().dot(new) - 1) < E_TOL
    assert abs(new.expectation(mpo) - float(g[f"{tag}_post_energy"])) < 1e-9
    if tag == "sbm":
        def admixture(post):
            a, b = pre.copy(), post.copy()
            a.coeff = b.coeff = 1
            return b.add(a.scale(-a.conj().dot(b) / a.conj().dot(a)))
        x, y = admixture(new), admixture(ref)
        assert abs(abs(x.conj().dot(y)) / (x.mp_norm * y.mp_norm) - 1) < 1e-4
    with pytest.raises(NotImplementedError):
        pre.copy().expand_bond_dimension(mpo)               # include_ex=True needs a model-specific state
    new.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps)
    e0 = new.expectation(mpo)
    stepped = new.evolve(mpo, 0.05).evolve(mpo, 0.05)
    assert stepped.bond_dims == new.bond_dims
    assert abs(stepped.expectation(mpo) - e0) < 1e-8
    assert abs(stepped.mp_norm - 1) < 1e-12


def test_prop_and_compress_adaptive_golden(golden):
    """Adaptive propagate-and-compress (mps.py:826-880) through Mps.evolve: the controller accepts
    the same sub-steps (bond dimensions and occupations follow the reference); guess_dt is derived
    from the distance of two nearly equal states -- a cancellation of ten digits -- so the
    reference's own value is defined to ~1e-5 only."""
    from renormalizer_b200.configs import CompressConfig, CompressCriteria
    from renormalizer_b200.mpo import Mpo
    g = golden("pc")
    mpo = _device_mpo_with_qn(g)
    occ = [Mpo(load_mpo(g, f"occ{i}")) for i in range(int(g["nmol"]))]
    mps = _device_mps_with_coeff(g, "mps0")
    mps.compress_config = CompressConfig(CompressCriteria.threshold, threshold=1e-5)
    mps.evolve_config = EvolveConfig(EvolveMethod.prop_and_compress, adaptive=True, guess_dt=0.4,
                                     adaptive_rtol=1e-4)
    occs, guesses, dims = [], [], []
    for _ in range(3):
        mps = mps.evolve(mpo, 2.0)
        occs.append([mps.expectation(o) for o in occ])
        guesses.append(mps.evolve_config.guess_dt)
        dims.append(mps.bond_dims)
    assert np.allclose(guesses, g["ada_guess_dt"], rtol=1e-3)
    assert np.array_equal(np.array(dims), g["ada_bond_dims"])
    assert np.abs(np.array(occs) - g["ada_occ"]).max() < 1e-9
    with pytest.raises(ValueError):
        mps.evolve(mpo, -1.0)                            # check_valid_dt: wrong direction


@pytest.mark.parametrize("tag", ["mps", "dm"])
def test_bond_entropy_golden(golden, tag):
    """Mps.calc_bond_singular_values / calc_bond_entropy (mps.py:1759-1793): a compression sweep
    that truncates nothing, on the device SVD."""
    g = golden("entropy")
    mps = to_device_mps(load_oracle_mps(g, tag, meta=tag))
    s = mps.calc_bond_singular_values()
    ref = g[f"{tag}_singular_values"]
    assert s.shape == ref.shape
    assert np.abs(s - ref).max() < 1e-12
    assert np.abs(mps.calc_entropy("bond") - g[f"{tag}_bond_entropy"]).max() < E_TOL
    with pytest.raises(NotImplementedError):
        mps.calc_entropy("1site")


@pytest.mark.parametrize("method", ["1site", "2site"])
def test_variational_compress_golden(golden, method):
    """Mpo.contract(mps, algo="variational") (mp.py:513-650): the default guess (SVD-compressed
    operator times SVD-compressed state), the sweeps with the guess as bra and the state as ket on
    the device environment / H_eff / SVD kernels; bond dimensions, norm, overlap with the exact
    product and the state itself follow the reference."""
    from renormalizer_b200.configs import CompressConfig, CompressCriteria
    from renormalizer_b200.mpo import Mpo
    g = golden("vcompress")
    n = int(g["mpo_n"])
    mpo = Mpo(load_mpo(g), qn=[g[f"mpo_qn_{i}"] for i in range(n + 1)], qntot=g["mpo_qntot"],
              qnidx=int(g["mpo_qnidx"]), sigmaqn=[g[f"mpo_sigmaqn_{i}"] for i in range(n)],
              to_right=bool(g["mpo_to_right"]))
    state = to_device_mps(load_oracle_mps(g, "mps", meta="mps"))
    state.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=8, vmethod=method)
    before = [t.clone() for t in state]
    np.random.seed(0)
    new = mpo.contract(state, algo="variational")
    assert all(torch.equal(a, b) for a, b in zip(before, state))          # `self` is not overwritten
    assert new.bond_dims == list(g[f"{method}_bond_dims"])
    assert abs(new.mp_norm - float(g[f"{method}_norm"])) < 1e-8
    exact = mpo.apply(state)
    assert abs(new.conj().dot(exact) - complex(g[f"{method}_overlap_exact"])) < 1e-8
    ref = to_device_mps(load_oracle_mps(g, f"{method}_new", meta=f"{method}_new"))
    assert abs(abs(ref.conj().dot(new)) / (ref.mp_norm * new.mp_norm) - 1) < T_TOL
    assert abs(new.conj().dot(exact)) / (new.mp_norm * exact.mp_norm) > 1 - 1e-6


@pytest.mark.parametrize("tag", ["rk4", "rkf", "rk3"])
def test_prop_and_compress_runge_kutta_golden(golden, tag):
    """Mps.evolve with the Runge-Kutta propagate-and-compress integrators (mps.py:664-793): classical
    RK4 with the MPO given as a function of time, the tableau integrator with the embedded Fehlberg pair
    and adaptive step control, and a fixed-step third-order tableau; Mpo.contract, Mps.add and the SVD
    compression run on the device."""
    from renormalizer_b200.configs import CompressConfig, CompressCriteria, EvolveConfig
    g = golden("pc")
    mpo = _device_mpo_with_qn(g)
    from renormalizer_b200.mpo import Mpo
    occ = [Mpo(load_mpo(g, f"occ{i}")) for i in range(int(g["nmol"]))]
    mps = _device_mps_with_coeff(g, "mps0")
    mps.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=10)
    mps.evolve_config = {
        "rk4": EvolveConfig(EvolveMethod.prop_and_compress_tdrk4),
        "rkf": EvolveConfig(EvolveMethod.prop_and_compress_tdrk, adaptive=True, guess_dt=0.3, adaptive_rtol=1e-4,
                            rk_solver="RKF45"),
        "rk3": EvolveConfig(EvolveMethod.prop_and_compress_tdrk, rk_solver="Kutta_RK3")}[tag]
    assert not mps.evolve_config.is_tdvp
    occs, guesses, dims = [], [], []
    for _ in range(3):
        mps = mps.evolve((lambda t, *a, **k: mpo) if tag == "rk4" else mpo, 0.5)
        occs.append([mps.expectation(o) for o in occ])
        guesses.append(mps.evolve_config.guess_dt)
        dims.append(mps.bond_dims)
    assert np.array_equal(np.array(dims), g[f"{tag}_bond_dims"])
    assert np.abs(np.array(occs) - g[f"{tag}_occ"]).max() < E_TOL
    assert abs(mps.expectation(mpo) - float(g[f"{tag}_energy"])) < E_TOL
    assert np.allclose(guesses, g[f"{tag}_guess_dt"], rtol=1e-6)
    ref = to_device_mps(load_oracle_mps(g, f"{tag}_mpsT", meta="mps0"))
    assert abs(abs(ref.conj().dot(mps)) - 1) < T_TOL

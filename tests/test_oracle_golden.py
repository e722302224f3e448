"""Pin the CPU oracle against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import contract, svdqn
from oracle.krylov import expm_krylov
from oracle.davidson import davidson
from oracle.sweep import optimize_mps, evolve_tdvp_ps
from helpers import load_mpo, load_oracle_mps, relerr

TOL = 1e-13


@pytest.mark.parametrize("t", ["r", "c"])
def test_hop_and_env(golden, t):
    g = golden("kernels")
    L, R, R1, R0, W1, W2 = (g[f"{t}_{k}"] for k in ("L", "R", "R1", "R0", "W1", "W2"))
    assert relerr(contract.hop_apply(L, R0, [], g[f"{t}_C0"]), g[f"{t}_hop0"]) < TOL
    assert relerr(contract.hop_apply(L, R1, [W1], g[f"{t}_C1"]), g[f"{t}_hop1"]) < TOL
    assert relerr(contract.hop_apply(L, R, [W1, W2], g[f"{t}_C2"]), g[f"{t}_hop2"]) < TOL
    assert relerr(contract.hop_apply(L, R1, [W1], g[f"{t}_C1a"]), g[f"{t}_hop1a"]) < TOL
    assert relerr(contract.hop_apply(L, R, [W1, W2], g[f"{t}_C2a"]), g[f"{t}_hop2a"]) < TOL
    assert relerr(contract.env_update(L, g[f"{t}_A3"], W1, "L"), g[f"{t}_envL3"]) < TOL
    assert relerr(contract.env_update(L, g[f"{t}_A4"], W1, "L"), g[f"{t}_envL4"]) < TOL
    assert relerr(contract.env_update(R1, g[f"{t}_A3"], W1, "R"), g[f"{t}_envR3"]) < TOL
    assert relerr(contract.env_update(R1, g[f"{t}_A4"], W1, "R"), g[f"{t}_envR4"]) < TOL


@pytest.mark.parametrize("t", ["r", "c"])
def test_hop_two_layer(golden, t):
    """Oracle restatement of the two-layer expressions pinned to the reference's own output."""
    g = golden("kernels")
    L4, R41, R42, W1, W2 = (g[f"{t}_{k}"] for k in ("L4", "R41", "R42", "W1", "W2"))
    got = contract.hop_apply_two_layer(L4, R41, [W1], g[f"{t}_C1"])
    assert relerr(got, g[f"{t}_hop1_2l"]) < 1e-13
    got = contract.hop_apply_two_layer(L4, R42, [W1, W2], g[f"{t}_C2"])
    assert relerr(got, g[f"{t}_hop2_2l"]) < 1e-13


def test_hop_diag_matches_dense_diagonal(golden):
    g = golden("kernels")
    L, R1, R, W1, W2 = g["r_L"], g["r_R1"], g["r_R"], g["r_W1"], g["r_W2"]
    dense = np.einsum("abc,bdef,lfk->adlcek", L, W1, R1)
    n = dense.shape[0] * dense.shape[1] * dense.shape[2]
    assert relerr(contract.hop_diag(L, R1, [W1]).ravel(), np.diag(dense.reshape(n, n))) < TOL
    dense = np.einsum("abc,bdef,fghj,ljk->adglcehk", L, W1, W2, R)
    n = int(np.prod(dense.shape[:4]))
    assert relerr(contract.hop_diag(L, R, [W1, W2]).ravel(), np.diag(dense.reshape(n, n))) < TOL


@pytest.mark.parametrize("t", ["r", "c"])
@pytest.mark.parametrize("system", ["L", "R"])
def test_svd_qn(golden, t, system):
    g = golden("svdqn")
    k = f"{t}_{system}"
    c, ql, qr, qntot = g[k + "_c"], g[k + "_qnbigl"], g[k + "_qnbigr"], g["qntot"]
    u, su, qnl, v, sv, qnr = svdqn.svd_qn(c, ql, qr, qntot, system=system, full_matrices=False)
    assert relerr(su, g[k + "_svd_s"]) < TOL
    assert relerr(u, g[k + "_svd_u"]) < 1e-10 and relerr(v, g[k + "_svd_v"]) < 1e-10
    assert np.array_equal(np.array(qnl), g[k + "_svd_qnl"])
    assert np.array_equal(np.array(qnr), g[k + "_svd_qnr"])
    np.random.seed(11)
    u, su, qnl, v, sv, qnr = svdqn.svd_qn(c, ql, qr, qntot, system=system, full_matrices=True)
    assert u.shape == g[k + "_fsvd_u"].shape
    assert relerr(u, g[k + "_fsvd_u"]) < 1e-10 and relerr(v, g[k + "_fsvd_v"]) < 1e-10
    assert relerr(su, g[k + "_fsvd_su"]) < TOL and relerr(sv, g[k + "_fsvd_sv"]) < TOL
    assert np.array_equal(np.array(qnl), g[k + "_fsvd_qnl"])
    u, qnl, v, qnr = svdqn.svd_qn(c, ql, qr, qntot, QR=True, system=system, full_matrices=False)
    assert relerr(u, g[k + "_qr_u"]) < 1e-12 and relerr(v, g[k + "_qr_v"]) < 1e-12
    assert np.array_equal(np.array(qnl), g[k + "_qr_qnl"])
    assert np.array_equal(np.array(qnr), g[k + "_qr_qnr"])


def test_svd_qn_invalid_qn_raises():
    c = np.ones((2, 2, 2))
    qnl = np.zeros((2, 1), dtype=int)
    sig = np.zeros((2, 1), dtype=int)
    with pytest.raises(ValueError):
        svdqn.svd_qn(c, svdqn.add_outer(qnl, sig), qnl, np.array([5]), system="L")


def test_expm_krylov(golden):
    g = golden("krylov")
    h, v = g["h"], g["v"]
    for i in range(3):
        res, j = expm_krylov(lambda y: h @ y, complex(g[f"dt{i}"]), v.copy())
        assert j == int(g[f"j{i}"])
        assert relerr(res, g[f"res{i}"]) < TOL


@pytest.mark.parametrize("nroots", [1, 3])
def test_davidson(golden, nroots):
    g = golden("davidson")
    a = g["a"]
    hd = np.diag(a).copy()
    count = [0]

    def hop(x):
        count[0] += 1
        return a @ x
    e, c = davidson(hop, [x.copy() for x in g[f"x0_{nroots}"]],
                    lambda x, e, *args: x / (hd - e + 1e-4), max_cycle=100, nroots=nroots)
    assert count[0] == int(g[f"nhop_{nroots}"])
    assert np.abs(np.array(e) - g[f"e_{nroots}"]).max() < 1e-13
    assert np.abs(np.array(c) - g[f"c_{nroots}"]).max() < 1e-10


@pytest.mark.parametrize("method", ["1site", "2site"])
def test_dmrg_holstein_trajectory(golden, method):
    """Whole optimize_mps trajectory: every micro-iteration energy of every sweep."""
    g = golden("holstein")
    mpo = load_mpo(g)
    mps = load_oracle_mps(g, "mps0")
    np.random.seed(99)
    micro = []
    proc = [(int(a), float(b)) for a, b in g["procedure"]]
    e, opt = optimize_mps(mps, mpo, proc, method=method, micro_out=micro)
    assert len(e) == len(g[f"{method}_energies"])
    assert np.abs(np.array(e) - g[f"{method}_energies"]).max() < 1e-12
    for i, m in enumerate(micro):
        assert np.abs(m - g[f"{method}_micro_{i}"]).max() < 1e-12
    assert abs(opt.expectation(mpo) - float(g[f"{method}_expectation"])) < 1e-12
    # the reference's own acceptance value (mps/tests/test_gs.py:22,36)
    assert e[-1] == pytest.approx(0.08401412 + float(g["gs_zpe"]), rel=1e-5)


def test_tdvp_ps_spin_boson(golden):
    g = golden("sbm")
    mpo = load_mpo(g)
    sz = load_mpo(g, "sigma_z")
    mps = load_oracle_mps(g, "mps0")
    dt = float(g["dt"])
    szs, es = [mps.expectation(sz)], [mps.expectation(mpo)]
    for i in range(int(g["nsteps"])):
        mps = evolve_tdvp_ps(mps, mpo, dt)
        szs.append(mps.expectation(sz))
        es.append(mps.expectation(mpo))
        if i == 0:
            ref1 = load_oracle_mps(g, "mps1")
            # same physical state (site tensors differ by a gauge on rank-deficient bonds)
            assert abs(abs(ref1.dot_conj(mps)) - 1) < 1e-12
    assert np.abs(np.array(szs) - g["sigma_z_t"]).max() < 1e-11
    assert np.abs(np.array(es) - g["energy_t"]).max() < 1e-12
    refT = load_oracle_mps(g, "mpsT")
    assert abs(abs(refT.dot_conj(mps)) - 1) < 1e-11


def test_tdvp_ps2_spin_boson(golden):
    """Two-site projector splitting (mps.py:1407-1517) pinned to the reference's trajectory."""
    from oracle.sweep import evolve_tdvp_ps2
    g = golden("sbm")
    mpo = load_mpo(g)
    sz = load_mpo(g, "sigma_z")
    mps = load_oracle_mps(g, "mps0")
    dt = float(g["dt"])
    szs, es = [mps.expectation(sz)], [mps.expectation(mpo)]
    for i in range(int(g["ps2_nsteps"])):
        mps = evolve_tdvp_ps2(mps, mpo, dt, 12)
        szs.append(mps.expectation(sz))
        es.append(mps.expectation(mpo))
    assert np.abs(np.array(szs) - g["ps2_sigma_z_t"]).max() < 1e-10
    assert np.abs(np.array(es) - g["ps2_energy_t"]).max() < 1e-10
    assert [s.shape[0] for s in mps.sites] + [1] == list(g["ps2_bond_dims"])
    ref = load_oracle_mps(g, "ps2_mpsT")
    ov = mps.conj().dot(ref) if hasattr(mps, "conj") else None
    if ov is not None:
        assert abs(abs(ov) - 1) < 1e-8


@pytest.mark.parametrize("method", ["1site", "2site"])
def test_dmrg_state_averaged(golden, method):
    """State-averaged DMRG for the three lowest states (gs.py nroots > 1, mp.py:780-838, eigh_qn
    svd_qn.py:243-302): every sweep energy of every root reproduces the reference run."""
    g = golden("holstein")
    mpo = load_mpo(g)
    mps = load_oracle_mps(g, "mps0")
    proc = [(int(a), float(b)) for a, b in g["procedure"]]
    np.random.seed(99)
    e, opts = optimize_mps(mps, mpo, proc, method=method, nroots=3)
    assert np.abs(np.array(e) - g[f"sa_{method}_energies"]).max() < 1e-12
    got = np.array([o.expectation(mpo) for o in opts])
    assert np.abs(got - g[f"sa_{method}_expectations"]).max() < 1e-12
    # the reference's own acceptance values (mps/tests/test_gs.py:80)
    std = np.array([0.08401412, 0.08449771, 0.08449801]) + float(g["gs_zpe"])
    assert np.allclose(got, std)


# ---------------------------------------------------------------- BASELINE.json parity cases
def _run_oracle_dmrg(g, mpo, method, prefix="mps0", **kw):
    from oracle.sweep import optimize_mps
    mps = load_oracle_mps(g, prefix)
    np.random.seed(99)
    return optimize_mps(mps, mpo, [tuple(p) for p in g["procedure"]], method=method, **kw)


@pytest.mark.parametrize("method", ["1site", "2site"])
def test_dmrg_stacked_mpo(golden, method):
    """StackedMpo (mpo.py:483-494; mps/tests/test_gs.py:148-158): H + H as two members gives twice
    the energies; the trajectory reproduces the reference's."""
    from oracle.sweep import StackedMpo
    g = golden("stacked")
    mpo = load_mpo(g)
    e2, _ = _run_oracle_dmrg(g, StackedMpo([mpo, mpo]), method)
    assert np.abs(np.array(e2) - g[f"{method}_double_energies"]).max() < 1e-12
    assert np.abs(np.array(e2) - 2 * g[f"{method}_single_energies"]).max() < 1e-8   # test_gs.py:158


def test_dmrg_stacked_split_hamiltonian(golden):
    from oracle.sweep import StackedMpo
    g = golden("stacked")
    e, opt = _run_oracle_dmrg(g, StackedMpo([load_mpo(g, "mpo_a"), load_mpo(g, "mpo_b")]), "2site")
    assert np.abs(np.array(e) - g["split_energies"]).max() < 1e-12
    assert abs(opt.expectation(load_mpo(g)) - float(g["split_expectation"])) < 1e-12


def test_dmrg_qc_h6_two_quantum_numbers(golden):
    """BASELINE configs[4] in miniature: ab initio DMRG from the reference's H6 FCIDUMP
    (mps/tests/test_gs.py:103-145), two conserved quantum numbers, M = 30."""
    g = golden("qc_h6")
    mpo = load_mpo(g)
    assert load_oracle_mps(g, "mps0").qntot.shape == (2,)
    e, opt = _run_oracle_dmrg(g, mpo, "2site")
    assert np.abs(np.array(e) - g["energies"]).max() < 1e-10
    assert abs(opt.expectation(mpo) - float(g["expectation"])) < 1e-10
    assert np.allclose(min(e), float(g["fci_e"]), atol=5e-3)                          # test_gs.py:145


def _load_with_coeff(g, prefix):
    mps = load_oracle_mps(g, prefix, meta=prefix)
    mps.coeff = complex(g[prefix + "_coeff"])
    return mps


def test_tdvp_ps_exciton_qn_blocked(golden):
    """BASELINE configs[3] in miniature (example/fmo.py): long-range J, one conserved exciton."""
    g = golden("exciton")
    mpo = load_mpo(g)
    occ = [load_mpo(g, f"occ{i}") for i in range(int(g["nmol"]))]
    mps = _load_with_coeff(g, "mps0")
    occs, es = [[mps.expectation(o) for o in occ]], [mps.expectation(mpo)]
    for _ in range(int(g["nsteps"])):
        mps = evolve_tdvp_ps(mps, mpo, float(g["dt"]))
        occs.append([mps.expectation(o) for o in occ])
        es.append(mps.expectation(mpo))
    assert np.abs(np.array(occs) - g["occ_t"]).max() < 1e-10
    assert np.abs(np.array(es) - g["energy_t"]).max() < 1e-10
    assert abs(np.sum(occs[-1]) - 1) < 1e-10
    assert mps.bond_dims == list(g["bond_dims"])


def test_tdvp_ps_density_operator(golden):
    """MpDm sites carry an ancilla index (hop_expr.py:83-117, lib.py:213-262)."""
    g = golden("exciton")
    mpo = load_mpo(g)
    occ = [load_mpo(g, f"occ{i}") for i in range(int(g["nmol"]))]
    dm = _load_with_coeff(g, "dm0")
    assert dm.sites[0].ndim == 4
    occs, es = [[dm.expectation(o) for o in occ]], [dm.expectation(mpo)]
    for _ in range(int(g["dm_nsteps"])):
        dm = evolve_tdvp_ps(dm, mpo, float(g["dt"]))
        occs.append([dm.expectation(o) for o in occ])
        es.append(dm.expectation(mpo))
    assert np.abs(np.array(occs) - g["dm_occ_t"]).max() < 1e-10
    assert np.abs(np.array(es) - g["dm_energy_t"]).max() < 1e-10


def test_two_spin_quickstart(golden):
    """BASELINE configs[0], the README quickstart (README.md:36-58): two half spins, 10 steps of
    dt = 0.05.  The README's own integrator (propagate-and-compress RK4) is outside the sweep
    path; the two-site TDVP sweep reproduces the reference's tdvp_ps2 run to 1e-12, the README's
    printed values to the RK4 error, and both follow -cos(2t)."""
    from oracle.sweep import evolve_tdvp_ps2
    g = golden("two_spin")
    mpo, z = load_mpo(g), load_mpo(g, "z")
    mps = load_oracle_mps(g, "mps0")
    zs = []
    for _ in range(10):
        mps = evolve_tdvp_ps2(mps, mpo, 0.05, 32)
        zs.append(mps.expectation(z))
    assert np.abs(np.array(zs) - g["ps2_z_t"]).max() < 1e-12
    assert np.abs(np.array(zs) - g["pc_z_t"]).max() < 1e-6
    assert np.abs(np.array(zs) + np.cos(2 * 0.05 * np.arange(1, 11))).max() < 1e-6


def test_thermal_imaginary_then_real_time(golden):
    """Finite temperature: imaginary-time TDVP-PS of a density operator (what
    mps/thermalprop.py:96-98 does every step), then real-time steps of the thermal state."""
    g = golden("thermal")
    mpo = load_mpo(g)
    occ = [load_mpo(g, f"occ{i}") for i in range(int(g["nmol"]))]
    dm = _load_with_coeff(g, "dm0")
    dbeta = float(g["beta"]) / int(g["nbeta"])
    occs, es = [[dm.expectation(o) for o in occ]], [dm.expectation(mpo)]
    for _ in range(int(g["nbeta"])):
        dm = evolve_tdvp_ps(dm, mpo, -0.5j * dbeta)
        occs.append([dm.expectation(o) for o in occ])
        es.append(dm.expectation(mpo))
    # Imaginary time is not unitary: rounding differences between two correct evaluations of the
    # same Lanczos recurrence are amplified by exp(dbeta * spectral width) per step (observed
    # 8e-12 after one step, 1.4e-10 after four), so this leg is held to 1e-9, not 1e-10.
    assert np.abs(np.array(occs) - g["imag_occ"]).max() < 1e-9
    assert np.abs(np.array(es) - g["imag_energy"]).max() < 1e-9
    assert np.abs(np.array(occs[1]) - g["imag_occ"][1]).max() < 1e-10
    assert es[-1] < es[0]                                   # cooling
    rocc, ren = [], []
    for _ in range(2):
        dm = evolve_tdvp_ps(dm, mpo, 2.0)
        rocc.append([dm.expectation(o) for o in occ])
        ren.append(dm.expectation(mpo))
    assert np.abs(np.array(rocc) - g["real_occ"]).max() < 1e-9
    assert np.abs(np.array(ren) - g["real_energy"]).max() < 1e-9


def test_adaptive_tdvp_ps(golden):
    """adaptive_tdvp (mps.py:46-115): same accepted sub-steps, same guess_dt, same observables."""
    from oracle.sweep import evolve_adaptive_tdvp_ps
    g = golden("thermal")
    mpo = load_mpo(g)
    occ = [load_mpo(g, f"occ{i}") for i in range(int(g["nmol"]))]
    mps = _load_with_coeff(g, "mps0")
    guess = 1.0
    occs, guesses = [], []
    for _ in range(3):
        mps, guess = evolve_adaptive_tdvp_ps(mps, mpo, 4.0, guess)
        occs.append([mps.expectation(o) for o in occ])
        guesses.append(guess)
    assert np.allclose(guesses, g["adaptive_guess_dt"], rtol=1e-8)
    # the controller doubles the step to dt = 4: the local Krylov spaces get large and the
    # reference's own stopping rule (successive iterates numpy.allclose, atol 1e-8) is what bounds
    # the agreement of two evaluations of the recurrence; the first (dt = 1, 2) step agrees to 1e-10
    assert np.abs(np.array(occs[0]) - g["adaptive_occ"][0]).max() < 1e-10
    assert np.abs(np.array(occs) - g["adaptive_occ"]).max() < 1e-7


@pytest.mark.parametrize("tag", ["thr", "fix"])
def test_prop_and_compress_exciton(golden, tag):
    """Propagate-and-compress (the default Mps.evolve, mps.py:796-884) with a conserved exciton:
    occupations, energy and the bond dimensions the truncation chooses, step by step."""
    from helpers import load_oracle_mpo
    from oracle.sweep import evolve_prop_and_compress, CompressSpec
    g = golden("pc")
    mpo = load_oracle_mpo(g)
    occ = [load_mpo(g, f"occ{i}") for i in range(int(g["nmol"]))]
    mps = _load_with_coeff(g, "mps0")
    spec = CompressSpec("threshold", threshold=1e-5) if tag == "thr" else CompressSpec("fixed", max_bonddim=12)
    occs, es, dims = [], [], []
    for _ in range(4):
        mps = evolve_prop_and_compress(mps, mpo, 1.0, spec)
        occs.append([mps.expectation(o) for o in occ])
        es.append(mps.expectation(mpo.sites))
        dims.append(mps.bond_dims)
    assert np.array_equal(np.array(dims), g[f"{tag}_bond_dims"])
    assert np.abs(np.array(occs) - g[f"{tag}_occ"]).max() < 1e-10
    assert np.abs(np.array(es) - g[f"{tag}_energy"]).max() < 1e-10
    ref = load_oracle_mps(g, f"{tag}_mpsT", meta="mps0")
    assert abs(abs(ref.dot_conj(mps)) - 1) < 1e-10


def test_two_spin_quickstart_default_integrator(golden):
    """The README quickstart exactly as printed: Mps.evolve with the default (propagate-and-compress)
    configuration, ten steps of 0.05 -- the values the reference prints."""
    from helpers import load_oracle_mpo
    from oracle.sweep import evolve_prop_and_compress, CompressSpec
    g = golden("two_spin")
    mpo, z = load_oracle_mpo(g), load_mpo(g, "z")
    mps = load_oracle_mps(g, "mps0")
    zs = []
    for _ in range(10):
        mps = evolve_prop_and_compress(mps, mpo, 0.05, CompressSpec())
        zs.append(mps.expectation(z))
    assert np.abs(np.array(zs) - g["pc_z_t"]).max() < 1e-12
    assert mps.bond_dims == list(g["pc_bond_dims"])


@pytest.mark.parametrize("tag", ["sbm", "ex", "dm"])
def test_expand_bond_dimension(golden, tag):
    """expand_bond_dimension with a hint MPO (mps.py:1934-2023), the preparation step of every
    TDVP-PS run: bond dimensions, norm carried to coeff, the whole state, and -- separately, because
    it enters with weight coef -- the admixed expander."""
    from helpers import load_oracle_mpo
    from oracle.sweep import expand_bond_dimension, mps_add, mps_scale
    g = golden("expand")
    mpo = load_oracle_mpo(g, f"{tag}_mpo")
    pre = load_oracle_mps(g, f"{tag}_pre", meta=f"{tag}_pre")
    pre.coeff = complex(g[f"{tag}_pre_coeff"])
    coef = float(g[f"{tag}_coef"])
    new = expand_bond_dimension(pre.copy(), mpo, int(g[f"{tag}_max_bonddim"]), coef)
    ref = load_oracle_mps(g, f"{tag}_post", meta=f"{tag}_post")
    assert new.bond_dims == list(g[f"{tag}_post_bond_dims"])
    assert abs(new.coeff - complex(g[f"{tag}_post_coeff"])) < 1e-12
    assert abs(new.mp_norm - 1) < 1e-12
    assert abs(ref.dot_conj(new) - 1) < 1e-12
    assert abs(new.expectation(mpo.sites) - float(g[f"{tag}_post_energy"])) < 1e-10
    # the expander: (post - <pre|post> pre) / coef has the same direction in both
    def admixture(post):
        a = pre.copy(); a.coeff = 1
        b = post.copy(); b.coeff = 1
        ov = a.dot_conj(b) / a.dot_conj(a)
        return mps_add(b, mps_scale(a, -ov))
    x, y = admixture(new), admixture(ref)
    assert abs(abs(x.dot_conj(y)) / (x.mp_norm * y.mp_norm) - 1) < 1e-6
    assert abs(x.mp_norm / y.mp_norm - 1) < 1e-6


def test_prop_and_compress_adaptive(golden):
    """Adaptive propagate-and-compress (mps.py:826-880): accepted sub-steps, guess_dt, bond
    dimensions and occupations follow the reference."""
    from helpers import load_oracle_mpo
    from oracle.sweep import evolve_prop_and_compress_adaptive, CompressSpec
    g = golden("pc")
    mpo = load_oracle_mpo(g)
    occ = [load_mpo(g, f"occ{i}") for i in range(int(g["nmol"]))]
    mps = _load_with_coeff(g, "mps0")
    spec = CompressSpec("threshold", threshold=1e-5)
    guess = 0.4
    occs, guesses, dims = [], [], []
    for _ in range(3):
        mps, guess = evolve_prop_and_compress_adaptive(mps, mpo, 2.0, spec, guess, rtol=1e-4)
        occs.append([mps.expectation(o) for o in occ])
        guesses.append(guess)
        dims.append(mps.bond_dims)
    # guess_dt comes from the distance of two nearly equal states, sqrt(l1 + l2 - 2 Re l12): a
    # cancellation of ten digits, so the reference's own value is defined to ~1e-5 only
    assert np.allclose(guesses, g["ada_guess_dt"], rtol=1e-3)
    assert np.array_equal(np.array(dims), g["ada_bond_dims"])
    assert np.abs(np.array(occs) - g["ada_occ"]).max() < 1e-10


@pytest.mark.parametrize("tag", ["mps", "dm"])
def test_bond_entropy(golden, tag):
    """Bond singular values and von Neumann entropies (mps.py:1759-1793)."""
    from oracle.sweep import calc_bond_singular_values, calc_bond_entropy
    g = golden("entropy")
    mps = load_oracle_mps(g, tag, meta=tag)
    s = calc_bond_singular_values(mps)
    assert s.shape == g[f"{tag}_singular_values"].shape
    assert np.abs(s - g[f"{tag}_singular_values"]).max() < 1e-12
    assert np.abs(calc_bond_entropy(mps) - g[f"{tag}_bond_entropy"]).max() < 1e-10


@pytest.mark.parametrize("method", ["1site", "2site"])
def test_variational_compress(golden, method):
    """Mpo.contract(mps, algo="variational") (mp.py:513-650): the oracle restatement follows the
    reference (bond dimensions, norm, overlap with the exact product, the state itself).  The
    device implementation is next round's work; this pins what it will be checked against."""
    from helpers import load_oracle_mpo
    from oracle.sweep import variational_compress, mpo_apply
    g = golden("vcompress")
    mpo = load_oracle_mpo(g)
    n = len(mpo)
    state = load_oracle_mps(g, "mps", meta="mps")
    np.random.seed(0)
    new = variational_compress(state, mpo, [g[f"mpo_sigmaqn_{i}"] for i in range(n)], bool(g["mpo_to_right"]),
                               max_bonddim=8, method=method)
    assert new.bond_dims == list(g[f"{method}_bond_dims"])
    assert abs(new.mp_norm - float(g[f"{method}_norm"])) < 1e-9
    exact = mpo_apply(mpo, state)
    assert abs(new.dot_conj(exact) - complex(g[f"{method}_overlap_exact"])) < 1e-9
    ref = load_oracle_mps(g, f"{method}_new", meta=f"{method}_new")
    assert abs(abs(ref.dot_conj(new)) / (ref.mp_norm * new.mp_norm) - 1) < 1e-8
    # a compression: close to the exact product
    assert abs(new.dot_conj(exact)) / (new.mp_norm * exact.mp_norm) > 1 - 1e-6


@pytest.mark.parametrize("tag", ["rk4", "rkf", "rk3"])
def test_prop_and_compress_runge_kutta(golden, tag):
    """The Runge-Kutta propagate-and-compress integrators (mps.py:664-793): classical RK4 with the MPO
    given as a function of time, the tableau integrator with the embedded Fehlberg pair and adaptive
    step control, and a fixed-step third-order tableau -- occupations, bond dimensions, guess_dt and
    the final state follow the reference (evolve's "mps_only" normalisation included)."""
    from helpers import load_oracle_mpo
    from oracle.sweep import evolve_pc_tdrk4, evolve_pc_tdrk, CompressSpec, normalize
    from renormalizer_b200.rk import RungeKutta
    g = golden("pc")
    mpo = load_oracle_mpo(g)
    occ = [load_mpo(g, f"occ{i}") for i in range(int(g["nmol"]))]
    mps = _load_with_coeff(g, "mps0")
    spec = CompressSpec("fixed", max_bonddim=10)
    guess = 0.3 if tag == "rkf" else 0.1
    occs, guesses, dims = [], [], []
    for _ in range(3):
        if tag == "rk4":
            mps = evolve_pc_tdrk4(mps, lambda t: mpo, 0.5, spec)
        else:
            rk = RungeKutta("RKF45" if tag == "rkf" else "Kutta_RK3")
            mps, guess = evolve_pc_tdrk(mps, lambda t: mpo, 0.5, spec, rk.tableau, rk.order, adaptive=tag == "rkf",
                                        guess_dt=guess, rtol=1e-4)
        normalize(mps, "mps_only")
        occs.append([mps.expectation(o) for o in occ])
        guesses.append(guess)
        dims.append(mps.bond_dims)
    assert np.array_equal(np.array(dims), g[f"{tag}_bond_dims"])
    assert np.abs(np.array(occs) - g[f"{tag}_occ"]).max() < 1e-10
    assert abs(mps.expectation(mpo.sites) - float(g[f"{tag}_energy"])) < 1e-10
    assert np.allclose(guesses, g[f"{tag}_guess_dt"], rtol=1e-6)
    ref = load_oracle_mps(g, f"{tag}_mpsT", meta="mps0")
    assert abs(abs(ref.dot_conj(mps)) - 1) < 1e-10

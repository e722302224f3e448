#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) in the build
container.  TEST TOOLING ONLY -- never imported by the product, never run on the GPU box.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference needs three third-party packages that this image lacks (opt_einsum, h5py,
print_tree2); `tests/golden/_shims/` holds minimal stand-ins for them (einsum -> numpy.einsum).
Every array below is produced by reference code paths:

  kernels.npz   hop_expr (mps/hop_expr.py:7), contract_one_site (mps/lib.py:172)
  svdqn.npz     svd_qn (mps/svd_qn.py:97) in SVD and QR modes
  krylov.npz    expm_krylov (lib/krylov/krylov.py:28)
  davidson.npz  davidson (lib/davidson/davidson.py:73)
  holstein.npz  Mpo(holstein_model), Mps.random, optimize_mps (mps/gs.py:54), 1site + 2site
  sbm.npz       SpinBosonModel MPO, expand_bond_dimension'd MPS, Mps.evolve with tdvp_ps
                (mps/mps.py:1268) and tdvp_ps2 (mps/mps.py:1407), sigma_z trajectory and final MPS
  stacked.npz   optimize_mps with StackedMpo (mps/mpo.py:483, mps/tests/test_gs.py:148)
  qc_h6.npz     ab initio DMRG on the reference's H6 FCIDUMP, two quantum numbers (test_gs.py:103)
  exciton.npz   FMO-like HolsteinModel with long-range J: TDVP-PS of the one-exciton state and of
                its density operator (MpDm)
  thermal.npz   imaginary-time TDVP-PS of a density operator (mps/thermalprop.py:96), real-time
                steps of the thermal state, adaptive_tdvp (mps/mps.py:46)
  two_spin.npz  the README quickstart with the default integrator, tdvp_ps and tdvp_ps2
  pc.npz        propagate-and-compress (mps/mps.py:796): fixed step with threshold / fixed
                truncation, adaptive step
  expand.npz    expand_bond_dimension(hint_mpo, include_ex=False) (mps/mps.py:1934)
  entropy.npz   calc_bond_singular_values / calc_bond_entropy (mps/mps.py:1759)
  vcompress.npz Mpo.contract(mps, algo="variational") (mps/mp.py:513)
"""
import os
import sys
import logging

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402

logging.disable(logging.INFO)

from renormalizer.mps.hop_expr import hop_expr  # noqa: E402
from renormalizer.mps.lib import contract_one_site  # noqa: E402
from renormalizer.mps import svd_qn as ref_svd_qn  # noqa: E402
from renormalizer.lib.krylov.krylov import expm_krylov  # noqa: E402
from renormalizer.lib import davidson  # noqa: E402


def crand(rng, shape, cplx):
    a = rng.standard_normal(shape)
    if cplx:
        a = a + 1j * rng.standard_normal(shape)
    return a


def gen_kernels():
    rng = np.random.default_rng(20261017)
    out = {}
    for cplx in (False, True):
        tag = "c" if cplx else "r"
        Ml, Mr, w0, w1, w2, d1, d2, anc = 7, 9, 3, 4, 5, 3, 4, 2
        L = crand(rng, (Ml, w0, Ml), cplx)
        R = crand(rng, (Mr, w2, Mr), cplx)
        W1 = crand(rng, (w0, d1, d1, w1), False)
        W2 = crand(rng, (w1, d2, d2, w2), False)
        # make the MPO site sparse like real ones
        W1[rng.random(W1.shape) < 0.5] = 0
        W2[rng.random(W2.shape) < 0.5] = 0
        # one-site: R must carry W1's right bond
        R1 = crand(rng, (Mr, w1, Mr), cplx)
        C1 = crand(rng, (Ml, d1, Mr), cplx)
        C2 = crand(rng, (Ml, d1, d2, Mr), cplx)
        C1a = crand(rng, (Ml, d1, anc + 1, Mr), cplx)
        C2a = crand(rng, (Ml, d1, anc + 1, d2, anc, Mr), cplx)
        R0 = crand(rng, (Mr, w0, Mr), cplx)
        C0 = crand(rng, (Ml, Mr), cplx)
        out.update({f"{tag}_L": L, f"{tag}_R": R, f"{tag}_R1": R1, f"{tag}_R0": R0,
                    f"{tag}_W1": W1, f"{tag}_W2": W2,
                    f"{tag}_C0": C0, f"{tag}_C1": C1, f"{tag}_C2": C2,
                    f"{tag}_C1a": C1a, f"{tag}_C2a": C2a})
        out[f"{tag}_hop0"] = hop_expr(L, R0, [], C0.shape)(C0)
        out[f"{tag}_hop1"] = hop_expr(L, R1, [W1.copy()], C1.shape)(C1)
        out[f"{tag}_hop2"] = hop_expr(L, R, [W1.copy(), W2.copy()], C2.shape)(C2)
        out[f"{tag}_hop1a"] = hop_expr(L, R1, [W1.copy()], C1a.shape)(C1a)
        out[f"{tag}_hop2a"] = hop_expr(L, R, [W1.copy(), W2.copy()], C2a.shape)(C2a)
        # two-layer (H - omega)^2 expressions (hop_expr.py:24-52): 4-index environments
        L4 = crand(rng, (Ml, w0, w0, Ml), cplx)
        R41 = crand(rng, (Mr, w1, w1, Mr), cplx)
        R42 = crand(rng, (Mr, w2, w2, Mr), cplx)
        out.update({f"{tag}_L4": L4, f"{tag}_R41": R41, f"{tag}_R42": R42})
        out[f"{tag}_hop1_2l"] = hop_expr(L4, R41, [W1.copy()], C1.shape, twolayer=True)(C1)
        out[f"{tag}_hop2_2l"] = hop_expr(L4, R42, [W1.copy(), W2.copy()], C2.shape, twolayer=True)(C2)
        # environment updates: MPS site (ndim 3) and MPDM site (ndim 4)
        A3 = crand(rng, (Ml, d1, Mr), cplx)
        A4 = crand(rng, (Ml, d1, anc, Mr), cplx)
        out[f"{tag}_A3"] = A3
        out[f"{tag}_A4"] = A4
        out[f"{tag}_envL3"] = contract_one_site(L, A3, W1, "L")
        out[f"{tag}_envL4"] = contract_one_site(L, A4, W1, "L")
        out[f"{tag}_envR3"] = contract_one_site(R1, A3, W1, "R")
        out[f"{tag}_envR4"] = contract_one_site(R1, A4, W1, "R")
    np.savez_compressed(os.path.join(HERE, "kernels.npz"), **out)


def gen_svdqn():
    rng = np.random.default_rng(7)
    out = {}
    # one-site tensor (l, sigma, r) with a U(1) quantum number, total qn = 1
    ml, d, mr = 6, 3, 5
    qnl = rng.integers(0, 2, size=(ml, 1))
    sigmaqn = np.array([[0], [1], [0]])
    qnr = rng.integers(0, 2, size=(mr, 1))
    qntot = np.array([1])
    for cplx in (False, True):
        tag = "c" if cplx else "r"
        for system in ("L", "R"):
            if system == "L":
                qnbigl = ref_svd_qn.add_outer(qnl, sigmaqn)
                qnbigr = qnr
            else:
                qnbigl = qnl
                qnbigr = ref_svd_qn.add_outer(sigmaqn, qnr)
            qnmat = ref_svd_qn.add_outer(qnbigl, qnbigr)
            mask = ref_svd_qn.get_qn_mask(qnmat, qntot)
            c = crand(rng, (ml, d, mr), cplx)
            c[~mask] = 0
            key = f"{tag}_{system}"
            out[key + "_c"] = c
            out[key + "_qnbigl"] = qnbigl
            out[key + "_qnbigr"] = qnbigr
            u, su, qnlnew, v, sv, qnrnew = ref_svd_qn.svd_qn(
                c, qnbigl, qnbigr, qntot, system=system, full_matrices=False)
            out[key + "_svd_u"], out[key + "_svd_s"], out[key + "_svd_v"] = u, su, v
            out[key + "_svd_qnl"], out[key + "_svd_qnr"] = np.array(qnlnew), np.array(qnrnew)
            np.random.seed(11)
            u, su, qnlnew, v, sv, qnrnew = ref_svd_qn.svd_qn(
                c, qnbigl, qnbigr, qntot, system=system, full_matrices=True)
            out[key + "_fsvd_u"], out[key + "_fsvd_su"], out[key + "_fsvd_v"] = u, su, v
            out[key + "_fsvd_sv"] = sv
            out[key + "_fsvd_qnl"], out[key + "_fsvd_qnr"] = np.array(qnlnew), np.array(qnrnew)
            u, qnlnew, v, qnrnew = ref_svd_qn.svd_qn(
                c, qnbigl, qnbigr, qntot, QR=True, system=system, full_matrices=False)
            out[key + "_qr_u"], out[key + "_qr_v"] = u, v
            out[key + "_qr_qnl"], out[key + "_qr_qnr"] = np.array(qnlnew), np.array(qnrnew)
    out["qntot"] = qntot
    np.savez_compressed(os.path.join(HERE, "svdqn.npz"), **out)


def gen_krylov():
    rng = np.random.default_rng(3)
    n = 80
    h = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    h = (h + h.conj().T) / 2
    v = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    out = {"h": h, "v": v}
    for i, dt in enumerate((-0.05j, 0.3j, -0.2)):
        res, j = expm_krylov(lambda y: h @ y, dt, v.copy())
        out[f"dt{i}"] = np.array(dt)
        out[f"res{i}"] = res
        out[f"j{i}"] = np.array(j)
    np.savez_compressed(os.path.join(HERE, "krylov.npz"), **out)


def gen_davidson():
    rng = np.random.default_rng(5)
    n = 300
    a = rng.standard_normal((n, n)) * 0.05
    a = (a + a.T) / 2 + np.diag(np.arange(n) * 0.1)
    hdiag = np.diag(a).copy()
    out = {"a": a}
    for nroots in (1, 3):
        count = [0]

        def hop(x):
            count[0] += 1
            return a @ x
        x0 = [np.eye(n)[:, i] + 0.01 * rng.standard_normal(n) for i in range(nroots)]
        out[f"x0_{nroots}"] = np.array(x0)
        precond = lambda x, e, *args: x / (hdiag - e + 1e-4)  # noqa: E731
        e, c = davidson(hop, [x.copy() for x in x0], precond, max_cycle=100, nroots=nroots,
                        max_memory=64000)
        out[f"e_{nroots}"] = np.array(e)
        out[f"c_{nroots}"] = np.array(c)
        out[f"nhop_{nroots}"] = np.array(count[0])
    np.savez_compressed(os.path.join(HERE, "davidson.npz"), **out)


def dump_mp(prefix, mp, out):
    out[prefix + "_n"] = np.array(len(mp))
    for i, mt in enumerate(mp):
        out[f"{prefix}_{i}"] = np.asarray(mt.array)


def dump_mpo_meta(prefix, mpo, out):
    """Quantum numbers of an MPO (needed by Mpo.apply, mpo.py:375-382)."""
    out[prefix + "_qntot"] = np.array(mpo.qntot)
    out[prefix + "_qnidx"] = np.array(mpo.qnidx)
    for i, q in enumerate(mpo.qn):
        out[f"{prefix}_qn_{i}"] = np.array(q)


def dump_mps_meta(prefix, mps, out):
    out[prefix + "_qntot"] = np.array(mps.qntot)
    out[prefix + "_qnidx"] = np.array(mps.qnidx)
    out[prefix + "_to_right"] = np.array(bool(mps.to_right))
    for i, q in enumerate(mps.qn):
        out[f"{prefix}_qn_{i}"] = np.array(q)
    for i in range(len(mps)):
        out[f"{prefix}_sigmaqn_{i}"] = np.array(mps._get_sigmaqn(i))


def gen_holstein():
    from renormalizer.mps.gs import construct_mps_mpo, optimize_mps
    from renormalizer.mps import gs as ref_gs
    from renormalizer.tests.parameter import holstein_model
    out = {}
    procedure = [[10, 0.4], [20, 0.2], [30, 0.1], [40, 0], [40, 0]]
    np.random.seed(2026)
    mps, mpo = construct_mps_mpo(holstein_model, procedure[0][0], 1)
    dump_mp("mpo", mpo, out)
    dump_mp("mps0", mps, out)
    dump_mps_meta("mps0", mps, out)
    out["gs_zpe"] = np.array(holstein_model.gs_zpe)
    out["procedure"] = np.array(procedure, dtype=float)
    for method in ("1site", "2site"):
        m = mps.copy()
        m.optimize_config.procedure = procedure
        m.optimize_config.method = method
        # record every micro iteration energy of every sweep
        micro = []
        orig = ref_gs.single_sweep

        def spy(*a, **k):
            res = orig(*a, **k)
            micro.append(np.array([e for e, _ in res[0]]))
            return res
        ref_gs.single_sweep = spy
        try:
            np.random.seed(99)
            energies, opt = optimize_mps(m, mpo)
        finally:
            ref_gs.single_sweep = orig
        out[f"{method}_energies"] = np.array(energies)
        for i, mi in enumerate(micro):
            out[f"{method}_micro_{i}"] = mi
        out[f"{method}_nsweeps"] = np.array(len(micro))
        out[f"{method}_expectation"] = np.array(opt.expectation(mpo))
        dump_mp(f"{method}_opt", opt, out)
    # state-averaged DMRG for the three lowest states (gs.py nroots > 1, mp.py:780-838)
    for method in ("1site", "2site"):
        m = mps.copy()
        m.optimize_config.procedure = procedure
        m.optimize_config.method = method
        m.optimize_config.nroots = 3
        np.random.seed(99)
        energies, opts = optimize_mps(m, mpo)
        out[f"sa_{method}_energies"] = np.array(energies)
        out[f"sa_{method}_expectations"] = np.array([o.expectation(mpo) for o in opts])
        ov = np.array([[abs(a.conj().dot(b)) for b in opts] for a in opts])
        out[f"sa_{method}_overlaps"] = ov
    # omega targeting: (H - omega)^2 with two-layer environments (gs.py:106-111, test_gs.py:66-86)
    m = mps.copy()
    m.optimize_config.procedure = procedure
    m.optimize_config.method = "2site"
    m.optimize_config.e_atol = 1e-6
    m.optimize_config.e_rtol = 1e-6
    np.random.seed(99)
    energies, opt = optimize_mps(m, mpo, omega=0.084)
    out["omega"] = np.array(0.084)
    out["omega_energies"] = np.array(energies)
    out["omega_expectation"] = np.array(opt.expectation(mpo))
    np.savez_compressed(os.path.join(HERE, "holstein.npz"), **out)


def gen_sbm():
    from renormalizer.model import Phonon, SpinBosonModel
    from renormalizer.model.op import Op
    from renormalizer.mps import Mps, Mpo
    from renormalizer.utils import Quantity, CompressConfig, EvolveConfig, EvolveMethod, \
        CompressCriteria
    out = {}
    nphonons, ph_levels = 6, 4
    omegas = [0.5, 0.8, 1.0, 1.3, 1.7, 2.2]
    disp = [1.0, 0.8, 0.6, 0.5, 0.4, 0.3]
    ph_list = [Phonon.simple_phonon(Quantity(o), Quantity(d), ph_levels)
               for o, d in zip(omegas, disp)]
    model = SpinBosonModel(Quantity(0.3), Quantity(1.0), ph_list)
    mpo = Mpo(model)
    np.random.seed(4242)
    mps = Mps.ground_state(model, False)
    mps.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=12)
    mps.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps, adaptive=False)
    mps = mps.expand_bond_dimension(mpo, coef=1e-6, include_ex=False)
    dump_mp("mpo", mpo, out)
    dump_mp("mps0", mps, out)
    dump_mps_meta("mps0", mps, out)
    out["mps0_coeff"] = np.array(mps.coeff)
    sigma_z = Mpo(model, Op("sigma_z", "spin"))
    dump_mp("sigma_z", sigma_z, out)
    dt = 0.05
    nsteps = 6
    sz = [mps.expectation(sigma_z)]
    energy = [mps.expectation(mpo)]
    for i in range(nsteps):
        mps = mps.evolve(mpo, dt)
        sz.append(mps.expectation(sigma_z))
        energy.append(mps.expectation(mpo))
        if i == 0:
            dump_mp("mps1", mps, out)
    out["dt"] = np.array(dt)
    out["nsteps"] = np.array(nsteps)
    out["sigma_z_t"] = np.array(sz)
    out["energy_t"] = np.array(energy)
    dump_mp("mpsT", mps, out)
    out["mpsT_coeff"] = np.array(mps.coeff)
    out["mpsT_qnidx"] = np.array(mps.qnidx)
    out["mpsT_to_right"] = np.array(bool(mps.to_right))
    # two-site projector splitting (mps.py:1407) from the same start state
    np.random.seed(4242)
    mps2 = Mps.ground_state(model, False)
    mps2.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=12)
    mps2.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps2, adaptive=False)
    mps2 = mps2.expand_bond_dimension(mpo, coef=1e-6, include_ex=False)
    sz2 = [mps2.expectation(sigma_z)]
    en2 = [mps2.expectation(mpo)]
    for i in range(4):
        mps2 = mps2.evolve(mpo, dt)
        sz2.append(mps2.expectation(sigma_z))
        en2.append(mps2.expectation(mpo))
    out["ps2_nsteps"] = np.array(4)
    out["ps2_sigma_z_t"] = np.array(sz2)
    out["ps2_energy_t"] = np.array(en2)
    out["ps2_bond_dims"] = np.array(mps2.bond_dims)
    dump_mp("ps2_mpsT", mps2, out)
    np.savez_compressed(os.path.join(HERE, "sbm.npz"), **out)


def gen_stacked():
    """StackedMpo (mps/mpo.py:483-494, gs.py:113-114,226-241): mps/tests/test_gs.py:148-158
    (H + H as two stacked members) and a genuinely split Hamiltonian (on-site part + hopping)."""
    from renormalizer.mps.gs import construct_mps_mpo, optimize_mps
    from renormalizer.mps import Mpo, StackedMpo
    from renormalizer.model import Model
    from renormalizer.tests.parameter import holstein_model
    out = {}
    procedure = [[10, 0.4], [20, 0.2], [30, 0.1], [40, 0], [40, 0]]
    np.random.seed(2026)
    model = holstein_model.switch_scheme(1)
    mps, mpo = construct_mps_mpo(model, procedure[0][0], 1)
    dump_mp("mpo", mpo, out)
    dump_mp("mps0", mps, out)
    dump_mps_meta("mps0", mps, out)
    out["procedure"] = np.array(procedure, dtype=float)
    for method in ("1site", "2site"):
        m = mps.copy()
        m.optimize_config.procedure = procedure
        m.optimize_config.method = method
        np.random.seed(99)
        e1, _ = optimize_mps(m.copy(), mpo)
        np.random.seed(99)
        e2, opt2 = optimize_mps(m.copy(), StackedMpo([mpo, mpo]))
        out[f"{method}_single_energies"] = np.array(e1)
        out[f"{method}_double_energies"] = np.array(e2)
    # split Hamiltonian: terms touching one dof / terms touching several
    terms_a = [t for t in model.ham_terms if len(t.dofs) == 1]
    terms_b = [t for t in model.ham_terms if len(t.dofs) != 1]
    mpo_a = Mpo(Model(model.basis, terms_a))
    mpo_b = Mpo(Model(model.basis, terms_b))
    dump_mp("mpo_a", mpo_a, out)
    dump_mp("mpo_b", mpo_b, out)
    m = mps.copy()
    m.optimize_config.procedure = procedure
    m.optimize_config.method = "2site"
    np.random.seed(99)
    e3, opt3 = optimize_mps(m.copy(), StackedMpo([mpo_a, mpo_b]))
    out["split_energies"] = np.array(e3)
    out["split_expectation"] = np.array(opt3.expectation(mpo))
    np.savez_compressed(os.path.join(HERE, "stacked.npz"), **out)


def gen_qc():
    """Ab initio DMRG on the reference's own H6 / STO-3G FCIDUMP (mps/tests/test_gs.py:103-145,
    the small sibling of example/h2o_qc.py): two conserved quantum numbers (alpha, beta electron
    counts), M = 30, two-site sweeps.  fci_e is the reference test's acceptance value."""
    from renormalizer.mps.gs import optimize_mps
    from renormalizer.mps import Mpo, Mps
    from renormalizer.model import Model, h_qc
    out = {}
    ref_dir = "/root/reference/renormalizer/mps/tests"
    h1e, h2e, nuc = h_qc.read_fcidump(os.path.join(ref_dir, "H6.txt"), 6)
    basis, ham_terms = h_qc.qc_model(h1e, h2e)
    model = Model(basis, ham_terms)
    mpo = Mpo(model)
    # the spin-orbital integrals read_fcidump returns (h_qc.py:15-72): input of the numeric MPO builder
    # renormalizer_b200.models.qc_mpo, whose MPO must equal the reference's symbolic one as an operator
    out["h1e"], out["h2e"] = h1e, h2e
    nelec = [3, 3]
    M = 30
    procedure = [[M, 0.4], [M, 0.2], [M, 0.1], [M, 0], [M, 0], [M, 0], [M, 0]]
    np.random.seed(2023)
    mps = Mps.random(model, nelec, M, percent=1.0)
    hf = Mps.hartree_product_state(model, {i: 1 for i in range(sum(nelec))})
    mps = mps.scale(1e-8) + hf
    dump_mp("mpo", mpo, out)
    dump_mp("mps0", mps, out)
    dump_mps_meta("mps0", mps, out)
    out["procedure"] = np.array(procedure, dtype=float)
    out["fci_e"] = np.array(-3.23747673055271 - nuc)
    out["hf_e"] = np.array(mps.expectation(mpo))
    m = mps.copy()
    m.optimize_config.procedure = procedure
    m.optimize_config.method = "2site"
    np.random.seed(99)
    energies, opt = optimize_mps(m, mpo)
    out["energies"] = np.array(energies)
    out["expectation"] = np.array(opt.expectation(mpo))
    out["bond_dims"] = np.array(opt.bond_dims)
    np.savez_compressed(os.path.join(HERE, "qc_h6.npz"), **out)


def gen_exciton():
    """FMO-like exciton dynamics (example/fmo.py in miniature): HolsteinModel with a full
    long-range J matrix, one conserved exciton, TDVP-PS at fixed bond dimension -- the
    quantum-number-blocked TDVP case; plus the same evolution of a density operator (MpDm,
    ancilla index riding along: hop_expr.py:83-117) started from the maximally entangled state."""
    from renormalizer.model import Phonon, Mol, HolsteinModel
    from renormalizer.model.op import Op
    from renormalizer.mps import Mps, Mpo, MpDm
    from renormalizer.utils import Quantity, CompressConfig, EvolveConfig, EvolveMethod, \
        CompressCriteria
    out = {}
    jm = np.array([[0.31, -0.098, 0.006, -0.006],
                   [-0.098, 0.23, 0.030, 0.007],
                   [0.006, 0.030, 0.0, -0.059],
                   [-0.006, 0.007, -0.059, 0.18]])
    phs = [Phonon.simple_phonon(Quantity(o), Quantity(d), 4)
           for o, d in zip([0.12, 0.25], [1.1, 0.6])]
    mols = [Mol(Quantity(e), phs) for e in np.diag(jm)]
    model = HolsteinModel(mols, jm)
    mpo = Mpo(model)
    np.random.seed(777)
    gs = Mps.ground_state(model, False)
    mps = Mpo.onsite(model, r"a^\dagger", dof_set={0}) @ gs
    mps.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=10)
    mps.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps, adaptive=False)
    mps = mps.expand_bond_dimension(mpo, include_ex=False)
    dump_mp("mpo", mpo, out)
    dump_mp("mps0", mps, out)
    dump_mps_meta("mps0", mps, out)
    out["mps0_coeff"] = np.array(mps.coeff)
    occ_ops = [Mpo(model, Op(r"a^\dagger a", i)) for i in range(len(mols))]
    for i, o in enumerate(occ_ops):
        dump_mp(f"occ{i}", o, out)
    out["nmol"] = np.array(len(mols))
    dt, nsteps = 2.0, 5
    occ = [[mps.expectation(o) for o in occ_ops]]
    en = [mps.expectation(mpo)]
    for i in range(nsteps):
        mps = mps.evolve(mpo, dt)
        occ.append([mps.expectation(o) for o in occ_ops])
        en.append(mps.expectation(mpo))
    out["dt"] = np.array(dt)
    out["nsteps"] = np.array(nsteps)
    out["occ_t"] = np.array(occ)
    out["energy_t"] = np.array(en)
    out["bond_dims"] = np.array(mps.bond_dims)
    dump_mp("mpsT", mps, out)
    # density operator: infinite-temperature state with one exciton, real-time TDVP-PS
    np.random.seed(778)
    dm = MpDm.max_entangled_ex(model)
    dm.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=8)
    dm.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps, adaptive=False)
    dm = dm.expand_bond_dimension(mpo, include_ex=False)
    dump_mp("dm0", dm, out)
    dump_mps_meta("dm0", dm, out)
    out["dm0_coeff"] = np.array(dm.coeff)
    docc = [[dm.expectation(o) for o in occ_ops]]
    den = [dm.expectation(mpo)]
    for i in range(3):
        dm = dm.evolve(mpo, dt)
        docc.append([dm.expectation(o) for o in occ_ops])
        den.append(dm.expectation(mpo))
    out["dm_nsteps"] = np.array(3)
    out["dm_occ_t"] = np.array(docc)
    out["dm_energy_t"] = np.array(den)
    dump_mp("dmT", dm, out)
    np.savez_compressed(os.path.join(HERE, "exciton.npz"), **out)


def _exciton_model():
    from renormalizer.model import Phonon, Mol, HolsteinModel
    from renormalizer.utils import Quantity
    jm = np.array([[0.31, -0.098, 0.006, -0.006],
                   [-0.098, 0.23, 0.030, 0.007],
                   [0.006, 0.030, 0.0, -0.059],
                   [-0.006, 0.007, -0.059, 0.18]])
    phs = [Phonon.simple_phonon(Quantity(o), Quantity(d), 4)
           for o, d in zip([0.12, 0.25], [1.1, 0.6])]
    mols = [Mol(Quantity(e), phs) for e in np.diag(jm)]
    return HolsteinModel(mols, jm), len(mols)


def gen_thermal():
    """Finite temperature (BASELINE configs[3]): imaginary-time TDVP-PS of a density operator from
    the maximally entangled one-exciton state down to beta (what mps/thermalprop.py:96-98 does each
    step: MpDm.evolve(h_mpo, -i dbeta/2), here with a fixed energy offset), followed by real-time
    steps of the thermal state; and the adaptive step-size controller (mps.py:46-115) on the
    zero-temperature exciton run."""
    from renormalizer.model.op import Op
    from renormalizer.mps import Mps, Mpo, MpDm
    from renormalizer.utils import CompressConfig, EvolveConfig, EvolveMethod, CompressCriteria
    out = {}
    model, nmol = _exciton_model()
    mpo = Mpo(model)
    occ_ops = [Mpo(model, Op(r"a^\dagger a", i)) for i in range(nmol)]
    dump_mp("mpo", mpo, out)
    for i, o in enumerate(occ_ops):
        dump_mp(f"occ{i}", o, out)
    out["nmol"] = np.array(nmol)
    np.random.seed(1234)
    dm = MpDm.max_entangled_ex(model)
    dm.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=8)
    dm.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps, adaptive=False)
    dm = dm.expand_bond_dimension(mpo, include_ex=False)
    dump_mp("dm0", dm, out)
    dump_mps_meta("dm0", dm, out)
    out["dm0_coeff"] = np.array(dm.coeff)
    beta, nbeta = 8.0, 4
    dbeta = beta / nbeta
    occ = [[dm.expectation(o) for o in occ_ops]]
    en = [dm.expectation(mpo)]
    coeffs = [dm.coeff]
    for i in range(nbeta):
        dm = dm.evolve(mpo, -0.5j * dbeta)
        occ.append([dm.expectation(o) for o in occ_ops])
        en.append(dm.expectation(mpo))
        coeffs.append(dm.coeff)
    out["beta"] = np.array(beta)
    out["nbeta"] = np.array(nbeta)
    out["imag_occ"] = np.array(occ)
    out["imag_energy"] = np.array(en)
    out["imag_coeff"] = np.array(coeffs)
    dump_mp("dm_beta", dm, out)
    # real-time evolution of the thermal state
    rocc, ren = [], []
    for i in range(2):
        dm = dm.evolve(mpo, 2.0)
        rocc.append([dm.expectation(o) for o in occ_ops])
        ren.append(dm.expectation(mpo))
    out["real_occ"] = np.array(rocc)
    out["real_energy"] = np.array(ren)
    # adaptive TDVP-PS on the pure-state run
    np.random.seed(777)
    gs = Mps.ground_state(model, False)
    mps = Mpo.onsite(model, r"a^\dagger", dof_set={0}) @ gs
    mps.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=10)
    mps.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps, adaptive=True, guess_dt=1.0,
                                     adaptive_rtol=5e-4)
    mps = mps.expand_bond_dimension(mpo, include_ex=False)
    dump_mp("mps0", mps, out)
    dump_mps_meta("mps0", mps, out)
    out["mps0_coeff"] = np.array(mps.coeff)
    aocc, guess = [], []
    for i in range(3):
        mps = mps.evolve(mpo, 4.0)
        aocc.append([mps.expectation(o) for o in occ_ops])
        guess.append(mps.evolve_config.guess_dt)
    out["adaptive_occ"] = np.array(aocc)
    out["adaptive_guess_dt"] = np.array(guess)
    dump_mp("adaptive_mpsT", mps, out)
    np.savez_compressed(os.path.join(HERE, "thermal.npz"), **out)


def gen_pc():
    """Propagate-and-compress, the default integrator of Mps.evolve (mps.py:796-884: 4th-order
    Taylor expansion of the propagator, every H^k psi by Mpo.contract = apply + canonicalise +
    compress, compressed_sum of the scaled terms) on the exciton model with its conserved exciton."""
    from renormalizer.model.op import Op
    from renormalizer.mps import Mps, Mpo
    from renormalizer.utils import CompressConfig, CompressCriteria
    out = {}
    model, nmol = _exciton_model()
    mpo = Mpo(model)
    occ_ops = [Mpo(model, Op(r"a^\dagger a", i)) for i in range(nmol)]
    dump_mp("mpo", mpo, out)
    dump_mpo_meta("mpo", mpo, out)
    for i, o in enumerate(occ_ops):
        dump_mp(f"occ{i}", o, out)
    out["nmol"] = np.array(nmol)
    for tag, cfg in (("thr", CompressConfig(CompressCriteria.threshold, threshold=1e-5)),
                     ("fix", CompressConfig(CompressCriteria.fixed, max_bonddim=12))):
        np.random.seed(777)
        gs = Mps.ground_state(model, False)
        mps = Mpo.onsite(model, r"a^\dagger", dof_set={0}) @ gs
        mps.compress_config = cfg
        if tag == "thr":
            dump_mp("mps0", mps, out)
            dump_mps_meta("mps0", mps, out)
            out["mps0_coeff"] = np.array(mps.coeff)
        occ, en, dims = [], [], []
        for i in range(4):
            mps = mps.evolve(mpo, 1.0)
            occ.append([mps.expectation(o) for o in occ_ops])
            en.append(mps.expectation(mpo))
            dims.append(mps.bond_dims)
        out[f"{tag}_occ"] = np.array(occ)
        out[f"{tag}_energy"] = np.array(en)
        out[f"{tag}_bond_dims"] = np.array(dims)
        dump_mp(f"{tag}_mpsT", mps, out)
    # adaptive step control (mps.py:826-880): 5th-order expansion, error = distance of the 4th- and
    # 5th-order sums, recursion over the remaining time
    from renormalizer.utils import EvolveConfig, EvolveMethod
    np.random.seed(777)
    gs = Mps.ground_state(model, False)
    mps = Mpo.onsite(model, r"a^\dagger", dof_set={0}) @ gs
    mps.compress_config = CompressConfig(CompressCriteria.threshold, threshold=1e-5)
    mps.evolve_config = EvolveConfig(EvolveMethod.prop_and_compress, adaptive=True, guess_dt=0.4,
                                     adaptive_rtol=1e-4)
    occ, guess, dims = [], [], []
    for i in range(3):
        mps = mps.evolve(mpo, 2.0)
        occ.append([mps.expectation(o) for o in occ_ops])
        guess.append(mps.evolve_config.guess_dt)
        dims.append(mps.bond_dims)
    out["ada_occ"] = np.array(occ)
    out["ada_guess_dt"] = np.array(guess)
    out["ada_bond_dims"] = np.array(dims)
    # Runge-Kutta propagators (mps.py:664-793): classical RK4 with the MPO given as a function of time,
    # and the tableau integrator with the embedded Fehlberg pair and adaptive step control
    for tag, kw in (("rk4", dict(method=EvolveMethod.prop_and_compress_tdrk4)),
                    ("rkf", dict(method=EvolveMethod.prop_and_compress_tdrk, adaptive=True, guess_dt=0.3,
                                 adaptive_rtol=1e-4, rk_solver="RKF45")),
                    ("rk3", dict(method=EvolveMethod.prop_and_compress_tdrk, rk_solver="Kutta_RK3"))):
        np.random.seed(777)
        gs = Mps.ground_state(model, False)
        mps = Mpo.onsite(model, r"a^\dagger", dof_set={0}) @ gs
        mps.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=10)
        mps.evolve_config = EvolveConfig(**kw)
        occ, guess, dims = [], [], []
        for i in range(3):
            mps = mps.evolve((lambda t, *a, **k: mpo) if tag == "rk4" else mpo, 0.5)
            occ.append([mps.expectation(o) for o in occ_ops])
            guess.append(mps.evolve_config.guess_dt)
            dims.append(mps.bond_dims)
        out[f"{tag}_occ"] = np.array(occ)
        out[f"{tag}_energy"] = np.array(mps.expectation(mpo))
        out[f"{tag}_guess_dt"] = np.array(guess)
        out[f"{tag}_bond_dims"] = np.array(dims)
        dump_mp(f"{tag}_mpsT", mps, out)
    np.savez_compressed(os.path.join(HERE, "pc.npz"), **out)


def gen_expand():
    """expand_bond_dimension with a hint MPO and include_ex=False (mps.py:1934-2023), the preparation
    step of every TDVP-PS run in the reference's examples: a spin-boson product state (no conserved
    quantum number, coef 1e-6), the one-exciton state of the exciton model and its density operator."""
    from renormalizer.model import Phonon, SpinBosonModel
    from renormalizer.mps import Mps, Mpo, MpDm
    from renormalizer.utils import Quantity, CompressConfig, CompressCriteria
    out = {}

    def record(tag, mps, mpo, coef):
        dump_mp(f"{tag}_mpo", mpo, out)
        dump_mpo_meta(f"{tag}_mpo", mpo, out)
        dump_mp(f"{tag}_pre", mps, out)
        dump_mps_meta(f"{tag}_pre", mps, out)
        out[f"{tag}_pre_coeff"] = np.array(mps.coeff)
        out[f"{tag}_coef"] = np.array(coef)
        out[f"{tag}_max_bonddim"] = np.array(mps.compress_config.bond_dim_max_value)
        new = mps.expand_bond_dimension(mpo, coef=coef, include_ex=False)
        dump_mp(f"{tag}_post", new, out)
        dump_mps_meta(f"{tag}_post", new, out)
        out[f"{tag}_post_coeff"] = np.array(new.coeff)
        out[f"{tag}_post_bond_dims"] = np.array(new.bond_dims)
        out[f"{tag}_post_energy"] = np.array(new.expectation(mpo))

    ph_list = [Phonon.simple_phonon(Quantity(o), Quantity(d), 4)
               for o, d in zip([0.5, 0.8, 1.0, 1.3, 1.7, 2.2], [1.0, 0.8, 0.6, 0.5, 0.4, 0.3])]
    model = SpinBosonModel(Quantity(0.3), Quantity(1.0), ph_list)
    mps = Mps.ground_state(model, False)
    mps.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=12)
    record("sbm", mps, Mpo(model), 1e-6)

    model, nmol = _exciton_model()
    mpo = Mpo(model)
    mps = Mpo.onsite(model, r"a^\dagger", dof_set={0}) @ Mps.ground_state(model, False)
    mps.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=10)
    record("ex", mps, mpo, 1e-10)

    dm = MpDm.max_entangled_ex(model)
    dm.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=8)
    record("dm", dm, mpo, 1e-10)
    np.savez_compressed(os.path.join(HERE, "expand.npz"), **out)


def gen_entropy():
    """Bond singular values and von Neumann bond entropies (mps.py:1759-1793) of an evolved exciton
    state and of a thermal density operator."""
    from renormalizer.mps import Mps, Mpo, MpDm
    from renormalizer.utils import CompressConfig, EvolveConfig, EvolveMethod, CompressCriteria
    out = {}
    model, nmol = _exciton_model()
    mpo = Mpo(model)
    dump_mp("mpo", mpo, out)
    mps = Mpo.onsite(model, r"a^\dagger", dof_set={0}) @ Mps.ground_state(model, False)
    mps.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=10)
    mps.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps, adaptive=False)
    mps = mps.expand_bond_dimension(mpo, include_ex=False)
    for i in range(3):
        mps = mps.evolve(mpo, 2.0)
    dm = MpDm.max_entangled_ex(model)
    dm.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=8)
    dm.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps, adaptive=False)
    dm = dm.expand_bond_dimension(mpo, include_ex=False)
    for i in range(2):
        dm = dm.evolve(mpo, -1.0j)
    for tag, m in (("mps", mps), ("dm", dm)):
        dump_mp(tag, m, out)
        dump_mps_meta(tag, m, out)
        out[f"{tag}_singular_values"] = m.calc_bond_singular_values()
        out[f"{tag}_bond_entropy"] = m.calc_bond_entropy()
    np.savez_compressed(os.path.join(HERE, "entropy.npz"), **out)


def gen_vcompress():
    """Variational compression of mpo @ mps (mp.py:513-650, Mpo.contract(algo="variational")): the
    sweep over environments with bra = compressed guess and ket = the state, H_eff applied to the
    ket centre, SVD update of the guess -- for the next round's device implementation."""
    from renormalizer.mps import Mps, Mpo
    from renormalizer.utils import CompressConfig, EvolveConfig, EvolveMethod, CompressCriteria
    out = {}
    model, nmol = _exciton_model()
    mpo = Mpo(model)
    dump_mp("mpo", mpo, out)
    dump_mpo_meta("mpo", mpo, out)
    out["mpo_to_right"] = np.array(bool(mpo.to_right))
    for i in range(len(mpo)):
        out[f"mpo_sigmaqn_{i}"] = np.array(mpo._get_sigmaqn(i))
    mps = Mpo.onsite(model, r"a^\dagger", dof_set={0}) @ Mps.ground_state(model, False)
    mps.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=10)
    mps.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps, adaptive=False)
    mps = mps.expand_bond_dimension(mpo, include_ex=False)
    for i in range(2):
        mps = mps.evolve(mpo, 2.0)
    dump_mp("mps", mps, out)
    dump_mps_meta("mps", mps, out)
    for method in ("1site", "2site"):
        m = mps.copy()
        m.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=8, vmethod=method)
        new = mpo.contract(m, algo="variational")
        exact = mpo.apply(mps)
        out[f"{method}_bond_dims"] = np.array(new.bond_dims)
        out[f"{method}_norm"] = np.array(new.mp_norm)
        out[f"{method}_overlap_exact"] = np.array(new.conj().dot(exact))
        out[f"{method}_exact_norm"] = np.array(exact.mp_norm)
        dump_mp(f"{method}_new", new, out)
        dump_mps_meta(f"{method}_new", new, out)
    np.savez_compressed(os.path.join(HERE, "vcompress.npz"), **out)


def gen_two_spin():
    """The README quickstart (README.md:36-58): two half spins, sigma+ sigma- exchange, 10 steps
    of Mps.evolve with dt = 0.05, <Z_0> after every step -- with the default propagate-and-
    compress RK4 integrator (the README's own call) and with TDVP-PS."""
    from renormalizer import Mps, Mpo, Op, Model, BasisHalfSpin
    from renormalizer.utils import EvolveConfig, EvolveMethod
    out = {}
    basis = [BasisHalfSpin(0), BasisHalfSpin(1)]
    ham_terms = Op("sigma_+ sigma_-", [0, 1]) + Op("sigma_+ sigma_-", [1, 0])
    model = Model(basis, ham_terms)
    mpo = Mpo(model)
    z_op = Mpo(model, Op("Z", 0))
    dump_mp("mpo", mpo, out)
    dump_mpo_meta("mpo", mpo, out)
    dump_mp("z", z_op, out)
    for tag, cfg in (("pc", None), ("ps", EvolveConfig(EvolveMethod.tdvp_ps)),
                     ("ps2", EvolveConfig(EvolveMethod.tdvp_ps2))):
        mps = Mps.hartree_product_state(model, condition={0: [0, 1]})
        if cfg is not None:
            mps.evolve_config = cfg
        if tag == "pc":
            dump_mp("mps0", mps, out)
            dump_mps_meta("mps0", mps, out)
        zs = []
        for i in range(10):
            mps = mps.evolve(mpo, 0.05)
            zs.append(mps.expectation(z_op))
        out[f"{tag}_z_t"] = np.array(zs)
        out[f"{tag}_bond_dims"] = np.array(mps.bond_dims)
    np.savez_compressed(os.path.join(HERE, "two_spin.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["kernels", "svdqn", "krylov", "davidson", "holstein", "sbm", "stacked", "qc", "exciton",
                             "two_spin", "thermal", "pc", "expand", "entropy", "vcompress"]
    for name in which:
        print("generating", name, flush=True)
        globals()["gen_" + name]()
    print("done")

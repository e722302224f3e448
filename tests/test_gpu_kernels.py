"""GPU parity tests of the C-ABI kernels against the CPU oracle (run with -m gpu on a B200)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import contract as oc
from helpers import relerr

pytestmark = pytest.mark.gpu

TOL = 1e-12  # float64 contractions: relative to the largest output element


def dev(a):
    from renormalizer_b200.backend import asxp
    return asxp(a)


def host(t):
    return t.detach().cpu().numpy()


def rnd(rng, shape, cplx):
    a = rng.standard_normal(shape)
    if cplx:
        a = a + 1j * rng.standard_normal(shape)
    return a


@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (7, 5, 3), (128, 128, 16), (130, 257, 33),
                                   (300, 64, 1000), (64, 300, 17), (513, 129, 255)])
def test_dgemm_tn(m, n, k):
    from renormalizer_b200 import ops
    rng = np.random.default_rng(m * 1000 + n * 10 + k)
    a, b = rng.standard_normal((m, k)), rng.standard_normal((n, k))
    c = ops.gemm_tn(dev(a), dev(b), m, n, k, k, k)
    assert relerr(host(c), a @ b.T) < TOL
    # accumulate + leading dimensions larger than the logical sizes
    c0 = rng.standard_normal((m, n + 3))
    cd = dev(c0)
    ops.gemm_tn(dev(a), dev(b), m, n, k, k, k, out=cd, ldc=n + 3, accumulate=True)
    exp = c0.copy()
    exp[:, :n] += a @ b.T
    assert relerr(host(cd), exp) < TOL


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("M,K,N", [(5, 7, 3), (33, 65, 40), (256, 31, 100)])
def test_matmul_pack(cplx, M, K, N):
    from renormalizer_b200 import ops
    rng = np.random.default_rng(11)
    a, b = rnd(rng, (M, K), cplx), rnd(rng, (K, N), cplx)
    assert relerr(host(ops.matmul(dev(a), dev(b))), a @ b) < TOL


@pytest.mark.parametrize("t", ["r", "c"])
def test_hop_golden(golden, t):
    """H_eff.C against vectors produced by the reference's hop_expr."""
    from renormalizer_b200.hop_expr import hop_expr
    g = golden("kernels")
    L, R, R1, R0, W1, W2 = (g[f"{t}_{k}"] for k in ("L", "R", "R1", "R0", "W1", "W2"))
    cases = [("hop0", R0, [], "C0"), ("hop1", R1, [W1], "C1"), ("hop2", R, [W1, W2], "C2"),
             ("hop1a", R1, [W1], "C1a"), ("hop2a", R, [W1, W2], "C2a")]
    for name, r, cmo, ck in cases:
        c = g[f"{t}_{ck}"]
        expr = hop_expr(dev(L), dev(r), list(cmo), c.shape)
        got = host(expr(dev(c)))
        assert got.shape == g[f"{t}_{name}"].shape
        assert relerr(got, g[f"{t}_{name}"]) < TOL, name


@pytest.mark.parametrize("t", ["r", "c"])
def test_env_golden(golden, t):
    from renormalizer_b200.lib import contract_one_site
    g = golden("kernels")
    L, R1, W1 = g[f"{t}_L"], g[f"{t}_R1"], g[f"{t}_W1"]
    for name, env, a, dom in [("envL3", L, "A3", "L"), ("envL4", L, "A4", "L"),
                              ("envR3", R1, "A3", "R"), ("envR4", R1, "A4", "R")]:
        got = host(contract_one_site(dev(env), dev(g[f"{t}_{a}"]), W1, dom))
        assert relerr(got, g[f"{t}_{name}"]) < TOL, name


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("Ml,Mr,w,d", [(1, 1, 1, 2), (13, 9, 3, 4), (64, 48, 5, 8), (130, 100, 4, 3)])
def test_hop_env_random(cplx, Ml, Mr, w, d):
    """Oracle comparison at ragged sizes, including different bra/ket bond dimensions."""
    from renormalizer_b200.hop_expr import hop_expr
    from renormalizer_b200.lib import contract_one_site
    rng = np.random.default_rng(Ml * 7 + Mr)
    Ml2, Mr2 = Ml + 2, max(1, Mr - 1)
    L = rnd(rng, (Ml2, w, Ml), cplx)
    R = rnd(rng, (Mr2, w + 1, Mr), cplx)
    W = rng.standard_normal((w, d, d, w + 1)) * (rng.random((w, d, d, w + 1)) < 0.4)
    W2 = rng.standard_normal((w + 1, d, d, w + 1)) * (rng.random((w + 1, d, d, w + 1)) < 0.4)
    C1 = rnd(rng, (Ml, d, Mr), cplx)
    got = host(hop_expr(dev(L), dev(R), [W], C1.shape)(dev(C1)))
    assert relerr(got, oc.hop_apply(L, R, [W], C1)) < TOL
    C2 = rnd(rng, (Ml, d, d, Mr), cplx)
    got = host(hop_expr(dev(L), dev(R), [W, W2], C2.shape)(dev(C2)))
    assert relerr(got, oc.hop_apply(L, R, [W, W2], C2)) < TOL
    R0 = rnd(rng, (Mr2, w, Mr), cplx)
    C0 = rnd(rng, (Ml, Mr), cplx)
    got = host(hop_expr(dev(L), dev(R0), [], C0.shape)(dev(C0)))
    assert relerr(got, oc.hop_apply(L, R0, [], C0)) < TOL
    # environments with a bra different from the ket
    ket = rnd(rng, (Ml, d, Mr), cplx)
    bra = rnd(rng, (Ml2, d, Mr2), cplx)
    got = host(contract_one_site(dev(L), dev(ket), W, "L", ms_conj=dev(bra.conj())))
    assert relerr(got, oc.env_update(L, ket, W, "L", ms_conj=bra.conj())) < TOL
    Wr = rng.standard_normal((w, d, d, w + 1)) * (rng.random((w, d, d, w + 1)) < 0.4)
    got = host(contract_one_site(dev(R), dev(ket), Wr, "R", ms_conj=dev(bra.conj())))
    assert relerr(got, oc.env_update(R, ket, Wr, "R", ms_conj=bra.conj())) < TOL


def test_hop_linearity_full_size():
    """Size-independent property at the bench shape (M=256): H(a x + b y) = a H x + b H y and
    <y, H x> = conj(<x, H y>) for Hermitian environments."""
    from renormalizer_b200.hop_expr import hop_expr
    rng = np.random.default_rng(5)
    M, w, d = 256, 4, 10
    def herm_env():
        e = rnd(rng, (M, w, M), True)
        return e + e.conj().transpose(2, 1, 0)
    L, R = herm_env(), herm_env()
    W = rng.standard_normal((w, d, d, w))
    W = W + W.transpose(0, 2, 1, 3)
    expr = hop_expr(dev(L), dev(R), [W], (M, d, M))
    x, y = dev(rnd(rng, (M, d, M), True)), dev(rnd(rng, (M, d, M), True))
    hx, hy = expr(x), expr(y)
    comb = expr(0.3 * x + (0.2 - 0.7j) * y)
    ref = 0.3 * hx + (0.2 - 0.7j) * hy
    assert float((comb - ref).abs().max() / ref.abs().max()) < 1e-12
    a = torch.vdot(y.flatten(), hx.flatten())
    b = torch.vdot(x.flatten(), hy.flatten())
    assert abs(complex(a) - complex(b).conjugate()) / abs(complex(a)) < 1e-11


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("m,n", [(1, 1), (5, 3), (3, 5), (40, 40), (300, 17), (64, 200), (513, 130)])
def test_qr_lq(cplx, m, n):
    from renormalizer_b200 import ops
    rng = np.random.default_rng(m + 31 * n)
    a = rnd(rng, (m, n), cplx)
    k = min(m, n)
    q, r = (host(x) for x in ops.qr(dev(a)))
    assert q.shape == (m, k) and r.shape == (k, n)
    assert np.abs(q.conj().T @ q - np.eye(k)).max() < 1e-12
    assert relerr(q @ r, a) < 1e-12
    assert np.abs(np.tril(r, -1)).max() == 0
    if m >= n:
        # same Householder convention as LAPACK: compare with numpy directly
        qn, rn_ = np.linalg.qr(a)
        assert relerr(r, rn_) < 1e-10 and relerr(q, qn) < 1e-10
    l, q2 = (host(x) for x in ops.qr(dev(a), lq=True))
    assert l.shape == (m, k) and q2.shape == (k, n)
    assert np.abs(q2 @ q2.conj().T - np.eye(k)).max() < 1e-12
    assert relerr(l @ q2, a) < 1e-12


@pytest.mark.parametrize("cplx", [False, True])
def test_qr_rank_deficient(cplx):
    """Padded bond dimensions give exactly rank-deficient matrices; Q must stay orthonormal."""
    from renormalizer_b200 import ops
    rng = np.random.default_rng(3)
    a = rnd(rng, (60, 4), cplx) @ rnd(rng, (4, 20), cplx)
    a[:, 7] = 0
    q, r = (host(x) for x in ops.qr(dev(a)))
    assert np.abs(q.conj().T @ q - np.eye(20)).max() < 1e-12
    assert relerr(q @ r, a) < 1e-12


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("m,n", [(1, 1), (6, 4), (4, 6), (33, 33), (200, 31), (50, 120), (256, 256)])
def test_svd_jacobi(cplx, m, n):
    from renormalizer_b200 import ops
    rng = np.random.default_rng(m * 3 + n)
    a = rnd(rng, (m, n), cplx)
    # graded singular values over 12 decades: Jacobi keeps relative accuracy
    k = min(m, n)
    u0, _, v0 = np.linalg.svd(a, full_matrices=False)
    s0 = np.logspace(0, -12, k)
    a = (u0 * s0) @ v0
    u, s, vh = (host(x) for x in ops.svd(dev(a)))
    assert np.all(np.diff(s) <= 0)
    assert np.abs(s - s0).max() < 1e-13
    # forming `a` in floating point perturbs sigma_min = 1e-12 by ~1e-16 absolute already
    assert np.abs(s / s0 - 1).max() < 1e-3
    assert relerr((u * s) @ vh, a) < 1e-12
    assert np.abs(u.conj().T @ u - np.eye(k)).max() < 1e-11
    assert np.abs(vh @ vh.conj().T - np.eye(k)).max() < 1e-11


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("m,n", [(700, 300), (300, 700), (512, 512)])
def test_svd_preconditioned_graded(cplx, m, n):
    """Exponentially decaying singular values (what a converged bond matrix looks like): the
    QR-preconditioned iteration converges in a few sweeps where the bare one needs > 30, and the
    vectors of the tiny singular values stay orthonormal.  Also a rank-deficient block."""
    from renormalizer_b200 import ops
    rng = np.random.default_rng(m + 7 * n)
    k = min(m, n)
    u0, _, v0 = np.linalg.svd(rnd(rng, (m, n), cplx), full_matrices=False)
    for s0 in (np.exp(-0.08 * np.arange(k)), np.where(np.arange(k) < k // 2, 1.0 / (1 + np.arange(k)), 0.0)):
        a = (u0 * s0) @ v0
        u, s, vh = (host(x) for x in ops.svd(dev(a)))
        assert ops.svd.last_sweeps <= 14
        assert np.all(np.diff(s) <= 0)
        assert np.abs(s - s0).max() < 1e-13
        assert relerr((u * s) @ vh, a) < 1e-12
        assert np.abs(u.conj().T @ u - np.eye(k)).max() < 1e-11
        assert np.abs(vh @ vh.conj().T - np.eye(k)).max() < 1e-11
        # same decomposition as the bare iteration (singular values; the vectors up to a gauge)
        _, s_bare, _ = (host(x) for x in ops.svd(dev(a), precondition=False))
        assert np.abs(s - s_bare).max() < 1e-12


def test_vector_kernels():
    from renormalizer_b200 import ops
    rng = np.random.default_rng(0)
    n, nvec = 100003, 5
    for cplx in (False, True):
        V = rnd(rng, (nvec, n), cplx)
        x = rnd(rng, (n,), cplx)
        ws = ops.VecWorkspace(dev(x).device)
        out = host(ops.multi_dot(dev(V), dev(x), nvec, n, cplx, ws)).reshape(nvec, 2)
        ref = V.conj() @ x
        assert relerr(out[:, 0] + 1j * out[:, 1], ref) < 1e-12
        coef = rnd(rng, (nvec,), cplx)
        o = dev(np.zeros(n, dtype=V.dtype))
        ops.lincomb(dev(V), dev(coef), nvec, n, cplx, o)
        assert relerr(host(o), coef @ V) < 1e-12
        w, vj, vm = rnd(rng, (n,), cplx), rnd(rng, (n,), cplx), rnd(rng, (n,), cplx)
        wd = dev(w)
        alpha, beta = dev(np.array([0.37, 0.0])), dev(np.array([1.9, 0.0]))
        bout = dev(np.zeros(2))
        ops.lanczos_update(wd, dev(vj), dev(vm), alpha, beta, ws, bout)
        exp = w - 0.37 * vj - 1.9 * vm
        assert relerr(host(wd), exp) < 1e-14
        assert abs(host(bout)[0] - np.linalg.norm(exp)) / np.linalg.norm(exp) < 1e-13
        o2 = dev(np.zeros(n, dtype=V.dtype))
        ops.scale_inv(wd, bout, o2)
        assert relerr(host(o2), exp / np.linalg.norm(exp)) < 1e-13


def test_host_buffer_entry_points(golden):
    """The host-pointer C-ABI calls (what a NumPy-side caller binds)."""
    import ctypes
    from renormalizer_b200 import _lib
    lib = _lib.get()
    g = golden("kernels")
    for t, cplx in (("r", 0), ("c", 1)):
        L, R1, W1, C1 = (np.ascontiguousarray(g[f"{t}_{k}"]) for k in ("L", "R1", "W1", "C1"))
        out = np.zeros_like(g[f"{t}_hop1"])
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        err = lib.rn_hop_apply_host(cplx, 1, p(L), *L.shape, p(R1), *R1.shape, C1.shape[1], 1, 1, 1,
                                    p(W1), W1.shape[3], None, 0, p(C1), p(out), 0)
        assert err == 0
        assert relerr(out, g[f"{t}_hop1"]) < TOL
        A3 = np.ascontiguousarray(g[f"{t}_A3"])
        for dom, env, name in ((0, L, "envL3"), (1, R1, "envR3")):
            out = np.zeros_like(g[f"{t}_{name}"])
            Mf = Mh = A3.shape[2] if dom == 0 else A3.shape[0]
            err = lib.rn_env_update_host(cplx, dom, p(env), *env.shape, p(A3), p(A3), A3.shape[1], 1,
                                         Mf, Mh, p(W1), W1.shape[0], W1.shape[3], p(out), 0)
            assert err == 0
            assert relerr(out, g[f"{t}_{name}"]) < TOL


@pytest.mark.parametrize("cplx", [0, 1])
@pytest.mark.parametrize("m,n", [(20, 12), (90, 200), (260, 130)])
def test_svd_host_entry_point(cplx, m, n):
    """rn_svd_host: the scipy.linalg.svd(a, full_matrices=False) call of optimized_svd
    (svd_qn.py:13-49) with NumPy buffers -- sorted singular values, A = U S Vh."""
    import ctypes
    from renormalizer_b200 import _lib
    lib = _lib.get()
    rng = np.random.default_rng(m + n + cplx)
    k = min(m, n)
    u0, _, v0 = np.linalg.svd(rnd(rng, (m, n), bool(cplx)), full_matrices=False)
    s0 = np.exp(-0.1 * np.arange(k))
    a = np.ascontiguousarray((u0 * s0) @ v0)
    u = np.zeros((m, k), dtype=a.dtype)
    s = np.zeros(k)
    vh = np.zeros((k, n), dtype=a.dtype)
    p = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    assert lib.rn_svd_host(cplx, m, n, p(a), p(u), p(s), p(vh), 1) == 0
    assert np.all(np.diff(s) <= 0)
    assert np.abs(s - np.linalg.svd(a, compute_uv=False)).max() < 1e-13
    assert relerr((u * s) @ vh, a) < 1e-12
    assert np.abs(u.conj().T @ u - np.eye(k)).max() < 1e-11
    assert np.abs(vh @ vh.conj().T - np.eye(k)).max() < 1e-11


@pytest.mark.parametrize("m,n,k", [(128, 128, 128), (1, 1, 1), (100, 37, 50), (256, 384, 1000),
                                   (300, 130, 129), (768, 4096, 512), (2048, 512, 1536)])
@pytest.mark.parametrize("nslices", [7, 8])
def test_ozaki_gemm_tcgen05(m, n, k, nslices):
    """tcgen05 int8 split GEMM against NumPy FP64; rows/columns with very different scales."""
    from renormalizer_b200 import ops
    rng = np.random.default_rng(m + n + k)
    a, b = rng.standard_normal((m, k)), rng.standard_normal((n, k))
    a *= np.exp(rng.uniform(-20, 20, size=(m, 1)))
    b *= np.exp(rng.uniform(-20, 20, size=(n, 1)))
    c = host(ops.ozaki_gemm_tn(dev(a), dev(b), m, n, k, k, k, nslices=nslices))
    ref = a @ b.T
    bound = np.abs(a).max(axis=1)[:, None] * np.abs(b).max(axis=1)[None, :] * k
    err = (np.abs(c - ref) / bound).max()
    assert err < (4e-14 if nslices == 7 else 4e-16), err


def test_ozaki_gemm_is_exact_on_small_integers():
    """Integer inputs below 2^6 need a single digit: the result must be bit exact."""
    from renormalizer_b200 import ops
    rng = np.random.default_rng(0)
    m, n, k = 200, 150, 300
    a = rng.integers(-60, 60, size=(m, k)).astype(np.float64)
    b = rng.integers(-60, 60, size=(n, k)).astype(np.float64)
    c = host(ops.ozaki_gemm_tn(dev(a), dev(b), m, n, k, k, k, nslices=3))
    assert np.array_equal(c, a @ b.T)


@pytest.mark.parametrize("cplx", [False, True])
def test_hop_tensor_path_matches_fp64_path(cplx):
    """H_eff.C with path=1 (tcgen05 split GEMMs) against path=0 (FP64 DMMA) and the oracle."""
    from renormalizer_b200 import ops
    rng = np.random.default_rng(4)
    Ml, Mr, w, d = 96, 80, 4, 6
    L, R = rnd(rng, (Ml, w, Ml), cplx), rnd(rng, (Mr, w, Mr), cplx)
    W = rng.standard_normal((w, d, d, w)) * (rng.random((w, d, d, w)) < 0.5)
    C = rnd(rng, (Ml, d, Mr), cplx)
    dt = torch.complex128 if cplx else torch.float64
    ref = oc.hop_apply(L, R, [W], C)
    for path in (0, 1):
        plan = ops.HopPlan(dev(L), dev(R), [ops.MpoSite(W)], C.shape, dt, path=path)
        got = host(plan.apply(dev(C)))
        plan.close()
        assert relerr(got, ref) < (1e-12 if path == 0 else 1e-11), path


@pytest.fixture
def all_tensor_path():
    """Route every contraction GEMM, however small, through the tcgen05 split path (exercises the
    fused split kernels and split-K at ragged sizes)."""
    from renormalizer_b200 import _lib
    lib = _lib.get()
    lib.rn_set_ozaki(7, 0.0)
    yield
    lib.rn_set_ozaki(7, 4.0e6)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("Ml,Mr,w,d,g", [(1, 1, 1, 2, 1), (13, 9, 3, 4, 1), (70, 48, 5, 8, 1),
                                         (33, 21, 3, 3, 2), (130, 100, 4, 3, 1)])
def test_hop_fused_split_kernels(all_tensor_path, cplx, Ml, Mr, w, d, g):
    """0-, 1- and 2-site H_eff.C with the transposing split, the MPO-apply+split kernel and the
    B-form split of R (all on the tcgen05 path), with and without ancilla indices."""
    from renormalizer_b200 import ops
    rng = np.random.default_rng(Ml * 5 + Mr + g)
    dt = torch.complex128 if cplx else torch.float64
    L = rnd(rng, (Ml + 1, w, Ml), cplx)
    R = rnd(rng, (Mr + 2, w + 1, Mr), cplx)
    R0 = rnd(rng, (Mr + 2, w, Mr), cplx)
    W1 = rng.standard_normal((w, d, d, w + 1)) * (rng.random((w, d, d, w + 1)) < 0.4)
    W2 = rng.standard_normal((w + 1, d, d, w + 1)) * (rng.random((w + 1, d, d, w + 1)) < 0.4)
    anc = (g,) if g > 1 else ()
    cases = [([], R0, (Ml, Mr)), ([W1], R, (Ml, d) + anc + (Mr,)),
             ([W1, W2], R, (Ml, d) + anc + (d,) + anc + (Mr,))]
    for cmo, r, shape in cases:
        C = rnd(rng, shape, cplx)
        ref = oc.hop_apply(L, r, cmo, C)
        plan = ops.HopPlan(dev(L), dev(r), [ops.MpoSite(x) for x in cmo], C.shape, dt, path=1)
        got = host(plan.apply(dev(C)))
        plan.close()
        assert relerr(got, ref) < 1e-11, (len(cmo), shape)


@pytest.mark.parametrize("m,n,k,ks", [(256, 512, 1536, 12), (768, 512, 512, 4), (130, 70, 1000, 3),
                                      (128, 128, 4096, 16), (2048, 512, 1536, 2)])
def test_ozaki_gemm_split_k(m, n, k, ks, monkeypatch):
    """Split-K partition of the digit GEMMs: deterministic and equal to the unsplit product."""
    import os
    from renormalizer_b200 import ops
    rng = np.random.default_rng(m + n + k)
    a, b = rng.standard_normal((m, k)), rng.standard_normal((n, k))
    ref = a @ b.T
    bound = np.abs(a).max(axis=1)[:, None] * np.abs(b).max(axis=1)[None, :] * k
    c1 = host(ops.ozaki_gemm_tn(dev(a), dev(b), m, n, k, k, k, nslices=7))
    c2 = host(ops.ozaki_gemm_tn(dev(a), dev(b), m, n, k, k, k, nslices=7))
    assert np.array_equal(c1, c2)                      # run-to-run reproducible
    assert (np.abs(c1 - ref) / bound).max() < 4e-14


def _herm_env(rng, M, w, cplx):
    e = rnd(rng, (M, w, M), cplx)
    return e + e.conj().transpose(2, 1, 0)


@pytest.mark.parametrize("cplx_dt", [True, False])
@pytest.mark.parametrize("Ml,Mr,w,d", [(1, 1, 1, 1), (1, 2, 1, 1), (2, 3, 2, 1), (6, 5, 2, 3), (24, 20, 3, 4)])
def test_expm_krylov_plan_vs_oracle(cplx_dt, Ml, Mr, w, d):
    """rn_expm_krylov (C++ Lanczos loop, device-side tridiagonal eigen-solver and convergence
    test) against the oracle's expm_krylov on the same H_eff: same result, same step count."""
    from renormalizer_b200 import ops
    from renormalizer_b200.hop_expr import hop_expr_dtype
    from renormalizer_b200.krylov import expm_krylov
    from oracle.krylov import expm_krylov as oracle_expm
    rng = np.random.default_rng(Ml * 11 + Mr)
    L, R = _herm_env(rng, Ml, w, True), _herm_env(rng, Mr, w, True)
    W = rng.standard_normal((w, d, d, w))
    W = W + W.transpose(0, 2, 1, 3)
    scale = 1.0 / max(1.0, np.abs(L).max() * np.abs(R).max() * np.abs(W).max() * Ml * Mr * w * d)
    L = L * scale
    C = rnd(rng, (Ml, d, Mr), True)
    dt = -0.3j if cplx_dt else -0.3
    ref, jref = oracle_expm(lambda y: oc.hop_apply(L, R, [W], y.reshape(C.shape)).ravel(), dt, C.ravel())
    hop = hop_expr_dtype(dev(L), dev(R), [W], C.shape, torch.complex128)
    got, j = expm_krylov(hop, dt, dev(C).reshape(-1))
    hop.close()
    assert j == jref
    assert relerr(host(got), ref) < 1e-10


def test_expm_krylov_plan_breakdown():
    """Start vector inside a small invariant subspace: the Lanczos recurrence breaks down and the
    device-side check must stop at the reference's step."""
    from renormalizer_b200.hop_expr import hop_expr_dtype
    from renormalizer_b200.krylov import expm_krylov
    from oracle.krylov import expm_krylov as oracle_expm
    M = 12
    L = np.zeros((M, 1, M), dtype=complex)
    L[np.arange(M), 0, np.arange(M)] = np.arange(1, M + 1)
    R = np.ones((1, 1, 1), dtype=complex)
    C = np.zeros((M, 1), dtype=complex)
    C[2, 0], C[5, 0], C[7, 0] = 1.0, -2.0, 0.5j
    ref, jref = oracle_expm(lambda y: oc.hop_apply(L, R, [], y.reshape(C.shape)).ravel(), -0.2j, C.ravel())
    hop = hop_expr_dtype(dev(L), dev(R), [], C.shape, torch.complex128)
    got, j = expm_krylov(hop, -0.2j, dev(C).reshape(-1))
    hop.close()
    assert j == jref
    assert relerr(host(got), ref) < 1e-10


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("m,n", [(2048, 256), (256, 2048), (256, 256), (1000, 97), (96, 96), (33, 31)])
def test_qr_panel_bench_shapes(cplx, m, n):
    """Blocked cluster-panel Householder QR / LQ at the sweep's bond-matrix shapes: orthonormal
    factor, reconstruction, triangular factor, and LAPACK's reflector convention (tall case)."""
    from renormalizer_b200 import ops
    rng = np.random.default_rng(m + 7 * n)
    a = rnd(rng, (m, n), cplx)
    k = min(m, n)
    q, r = (host(x) for x in ops.qr(dev(a)))
    assert np.abs(q.conj().T @ q - np.eye(k)).max() < 1e-12
    assert relerr(q @ r, a) < 1e-12
    assert np.abs(np.tril(r, -1)).max() == 0
    if m >= n:
        qn, rn_ = np.linalg.qr(a)
        assert relerr(r, rn_) < 1e-10 and relerr(q, qn) < 1e-10
    l, q2 = (host(x) for x in ops.qr(dev(a), lq=True))
    assert np.abs(q2 @ q2.conj().T - np.eye(k)).max() < 1e-12
    assert relerr(l @ q2, a) < 1e-12
    assert np.abs(np.triu(l, 1)).max() == 0


def test_qr_graded_columns():
    """Columns spanning 14 decades (the conditioning of an evolved TDVP site tensor): Householder
    QR must stay orthonormal and reconstruct to round-off."""
    from renormalizer_b200 import ops
    rng = np.random.default_rng(9)
    m, n = 700, 150
    a = rnd(rng, (m, n), True) * np.logspace(0, -14, n)[None, :]
    a[:, 40] = a[:, 3] * (0.3 - 2j)            # exactly dependent column
    q, r = (host(x) for x in ops.qr(dev(a)))
    assert np.abs(q.conj().T @ q - np.eye(n)).max() < 1e-12
    assert relerr(q @ r, a) < 1e-12


@pytest.mark.parametrize("t", ["r", "c"])
def test_hop_two_layer_golden(golden, t):
    """The (H - omega)^2 expressions of hop_expr.py:24-52 (4-index environments, two MPO layers)
    against vectors produced by the reference's hop_expr(twolayer=True)."""
    from renormalizer_b200.hop_expr import hop_expr
    g = golden("kernels")
    L4, R41, R42, W1, W2 = (g[f"{t}_{k}"] for k in ("L4", "R41", "R42", "W1", "W2"))
    for name, r, cmo, ck in [("hop1_2l", R41, [W1], "C1"), ("hop2_2l", R42, [W1, W2], "C2")]:
        c = g[f"{t}_{ck}"]
        got = host(hop_expr(dev(L4), dev(r), list(cmo), c.shape, twolayer=True)(dev(c)))
        assert got.shape == g[f"{t}_{name}"].shape
        assert relerr(got, g[f"{t}_{name}"]) < TOL, name


def test_qr_fallback_path_subprocess():
    """The launch-per-reflector fallback (used when a panel does not fit on chip) still factors
    correctly; it is selected with RN_QR_PANEL=0, read once per process."""
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np, torch, sys\n"
        "sys.path.insert(0, '.')\n"
        "from renormalizer_b200 import ops\n"
        "from renormalizer_b200.backend import asxp\n"
        "rng = np.random.default_rng(1)\n"
        "a = rng.standard_normal((300, 70)) + 1j * rng.standard_normal((300, 70))\n"
        "q, r = (x.cpu().numpy() for x in ops.qr(asxp(a)))\n"
        "assert np.abs(q.conj().T @ q - np.eye(70)).max() < 1e-12\n"
        "assert np.abs(q @ r - a).max() < 1e-11\n"
        "print('fallback ok')\n")
    env = dict(os.environ, RN_QR_PANEL="0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "fallback ok" in out.stdout, out.stderr[-2000:]


def test_ozaki_gemm_cta_pair_subprocess():
    """The opt-in cta_group::2 variant of the digit GEMM (two CTAs per 256 x 128 tile, RN_OZ_CTA2=1,
    read once per process) gives the same product, with and without split-K."""
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np, torch, sys\n"
        "sys.path.insert(0, '.')\n"
        "from renormalizer_b200 import ops\n"
        "from renormalizer_b200.backend import asxp\n"
        "rng = np.random.default_rng(2)\n"
        "for (m, n, k) in [(300, 130, 129), (768, 4096, 512), (2048, 512, 1536), (1000, 257, 4000)]:\n"
        "    a, b = rng.standard_normal((m, k)), rng.standard_normal((n, k))\n"
        "    c = ops.ozaki_gemm_tn(asxp(a), asxp(b), m, n, k, k, k, nslices=7).cpu().numpy()\n"
        "    bound = np.abs(a).max(axis=1)[:, None] * np.abs(b).max(axis=1)[None, :] * k\n"
        "    assert (np.abs(c - a @ b.T) / bound).max() < 4e-14, (m, n, k)\n"
        "print('pair ok')\n")
    env = dict(os.environ, RN_OZ_CTA2="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "pair ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("shape", [((1, 16), (16, 1)), ((9, 16), (16, 7)), ((40, 40), (40, 40))])
def test_ops_take_lazy_conj_views(shape):
    """Regression (round 1, variational compression): torch's `.conj()` is a lazy view (conj bit) and
    `.contiguous()` keeps it on an already contiguous tensor, so a kernel reading data_ptr() saw the
    UN-conjugated data.  Every op must materialise such views."""
    from renormalizer_b200 import ops
    rng = np.random.default_rng(5)
    a, b = rnd(rng, shape[0], True), rnd(rng, shape[1], True)
    ta, tb = dev(a), dev(b)
    assert relerr(host(ops.matmul(ta.conj(), tb)), a.conj() @ b) < TOL
    assert relerr(host(ops.matmul(ta, tb.conj())), a @ b.conj()) < TOL
    # conj().transpose().contiguous() of a single-row / single-column tensor is still a lazy view
    at = ta.conj().transpose(0, 1).contiguous()
    assert relerr(host(ops.matmul(at, dev(a))), a.conj().T @ a) < TOL
    q, r = ops.qr(dev(a.T.copy()).conj())
    assert relerr(host(q) @ host(r), a.T.conj()) < 1e-12
    u, s, vh = ops.svd(ta.conj())
    assert relerr((host(u) * host(s)) @ host(vh), a.conj()) < 1e-12


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("m,n,r", [(16, 2, 1), (24, 24, 9), (40, 64, 13), (200, 96, 30), (96, 300, 50), (130, 130, 0)])
def test_block_svd_rank_deficient_orthonormal(cplx, m, n, r):
    """Rank-deficient blocks (zero-padded bonds, bond dimension above the rank): the vectors one-sided
    Jacobi returns for (numerically) zero singular values are replaced by an orthonormal completion;
    U and V must be orthonormal and reproduce the block (svd_qn.py:13-66 semantics)."""
    from renormalizer_b200.svd_qn import _block_svd
    rng = np.random.default_rng(m * 7 + n)
    a = rnd(rng, (m, r), cplx) @ rnd(rng, (r, n), cplx) if r else np.zeros((m, n), dtype=complex if cplx else float)
    for full in (False, True):
        np.random.seed(3)
        u, s, vh = _block_svd(dev(a), full, True)
        u, s, vh = host(u), host(s), host(vh)
        k = min(m, n)
        assert np.abs(u.conj().T @ u - np.eye(u.shape[1])).max() < 1e-12
        assert np.abs(vh @ vh.conj().T - np.eye(vh.shape[0])).max() < 1e-12
        assert np.abs((u[:, :k] * s) @ vh[:k] - a).max() < 1e-12 * max(1.0, np.abs(a).max())
        if r:
            assert np.abs(s[:r] - np.linalg.svd(a, compute_uv=False)[:r]).max() < 1e-12 * s[0]
        assert np.all(s[r:] < 1e-12 * max(s[0], 1e-300)) or r == 0


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("domain", ["L", "R"])
@pytest.mark.parametrize("pad", [0, 5])
def test_env_update_bra_differs_from_ket_zero_padded(cplx, domain, pad):
    """contract_one_site with ms_conj != conj(ms) (variational compression, mp.py:600-607): bra and
    ket of different bond dimensions, embedded in zero-padded square tensors as
    Mps.variational_compress does; rectangular operands without padding as well."""
    from renormalizer_b200.lib import contract_one_site
    rng = np.random.default_rng(17)
    w, d = 3, 4
    ea, ec, mf, mh = 6, 9, 7, 11                     # bra / ket bonds before and after the site
    big_in, big_out = (max(ea, ec) + pad, max(mf, mh) + pad) if pad else (None, None)
    env = rnd(rng, (ea, w, ec), cplx)
    mo = rng.standard_normal((w, d, d, w)) * (rng.random((w, d, d, w)) < 0.6)
    if domain == "L":
        ket, bra = rnd(rng, (ec, d, mh), cplx), rnd(rng, (ea, d, mf), cplx)
    else:
        ket, bra = rnd(rng, (mh, d, ec), cplx), rnd(rng, (mf, d, ea), cplx)

    def padded(t, shape):
        out = np.zeros(shape, dtype=t.dtype)
        out[tuple(slice(0, s) for s in t.shape)] = t
        return out
    if pad:
        env = padded(env, (big_in, w, big_in))
        shp = (big_in, d, big_out) if domain == "L" else (big_out, d, big_in)
        ket, bra = padded(ket, shp), padded(bra, shp)
    ref = oc.env_update(env, ket, mo, domain, ms_conj=bra.conj())
    got = host(contract_one_site(dev(env), dev(ket), mo, domain, ms_conj=dev(bra).conj()))
    assert got.shape == ref.shape
    assert relerr(got, ref) < TOL


@pytest.mark.parametrize("cplx", [False, True])
def test_wide_mpo_bond_hop_and_environment(cplx):
    """Wide MPO bonds (ab initio Hamiltonians, D*F >= 128): the MPO application runs on
    wapply_wide_kernel (one thread per output row of a 16-wide y tile) -- H_eff.C for one and two
    sites and both environment updates against the oracle."""
    from renormalizer_b200.hop_expr import hop_expr
    from renormalizer_b200.lib import contract_one_site
    rng = np.random.default_rng(23)
    M, w, d = 40, 70, 2
    L, R = rnd(rng, (M, w, M), cplx), rnd(rng, (M, w, M), cplx)
    W1 = rng.standard_normal((w, d, d, w)) * (rng.random((w, d, d, w)) < 0.1)
    W2 = rng.standard_normal((w, d, d, w)) * (rng.random((w, d, d, w)) < 0.1)
    C1, C2 = rnd(rng, (M, d, M), cplx), rnd(rng, (M, d, d, M), cplx)
    got = host(hop_expr(dev(L), dev(R), [W1], C1.shape)(dev(C1)))
    assert relerr(got, oc.hop_apply(L, R, [W1], C1)) < 1e-11
    got = host(hop_expr(dev(L), dev(R), [W1, W2], C2.shape)(dev(C2)))
    assert relerr(got, oc.hop_apply(L, R, [W1, W2], C2)) < 1e-11
    # density-operator form (ancilla index rides along)
    Ca = rnd(rng, (M, d, 3, M), cplx)
    got = host(hop_expr(dev(L), dev(R), [W1], Ca.shape)(dev(Ca)))
    assert relerr(got, oc.hop_apply(L, R, [W1], Ca)) < 1e-11
    for domain in ("L", "R"):
        got = host(contract_one_site(dev(L), dev(C1), W1, domain))
        assert relerr(got, oc.env_update(L, C1, W1, domain)) < 1e-11


def test_ozaki_gemm_long_contraction():
    """Contractions longer than a CTA's int32 accumulator range (K > 65536: the G3 of a wide MPO bond
    has K = w M) run split-K with at least K / 65536 splits whose partial tiles are summed in FP64."""
    from renormalizer_b200 import ops
    rng = np.random.default_rng(4)
    for (m, n, k) in [(130, 140, 70000), (64, 256, 200000)]:
        a, b = rng.standard_normal((m, k)), rng.standard_normal((n, k))
        c = host(ops.ozaki_gemm_tn(dev(a), dev(b), m, n, k, k, k, nslices=7))
        bound = np.abs(a).max(axis=1)[:, None] * np.abs(b).max(axis=1)[None, :] * k
        assert (np.abs(c - a @ b.T) / bound).max() < 4e-14, (m, n, k)

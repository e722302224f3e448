"""Oracle: effective-Hamiltonian application and environment update (NumPy tensordot chains).

Index conventions follow the reference: environments are (bra bond, MPO bond, ket bond);
MPO sites are (left bond, up/bra physical, down/ket physical, right bond); MPS sites are
(left bond, physical[, ancilla], right bond).
"""
import numpy as np


def hop_apply(ltensor, rtensor, cmo, c):
    """H_eff . C for 0/1/2 centre sites, with or without ancilla indices.

    Reference: renormalizer/mps/hop_expr.py:7-117 (single-layer expressions)
      0 site : "abc, lbk, ck -> al"
      1 site : "abc, bdef, lfk, cek -> adl"         (ancilla: "cegk -> adgl")
      2 sites: "abc, bdef, fghj, ljk, cehk -> adgl" (ancilla: "cemhnk -> admgnl")
    """
    nsite = len(cmo)
    if nsite == 0:
        t = np.tensordot(ltensor, c, axes=(2, 0))            # a b k
        return np.tensordot(t, rtensor, axes=([1, 2], [1, 2]))  # a l
    ancilla = 2 * nsite + 2 == c.ndim
    if not ancilla:
        assert nsite + 2 == c.ndim
    if nsite == 1:
        w = cmo[0]
        if not ancilla:
            t = np.tensordot(ltensor, c, axes=(2, 0))                # a b e k
            t = np.tensordot(t, w, axes=([1, 2], [0, 2]))            # a k d f
            return np.tensordot(t, rtensor, axes=([3, 1], [1, 2]))   # a d l
        t = np.tensordot(ltensor, c, axes=(2, 0))                    # a b e g k
        t = np.tensordot(t, w, axes=([1, 2], [0, 2]))                # a g k d f
        t = np.tensordot(t, rtensor, axes=([4, 2], [1, 2]))          # a g d l
        return t.transpose(0, 2, 1, 3)                               # a d g l
    assert nsite == 2
    w1, w2 = cmo
    if not ancilla:
        t = np.tensordot(ltensor, c, axes=(2, 0))                    # a b e h k
        t = np.tensordot(t, w1, axes=([1, 2], [0, 2]))               # a h k d f
        t = np.tensordot(t, w2, axes=([4, 1], [0, 2]))               # a k d g j
        return np.tensordot(t, rtensor, axes=([4, 1], [1, 2]))       # a d g l
    t = np.tensordot(ltensor, c, axes=(2, 0))                        # a b e m h n k
    t = np.tensordot(t, w1, axes=([1, 2], [0, 2]))                   # a m h n k d f
    t = np.tensordot(t, w2, axes=([6, 2], [0, 2]))                   # a m n k d g j
    t = np.tensordot(t, rtensor, axes=([6, 3], [1, 2]))              # a m n d g l
    return t.transpose(0, 3, 1, 4, 2, 5)                             # a d m g n l


def hop_apply_two_layer(ltensor, rtensor, cmo, c):
    """(H - omega)^2-type H_eff with two MPO layers and 4-index environments.

    Reference: renormalizer/mps/hop_expr.py:24-52
      1 site : "abcd, befg, cfhi, jgik, aej -> dhk"
      2 sites: "abcd, befg, cfhi, gjkl, ikmn, olnp, aejo -> dhmp"
    """
    if len(cmo) == 1:
        w = cmo[0]
        t = np.tensordot(ltensor, c, axes=(0, 0))                    # b c d e j
        t = np.tensordot(t, w, axes=([0, 3], [0, 1]))                # c d j f g
        t = np.tensordot(t, w, axes=([0, 3], [0, 1]))                # d j g h i
        return np.tensordot(t, rtensor, axes=([1, 2, 4], [0, 1, 2])).reshape(
            ltensor.shape[3], w.shape[2], rtensor.shape[3])          # d h k
    assert len(cmo) == 2
    w1, w2 = cmo
    t = np.tensordot(ltensor, c, axes=(0, 0))                        # b c d e j o
    t = np.tensordot(t, w1, axes=([0, 3], [0, 1]))                   # c d j o f g
    t = np.tensordot(t, w1, axes=([0, 4], [0, 1]))                   # d j o g h i
    t = np.tensordot(t, w2, axes=([3, 1], [0, 1]))                   # d o h i k l
    t = np.tensordot(t, w2, axes=([3, 4], [0, 1]))                   # d o h l m n
    return np.tensordot(t, rtensor, axes=([1, 3, 5], [0, 1, 2]))     # d h m p


def hop_diag(ltensor, rtensor, cmo):
    """Diagonal of H_eff (Davidson preconditioner).

    Reference: renormalizer/mps/gs.py:422-445 (omega is None branch).
    """
    dl = np.einsum("aba->ba", ltensor)
    dr = np.einsum("aba->ba", rtensor)
    d0 = np.einsum("abbc->abc", cmo[0])
    if len(cmo) == 1:
        t = np.tensordot(dl, d0, axes=(0, 0))          # a c g
        return np.tensordot(t, dr, axes=(2, 0))        # a c f
    d1 = np.einsum("abbc->abc", cmo[1])
    t = np.tensordot(dl, d0, axes=(0, 0))              # a c e
    u = np.tensordot(d1, dr, axes=(2, 0))              # e d f
    return np.tensordot(t, u, axes=(2, 0))             # a c d f


def env_update(environ, ms, mo, domain, ms_conj=None):
    """Absorb one site into a left ("L") or right ("R") environment.

    Reference: renormalizer/mps/lib.py:172-262 (contract_one_site).
      L, MPS : "abc, adf -> bcdf" ; "bcdf, bdeg -> cfeg" ; "cfeg, ceh -> fgh"
      R, MPS : "fda, abc -> fdbc" ; "fdbc, gdeb -> fcge" ; "fcge, hec -> fgh"
    (MPDM sites carry one extra ancilla index that is traced between bra and ket.)
    """
    if ms_conj is None:
        ms_conj = ms.conj()
    if domain == "L":
        if ms.ndim == 3:
            t = np.tensordot(environ, ms_conj, axes=(0, 0))            # b c d f
            t = np.tensordot(t, mo, axes=([0, 2], [0, 1]))             # c f e g
            return np.tensordot(t, ms, axes=([0, 2], [0, 1]))          # f g h
        t = np.tensordot(environ, ms_conj, axes=(0, 0))                # b c d l f
        t = np.tensordot(t, mo, axes=([0, 2], [0, 1]))                 # c l f e g
        return np.tensordot(t, ms, axes=([0, 3, 1], [0, 1, 2]))        # f g h
    assert domain == "R"
    if ms.ndim == 3:
        t = np.tensordot(ms_conj, environ, axes=(2, 0))                # f d b c
        t = np.tensordot(t, mo, axes=([1, 2], [1, 3]))                 # f c g e
        return np.tensordot(t, ms, axes=([3, 1], [1, 2]))              # f g h
    t = np.tensordot(ms_conj, environ, axes=(3, 0))                    # f d l b c
    t = np.tensordot(t, mo, axes=([1, 3], [1, 3]))                     # f l c g e
    return np.tensordot(t, ms, axes=([4, 1, 2], [1, 2, 3]))            # f g h

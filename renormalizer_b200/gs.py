"""DMRG ground-state sweeps on the device -- mirror of renormalizer/mps/gs.py:54-576 for the
accelerated configuration (single root, no omega targeting, Davidson or direct eigensolver)."""
import logging

import numpy as np
import scipy.linalg
import torch

from . import ops
from .backend import asnumpy, asxp
from .configs import CompressConfig, CompressCriteria
from .davidson import davidson
from .hop_expr import hop_expr_dtype
from .lib import Environ
from .svd_qn import get_qn_mask, qn_mask_outer

logger = logging.getLogger(__name__)


def optimize_mps(mps, mpo, omega: float = None):
    """DMRG ground state algorithm (gs.py:54-171).  `mps` is overwritten during the optimisation;
    returns (energies of each macro sweep, optimised mps)."""
    if omega is not None:
        # gs.py:106-111: the variational function (H - omega)^2.  The reference keeps two MPO layers
        # in 4-index environments; here the same operator is ONE MPO whose bonds are the merged
        # pairs (mpo.squared), so environments, H_eff and the sweep run on the one-layer kernels.
        from .mpo import Mpo, StackedMpo
        if isinstance(mpo, StackedMpo):
            raise NotImplementedError("StackedMPO + omega is not implemented yet")     # as gs.py:107-108
        shifted = mpo.add(Mpo.identity_like(mpo).scale(-omega))
        mpo = shifted.squared()
    assert mps.optimize_config.method in ["2site", "1site"]
    if mps.is_left_canonical:
        mps.ensure_right_canonical()
        env = "R"
    else:
        mps.ensure_left_canonical()
        env = "L"
    compress_config_bk = mps.compress_config
    from .mpo import StackedMpo as _Stacked
    if isinstance(mpo, _Stacked):
        environ = [Environ(mps, item, env) for item in mpo.mpos]      # gs.py:113-114
    else:
        environ = Environ(mps, mpo, env)
    macro_iteration_result = []
    opt_e_idx = None
    res_mps = None
    for isweep, (compress_config, percent) in enumerate(mps.optimize_config.procedure):
        if isinstance(compress_config, CompressConfig):
            mps.compress_config = compress_config
        elif isinstance(compress_config, (int, np.integer)):
            mps.compress_config = CompressConfig(criteria=CompressCriteria.fixed,
                                                 max_bonddim=int(compress_config))
        else:
            assert False
        micro_iteration_result, res_mps, mpo = single_sweep(mps, mpo, environ, None, percent, opt_e_idx)
        opt_e = min(micro_iteration_result)
        macro_iteration_result.append(opt_e[0])
        opt_e_idx = opt_e[1]
        if isweep > 0 and percent == 0:
            v1, v2 = sorted(macro_iteration_result)[:2]
            if np.allclose(v1, v2, rtol=mps.optimize_config.e_rtol, atol=mps.optimize_config.e_atol):
                logger.info("DMRG has converged!")
                break
    else:
        logger.warning("DMRG did not converge! Please increase the procedure!")
    assert res_mps is not None
    if mps.optimize_config.nroots == 1:
        res_mps = res_mps.normalize("mps_only").ensure_left_canonical().canonicalise()
        res_mps.compress_config = compress_config_bk
    else:
        res_mps = [mp.normalize("mps_only").ensure_left_canonical().canonicalise() for mp in res_mps]
        for res in res_mps:
            res.compress_config = compress_config_bk
    return macro_iteration_result, res_mps


def single_sweep(mps, mpo, environ, omega, percent, last_opt_e_idx, site_filter=None, site_hook=None):
    """gs.py:174-304.  Measurement aids (not in the reference, used by bench.py only): `site_filter`, a
    set of site indices -- the sweep optimises only those and moves the canonical centre over the
    others with a QR and the environment update; `site_hook(stage, imps, info)` is called before
    ("pre": cidx, ltensor, rtensor, mps) and after ("done": e, nhop) every optimised site update."""
    method = mps.optimize_config.method
    nroots = mps.optimize_config.nroots
    # in a state-averaged calculation: the rotated centre tensors of every state (better guesses)
    averaged_ms = []
    res_mps = None
    micro_iteration_result = []
    hop_counts = []
    for imps in mps.iter_idx_list(full=True):
        if method == "2site" and ((mps.to_right and imps == mps.site_num - 1)
                                  or ((not mps.to_right) and imps == 0)):
            break
        if mps.to_right:
            lmethod, rmethod = "System", "Enviro"
        else:
            lmethod, rmethod = "Enviro", "System"
        if method == "1site":
            lidx, cidx, ridx = imps - 1, [imps], imps + 1
        elif mps.to_right:
            lidx, cidx, ridx = imps - 1, [imps, imps + 1], imps + 2
        else:
            lidx, cidx, ridx = imps - 2, [imps - 1, imps], imps + 1
        stacked = isinstance(environ, list)
        if stacked:      # gs.py:226-228, 239-240: one environment / centre MPO list per member
            ltensor = [env_i.GetLR("L", lidx, mps, op_i, itensor=None, method=lmethod)
                       for env_i, op_i in zip(environ, mpo.mpos)]
            rtensor = [env_i.GetLR("R", ridx, mps, op_i, itensor=None, method=rmethod)
                       for env_i, op_i in zip(environ, mpo.mpos)]
            cmo = [[op_i[idx] for idx in cidx] for op_i in mpo.mpos]
        else:
            ltensor = environ.GetLR("L", lidx, mps, mpo, itensor=None, method=lmethod)
            rtensor = environ.GetLR("R", ridx, mps, mpo, itensor=None, method=rmethod)
            cmo = [mpo[idx] for idx in cidx]
        if site_filter is not None and imps not in site_filter:
            mps._push_cano(imps)
            continue
        if site_hook is not None:
            site_hook("pre", imps, dict(cidx=cidx, ltensor=ltensor, rtensor=rtensor, mps=mps, cmo=cmo))
        qnbigl, qnbigr, _ = mps._get_big_qn(cidx, need_mat=False)
        qn_mask = qn_mask_outer(qnbigl, qnbigr, mps.qntot)
        cshape = qn_mask.shape
        use_direct_eigh = np.prod(cshape) < 1000 or mps.optimize_config.algo == "direct"
        if use_direct_eigh:
            e, cstruct = eigh_direct(mps, qn_mask, ltensor, rtensor, cmo)
            nhop = 0
        else:
            if nroots == 1:
                if method == "1site":
                    raw_cguess = [mps[cidx[0]]]
                else:
                    raw_cguess = [ops.tensordot1(mps[cidx[0]], mps[cidx[1]])]
            else:
                raw_cguess = []
                for ms in averaged_ms:
                    if method == "1site":
                        raw_cguess.append(ms)
                    elif mps.to_right:
                        raw_cguess.append(ops.tensordot1(ms, mps[cidx[1]]))
                    else:
                        raw_cguess.append(ops.tensordot1(mps[cidx[0]], ms))
            e, cstruct, nhop = eigh_iterative(mps, qn_mask, ltensor, rtensor, cmo, raw_cguess)
        hop_counts.append(nhop)
        if nroots > 1:
            e = np.asarray(e).tolist()
        micro_iteration_result.append((e, cidx))
        if cidx == last_opt_e_idx:
            if nroots == 1:
                res_mps = mps.copy()
                res_mps._update_mps(cstruct, cidx, qnbigl, qnbigr, percent)
            else:
                res_mps = [mps.copy() for _ in cstruct]
                for r, ci in zip(res_mps, cstruct):
                    r._update_mps(ci, cidx, qnbigl, qnbigr, percent)
        averaged_ms = mps._update_mps(cstruct, cidx, qnbigl, qnbigr, percent)
        if site_hook is not None:
            site_hook("done", imps, dict(e=e, nhop=nhop))
    mps._switch_direction()
    mps.hop_counts = hop_counts
    return micro_iteration_result, res_mps, mpo


def _sign_fix(c):
    """gs.py:372-380 on a device vector: make the largest-magnitude component positive."""
    idx = torch.argmax(c.abs())
    return c / torch.sgn(c[idx])


def eigh_direct(mps, qn_mask, ltensor, rtensor, cmo):
    """gs.py:307-407: tiny centre tensors are diagonalised densely.  The dense H_eff is assembled
    by applying the device H_eff to the unit vectors of the symmetry-allowed subspace."""
    cshape = qn_mask.shape
    if not isinstance(ltensor, list):
        ltensor, rtensor, cmo = [ltensor], [rtensor], [cmo]
    dtype = torch.complex128 if any(t.is_complex() for t in ltensor + rtensor) else torch.float64
    hops = [hop_expr_dtype(l, r, c, cshape, dtype) for l, r, c in zip(ltensor, rtensor, cmo)]
    idx = np.nonzero(qn_mask.reshape(-1))[0]
    nfull = int(np.prod(cshape))
    dev = ltensor[0].device
    cols = []
    sel = torch.from_numpy(idx).to(dev)
    for i in idx:
        x = torch.zeros(nfull, dtype=dtype, device=dev)
        x[i] = 1
        y = hops[0](x.reshape(cshape))
        for h in hops[1:]:
            y = y + h(x.reshape(cshape))
        cols.append(y.reshape(-1).index_select(0, sel))
    for h in hops:
        h.close()
    ham = asnumpy(torch.stack(cols, dim=1)) * mps.optimize_config.inverse
    w, v = scipy.linalg.eigh(ham)

    def scatter(c):
        c = c / np.sign(c[np.abs(c).argmax()])
        cstruct = np.zeros(nfull, dtype=c.dtype)
        cstruct[idx] = c
        return asxp(cstruct.reshape(cshape))
    nroots = mps.optimize_config.nroots
    if nroots == 1:
        return w[0], scatter(v[:, 0])
    return w[:nroots], [scatter(v[:, i]) for i in range(min(nroots, v.shape[1]))]


def _hdiag(ltensor, rtensor, cmo):
    """Diagonal of H_eff for the Davidson preconditioner (gs.py:422-445).  O(M^2 d w^2) setup
    work done with torch.einsum on the device."""
    dl = torch.einsum("aba->ba", ltensor)
    dr = torch.einsum("aba->ba", rtensor)
    d0 = torch.einsum("abbc->abc", cmo[0].dense).to(dl.dtype)
    if len(cmo) == 1:
        return torch.einsum("ba,bcg,gf->acf", dl, d0, dr)
    d1 = torch.einsum("abbc->abc", cmo[1].dense).to(dl.dtype)
    return torch.einsum("ba,bce,edg,gf->acdf", dl, d0, d1, dr)


def eigh_iterative(mps, qn_mask, ltensor, rtensor, cmo, raw_cguess):
    """gs.py:486-576 with algo == "davidson".  The Davidson vectors are kept dense on the device
    and multiplied by the quantum-number mask, which is the same subspace iteration as the
    reference's gather / scatter (cvec2cmat) through the mask."""
    inverse = mps.optimize_config.inverse
    if mps.optimize_config.algo != "davidson":
        raise NotImplementedError("only the Davidson eigensolver is accelerated")
    cshape = qn_mask.shape
    if not isinstance(ltensor, list):
        ltensor, rtensor, cmo = [ltensor], [rtensor], [cmo]
    cplx = any(t.is_complex() for t in ltensor + rtensor) or any(g.is_complex() for g in raw_cguess)
    dtype = torch.complex128 if cplx else torch.float64
    dev = ltensor[0].device
    mask = torch.from_numpy(qn_mask.reshape(-1)).to(dev)
    maskf = mask.to(dtype)
    # a stacked Hamiltonian is the sum of its members' effective Hamiltonians (gs.py:499-502)
    hdiag = sum(_hdiag(l, r, c).real.reshape(-1) for l, r, c in zip(ltensor, rtensor, cmo)) * inverse
    hops = [hop_expr_dtype(l, r, c, cshape, dtype) for l, r, c in zip(ltensor, rtensor, cmo)]
    count = [0]

    class _Sum:
        def close(self):
            for h in hops:
                h.close()
    hop = _Sum()

    def aop(x):
        count[0] += 1
        y = hops[0](x.reshape(cshape)).reshape(-1)
        for h in hops[1:]:
            y = y + h(x.reshape(cshape)).reshape(-1)
        y *= maskf
        if inverse != 1.0:
            y *= inverse
        return y

    # 1/(hdiag - e + 1e-4) restricted to the allowed subspace
    def precond(x, e, *args):
        return torch.where(mask, x / (hdiag - e + 1e-4).to(dtype), torch.zeros((), dtype=dtype, device=dev))

    nroots = mps.optimize_config.nroots
    guesses = [g.reshape(-1).to(dtype) * maskf for g in raw_cguess]
    # missing guesses are random in the allowed subspace (gs.py:268-271, same np.random stream)
    for _ in range(len(guesses), nroots):
        allowed = np.nonzero(qn_mask.reshape(-1))[0]
        r = np.zeros(mask.numel())
        r[allowed] = np.random.rand(len(allowed)) - 0.5
        guesses.append(asxp(r).to(dtype))
    if nroots == 1 and all(getattr(h, "plan", None) is not None and h.plan.dtype == dtype for h in hops):
        # the whole iteration in one C call (rn_davidson): vectors, inner products and the
        # preconditioner stay on the device, the host sees two scalar blocks per iteration
        mask_u8 = mask.to(torch.uint8)
        e, c, nhop, _ = ops.davidson_plans([h.plan for h in hops], guesses[0], mask_u8, hdiag.to(torch.float64),
                                           inverse=inverse, max_cycle=100)
        hop.close()
        return e, _sign_fix(c).reshape(cshape), nhop
    e, c = davidson(aop, guesses, precond, max_cycle=100, nroots=nroots)
    hop.close()
    if nroots > 1:
        return e, [_sign_fix(ci).reshape(cshape) for ci in c], count[0]
    c = _sign_fix(c)
    return e, c.reshape(cshape), count[0]

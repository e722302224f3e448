#!/usr/bin/env python
"""Benchmark of the sweep hot path: DMRG/TDVP sweep sites per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload sbm_tdvp|holstein_dmrg|qc_dmrg|fmo_thermal] [--bond M] [--path 0|1]

The headline line is BASELINE.json configs[1]: spin-boson model, TDVP-PS time evolution, 20 phonon
modes (8 levels each), bond dimension M=256, complex128.  A "step" is one Mps.evolve call (TDVP: one
forward and one backward half sweep = 2 * nsite site updates) or one DMRG sweep (nsite - 1 two-site
updates).  BASELINE.json quotes the metric at M=256/512/1024, so at N=1 the same JSON line carries
`sub_results` for the other named configurations:
    holstein_dmrg  M=512   Holstein chain ground state, 20 molecules x 8 levels, one conserved exciton
    qc_dmrg        M=1024  ab initio Hamiltonian (synthetic integrals of example/h2o_qc.py's shape),
                           two conserved quantum numbers (N_alpha, N_beta), wide MPO bond
    fmo_thermal    M=512   7-site exciton model with long-range couplings, density operator (ancilla
                           index), one real-time TDVP-PS step
each with its own value / e2e / roofline / cpu_baseline / parity_check.

`parity_check`: the CPU oracle and the CUDA path visit the SAME strided sample of full-size site
updates from the same initial state (the other sites are passed over with a QR and the environment
update); gauge-invariant scalars of every sampled update (Davidson eigenvalue; <C|H_eff|C> after the
Krylov evolution; <C|H_eff|C> and |H_eff C| where only H_eff applications are sampled) must agree to
1e-10 (relative to max(1, |E|)).  The oracle's time on those site updates is the `cpu_baseline`.
--impl reference times the CPU oracle port of the reference's NumPy path alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "sweep_sites_per_sec"
UNIT = "sites/s"
PARITY_TOL = 1e-10

# name -> (default bond dimension, steps, warmup, CPU-sampled site updates) for the sub-results
SUB_WORKLOADS = {"holstein_dmrg": (512, 3, 3, 3), "qc_dmrg": (1024, 1, 2, 1), "fmo_thermal": (512, 1, 1, 1)}


# --------------------------------------------------------------------------------------------
# workload definition (shared by both arms)
# --------------------------------------------------------------------------------------------
def make_workload(name, bond, args, seed):
    from renormalizer_b200 import models
    rng = np.random.default_rng(seed)
    if name == "sbm_tdvp":
        nmodes, d = args.modes, args.levels
        omega, g = models.ohmic_modes(nmodes, alpha=0.05, omega_c=20.0)
        w = models.spin_boson_mpo(0.0, 1.0, omega, g, d)
        pdims = [2] + [d] * nmodes
        sites = models.random_mps_sites(pdims, bond, rng, dtype=np.complex128)
        n = len(sites)
        qn = [np.zeros((s.shape[0], 1), dtype=int) for s in sites] + [np.zeros((1, 1), dtype=int)]
        sq = [np.zeros((p, 1), dtype=int) for p in pdims]
        meta = dict(qn=qn, sigmaqn=sq, qntot=np.array([0]), qnidx=n - 1, to_right=False)
        return dict(key=name, kind="tdvp", mpo=w, sites=sites, meta=meta, nsite=n, sites_per_step=2 * n,
                    bond=bond, dt=args.dt, dtype="c128",
                    name=f"spin-boson TDVP-PS, {nmodes} modes x {d} levels, M={bond}, dt={args.dt}")
    if name == "holstein_dmrg":
        nmol, d = args.mols, args.levels
        w = models.holstein_mpo(nmol, d, e0=0.0, j=-0.1, omega=0.2, g=1.0)
        sq = models.holstein_sigmaqn(nmol, d)
        sites, qn = models.random_mps_qn(sq, [1], bond, rng)
        n = len(sites)
        meta = dict(qn=qn, sigmaqn=sq, qntot=np.array([1]), qnidx=n - 1, to_right=False)
        return dict(key=name, kind="dmrg", mpo=w, sites=sites, meta=meta, nsite=n, sites_per_step=n - 1,
                    bond=bond, dtype="f64", hop_only=False,
                    name=f"Holstein chain DMRG 2-site, {nmol} mols x {d} levels ({n} sites), M={bond}")
    if name == "qc_dmrg":
        ns = args.orbitals
        h1e, h2e = models.random_qc_integrals(ns, rng)
        w, _ = models.qc_mpo(h1e, h2e)
        sq = models.qc_sigmaqn(2 * ns)
        nel = [ns // 2, ns // 2]
        # Hartree-Fock determinant (lowest orbitals doubly occupied) + 1e-3 x a random state, as the
        # reference seeds its ab initio runs (mps/tests/test_gs.py:131-134)
        sites, qn = models.seeded_mps_qn(sq, nel, bond, rng, [1] * (2 * nel[0]) + [0] * (2 * ns - 2 * nel[0]))
        n = len(sites)
        meta = dict(qn=qn, sigmaqn=sq, qntot=np.array(nel), qnidx=n - 1, to_right=False)
        return dict(key=name, kind="dmrg", mpo=w, sites=sites, meta=meta, nsite=n, sites_per_step=n - 1,
                    bond=bond, dtype="f64", hop_only=True,
                    name=f"ab initio DMRG 2-site, synthetic integrals, {ns} spatial orbitals ({n} spin-orbital "
                         f"sites), N_alpha=N_beta={nel[0]}, MPO bond {max(t.shape[-1] for t in w)}, M={bond}")
    if name == "fmo_thermal":
        nmol, nmode, d = 7, args.fmo_modes, args.levels
        jm = rng.standard_normal((nmol, nmol)) * 0.3
        jm = 0.5 * (jm + jm.T)
        omegas = np.linspace(0.2, 1.5, nmode)
        w, _ = models.exciton_phonon_mpo(np.linspace(0.0, 1.0, nmol), jm, omegas, 0.7 / np.sqrt(1 + np.arange(nmode)), d)
        sq1 = models.exciton_phonon_sigmaqn(nmol, nmode, d)
        sites, qn, sq = models.random_mpdm_qn(sq1, [1], bond, rng)
        sites = [s.astype(np.complex128) for s in sites]
        n = len(sites)
        meta = dict(qn=qn, sigmaqn=sq, qntot=np.array([1]), qnidx=n - 1, to_right=False)
        return dict(key=name, kind="tdvp", mpo=w, sites=sites, meta=meta, nsite=n, sites_per_step=2 * n,
                    bond=bond, dt=args.dt, dtype="c128",
                    name=f"FMO-like 7-site exciton model, long-range J, {nmode} modes x {d} levels per site, "
                         f"density operator (ancilla index, {n} sites), TDVP-PS, M={bond}, dt={args.dt}")
    raise SystemExit(f"unknown workload {name}")


def sample_sites(work, k):
    """k site updates spread evenly over the first half sweep: the site indices, in sweep order
    (TDVP starts at the right end and moves left; the DMRG sweep starts at site 0 and moves right,
    its last two-site update is at nsite - 2)."""
    n = work["nsite"]
    order = list(range(n - 1, -1, -1)) if work["kind"] == "tdvp" else list(range(n - 1))
    k = max(1, min(k, len(order)))
    return [order[p] for p in sorted({int((i + 0.5) * len(order) / k) for i in range(k)})]


# --------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# CPU oracle arm: a strided sample of full-size site updates
# --------------------------------------------------------------------------------------------
def blas_threads():
    """Use every host core for the NumPy/BLAS arm (torch.distributed.run exports OMP_NUM_THREADS=1
    when it starts more than one rank, which silently serialised the CPU arm in round 1)."""
    want = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits, threadpool_info
        threadpool_limits(limits=want)
        got = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(got) if got else want
    except Exception:
        return int(os.environ.get("OMP_NUM_THREADS", want))


def cpu_sample(work, nsample, nhops_by_site=None, hops_timed=2):
    """Oracle port of the reference's NumPy path on `nsample` site updates spread evenly over the
    first half sweep from the workload's initial state.  Returns dict(sites, seconds per sampled
    site, gauge-invariant scalars per sampled site).  With work["hop_only"] only `hops_timed`
    H_eff applications, the SVD update and the environment update of a site are timed and the
    Davidson iteration is extrapolated with the CUDA path's application count at that site
    (the schedule is deterministic and identical; tests assert equal counts)."""
    from oracle import sweep as osw
    meta = work["meta"]
    om = osw.Mps(work["sites"], meta["qn"], meta["sigmaqn"], meta["qntot"], meta["qnidx"], meta["to_right"])
    if work["kind"] == "tdvp":
        return _oracle_tdvp_sample(osw, om, work, nsample)
    return _oracle_dmrg_sample(osw, om, work, nsample, nhops_by_site, hops_timed)


def _oracle_tdvp_sample(osw, mps_in, work, nsample):
    """First half sweep of oracle.sweep.evolve_tdvp_ps (same statements, see oracle/sweep.py:467),
    evolving only the sampled sites."""
    from oracle.contract import hop_apply
    from oracle.krylov import expm_krylov
    from oracle.svdqn import svd_qn
    mpo, dt = work["mpo"], work["dt"]
    mps = mps_in.to_complex()
    n = len(mps)
    order = list(mps.iter_idx_list(full=True))
    chosen = set(sample_sites(work, nsample))
    environ = osw.Environ(mps, mpo, "R" if mps.to_right else "L")
    times, scalars = {}, {}
    for imps in order:
        t0 = time.perf_counter()
        evolve_site = imps in chosen
        system = "L" if mps.to_right else "R"
        l_array, r_array = environ.read("L", imps - 1), environ.read("R", imps + 1)
        shape = list(mps.sites[imps].shape)
        w = mpo[imps]
        if evolve_site:
            mps_t, _ = expm_krylov(lambda y: hop_apply(l_array, r_array, [w], y.reshape(shape)).ravel(),
                                   -1j * dt / 2, mps.sites[imps].ravel())
            t_probe = time.perf_counter()
            hc = hop_apply(l_array, r_array, [w], mps_t.reshape(shape)).ravel()
            scalars[imps] = [float(np.vdot(mps_t, hc).real)]
            t0 += time.perf_counter() - t_probe                  # the probe is not part of the step
            mps_t = mps_t.reshape(shape)
        else:
            mps_t = mps.sites[imps]
        qnbigl, qnbigr, _ = mps.big_qn([imps])
        u, qnlset, v, qnrset = svd_qn(mps_t, qnbigl, qnbigr, mps.qntot, QR=True, system=system,
                                      full_matrices=False)
        vt = v.T
        if not mps.to_right and imps != 0:
            mps.sites[imps] = vt.reshape([-1] + shape[1:])
            mps.qn[imps] = np.array(qnrset)
            mps.qnidx = imps - 1
            r_array = environ.get_lr("R", imps, mps, mpo, "System")
            su = u.shape
            back = u
            if evolve_site:
                back, _ = expm_krylov(lambda y: hop_apply(l_array, r_array, [], y.reshape(su)).ravel(),
                                      1j * dt / 2, u.ravel())
            mps.sites[imps - 1] = np.tensordot(mps.sites[imps - 1], back.reshape(su), axes=(-1, 0))
        elif mps.to_right and imps != n - 1:
            mps.sites[imps] = u.reshape(shape[:-1] + [-1])
            mps.qn[imps + 1] = np.array(qnlset)
            mps.qnidx = imps + 1
            l_array = environ.get_lr("L", imps, mps, mpo, "System")
            sv = vt.shape
            back = vt
            if evolve_site:
                back, _ = expm_krylov(lambda y: hop_apply(l_array, r_array, [], y.reshape(sv)).ravel(),
                                      1j * dt / 2, vt.ravel())
            mps.sites[imps + 1] = np.tensordot(back.reshape(sv), mps.sites[imps + 1], axes=(1, 0))
        else:
            mps.sites[imps] = mps_t
        if evolve_site:
            times[imps] = time.perf_counter() - t0
    return dict(sites=sorted(chosen), times=times, scalars=scalars, extrapolated=False)


def _oracle_dmrg_sample(osw, mps, work, nsample, nhops_by_site, hops_timed):
    from oracle.contract import hop_apply
    mpo, bond = work["mpo"], work["bond"]
    mps.ensure_right_canonical()
    environ = osw.Environ(mps, mpo, "R")
    n = len(mps)
    order = [i for i in mps.iter_idx_list(full=True) if i != n - 1]
    chosen = set(sample_sites(work, nsample))
    times, scalars, stamp = {}, {}, [time.perf_counter()]
    if not work.get("hop_only"):
        def site_done(imps, e):
            now = time.perf_counter()
            if e is not None:
                times[imps] = now - stamp[0]
                scalars[imps] = [float(e)]
            stamp[0] = now
        osw.dmrg_single_sweep(mps, mpo, environ, "2site", bond, 0.0, None, site_filter=chosen, site_done=site_done)
        return dict(sites=sorted(chosen), times=times, scalars=scalars, extrapolated=False)
    # H_eff applications only (sites too large for a whole Davidson run within the CPU budget)
    for imps in order:
        t0 = time.perf_counter()
        cidx = [imps, imps + 1]
        lt = environ.get_lr("L", imps - 1, mps, mpo, "System")
        rt = environ.get_lr("R", imps + 2, mps, mpo, "Enviro")
        if imps not in chosen:
            mps.push_cano(imps)
            continue
        t_env = time.perf_counter() - t0
        cmo = [mpo[i] for i in cidx]
        c = np.tensordot(mps.sites[cidx[0]], mps.sites[cidx[1]], axes=1)
        t1 = time.perf_counter()
        for _ in range(hops_timed):
            hc = hop_apply(lt, rt, cmo, c)
        t_hop = (time.perf_counter() - t1) / hops_timed
        scalars[imps] = [float(np.vdot(c, hc).real), float(np.linalg.norm(hc))]
        qnbigl, qnbigr, _ = mps.big_qn(cidx)
        t2 = time.perf_counter()
        osw.update_mps(mps, c, cidx, qnbigl, qnbigr, bond, 0.0)
        t_upd = time.perf_counter() - t2
        nhop = (nhops_by_site or {}).get(imps, 10)
        times[imps] = t_env + nhop * t_hop + t_upd
    return dict(sites=sorted(chosen), times=times, scalars=scalars, extrapolated=True, hops_timed=hops_timed)


# --------------------------------------------------------------------------------------------
# GPU arm: sampled site updates captured on the CUDA path and replayed by the CPU oracle
# --------------------------------------------------------------------------------------------
def flop_quantile_sites(work, bond_dims, k):
    """k site updates at the (i + 0.5) / k quantiles of the H_eff FLOP distribution over a half sweep
    (2 m n k of the two GEMMs of one application): every sampled update stands for an equal share of
    the sweep's cost, so T_cpu(sweep) ~ T_gpu(sweep) * mean_i(t_cpu_i / t_gpu_i)."""
    mpo, n = work["mpo"], work["nsite"]
    span = 1 if work["kind"] == "tdvp" else 2
    order = list(range(n - 1, -1, -1)) if work["kind"] == "tdvp" else list(range(n - 1))
    cost = []
    for i in order:
        ml, mr = bond_dims[i], bond_dims[i + span]
        phys = 1
        for j in range(i, i + span):
            phys *= int(np.prod(work["sites"][j].shape[1:-1]))
        wl, wr = mpo[i].shape[0], mpo[i + span - 1].shape[3]
        cost.append(2.0 * ml * wl * ml * phys * mr + 2.0 * ml * phys * mr * wr * mr)
    cum = np.cumsum(cost) / np.sum(cost)
    k = max(1, min(k, len(order)))
    picks = sorted({int(np.searchsorted(cum, (i + 0.5) / k)) for i in range(k)})
    return [order[min(p, len(order) - 1)] for p in picks]


def gpu_capture(work, mpo, mps, chosen):
    """Run the sampled site updates on the CUDA path starting from `mps` (the other sites are passed
    over with a QR and the environment update).  For every sampled update: the inputs (environments,
    state) are downloaded for the oracle, the update is timed with a synchronise on both sides, and
    the gauge-invariant scalars of the parity check are recorded."""
    import torch
    from renormalizer_b200 import ops
    from renormalizer_b200.backend import asnumpy
    from renormalizer_b200.gs import single_sweep
    from renormalizer_b200.hop_expr import hop_expr_dtype
    from renormalizer_b200.lib import Environ
    chosen = set(chosen)
    caps = []
    if work["kind"] == "tdvp":
        # pass 1: parity scalars (Re <C|H_eff|C> after the forward evolution of every sampled site)
        scal = {}
        mps._evolve_tdvp_ps(mpo, work["dt"], site_filter=chosen,
                            site_probe=lambda i, e: scal.__setitem__(i, [e]), half_sweeps=1)

        def hook(stage, imps, info):
            torch.cuda.synchronize()
            if stage == "pre":
                m = info["mps"]
                nb = imps - 1 if not m.to_right else imps + 1
                cap = dict(imps=imps, lt=asnumpy(info["l_array"]), rt=asnumpy(info["r_array"]),
                           site=asnumpy(m[imps]), nb_idx=nb,
                           nb_site=asnumpy(m[nb]) if 0 <= nb < len(m) else None,
                           qn=[q.copy() for q in m.qn], qnidx=m.qnidx, to_right=m.to_right, scalars=scal[imps])
                caps.append(cap)
                torch.cuda.synchronize()
                cap["t0"] = time.perf_counter()
            else:
                caps[-1]["t_gpu"] = time.perf_counter() - caps[-1]["t0"]
        mps._evolve_tdvp_ps(mpo, work["dt"], site_filter=chosen, half_sweeps=1, site_hook=hook)
        return caps
    m = mps.copy()
    m.ensure_right_canonical()
    env = Environ(m, mpo, "R")

    def hook(stage, imps, info):
        torch.cuda.synchronize()
        if stage == "pre":
            st, cidx = info["mps"], info["cidx"]
            cap = dict(imps=imps, cidx=list(cidx), lt=asnumpy(info["ltensor"]), rt=asnumpy(info["rtensor"]),
                       sites={i: asnumpy(st[i]) for i in cidx}, qn=[q.copy() for q in st.qn], qnidx=st.qnidx,
                       to_right=st.to_right)
            if work.get("hop_only"):
                c = ops.tensordot1(st[cidx[0]], st[cidx[1]])
                hop = hop_expr_dtype(info["ltensor"], info["rtensor"], info["cmo"], list(c.shape), c.dtype)
                hc = hop(c)
                hop.close()
                cap["scalars"] = [float(torch.vdot(c.reshape(-1), hc.reshape(-1)).real),
                                  float(torch.linalg.vector_norm(hc))]
            caps.append(cap)
            torch.cuda.synchronize()
            cap["t0"] = time.perf_counter()
        else:
            cap = caps[-1]
            cap["t_gpu"] = time.perf_counter() - cap["t0"]
            cap["nhop"] = int(info["nhop"])
            if not work.get("hop_only"):
                cap["scalars"] = [float(info["e"])]
    single_sweep(m, mpo, env, None, 0.0, None, site_filter=chosen, site_hook=hook)
    return caps


def cpu_replay(work, cap):
    """One captured site update on the CPU oracle: (seconds, scalars, note).  The oracle statements are
    those of oracle/sweep.py (dmrg_single_sweep / evolve_tdvp_ps) for a single site."""
    from oracle import sweep as osw
    from oracle.contract import hop_apply, hop_diag, env_update
    from oracle.davidson import davidson
    from oracle.krylov import expm_krylov
    from oracle.svdqn import svd_qn, get_qn_mask
    meta, mpo, n = work["meta"], work["mpo"], work["nsite"]
    dummy = np.zeros((1, 1, 1))
    lt, rt = cap["lt"], cap["rt"]
    if work["kind"] == "tdvp":
        imps, dt = cap["imps"], work["dt"]
        sites = [dummy] * n
        sites[imps] = cap["site"]
        if cap["nb_site"] is not None:
            sites[cap["nb_idx"]] = cap["nb_site"]
        om = osw.Mps(sites, cap["qn"], meta["sigmaqn"], meta["qntot"], cap["qnidx"], cap["to_right"])
        shape = list(cap["site"].shape)
        w = mpo[imps]
        t0 = time.perf_counter()
        mps_t, _ = expm_krylov(lambda y: hop_apply(lt, rt, [w], y.reshape(shape)).ravel(), -1j * dt / 2,
                               cap["site"].astype(np.complex128).ravel())
        t_probe = time.perf_counter()
        hc = hop_apply(lt, rt, [w], mps_t.reshape(shape)).ravel()
        scalars = [float(np.vdot(mps_t, hc).real)]
        t0 += time.perf_counter() - t_probe                      # the probe is not part of the step
        mps_t = mps_t.reshape(shape)
        qnbigl, qnbigr, _ = om.big_qn([imps])
        system = "L" if om.to_right else "R"
        u, qnlset, v, qnrset = svd_qn(mps_t, qnbigl, qnbigr, om.qntot, QR=True, system=system, full_matrices=False)
        vt = v.T
        if not om.to_right and imps != 0:
            new = vt.reshape([-1] + shape[1:])
            r_new = env_update(rt, new, w, "R")
            su = u.shape
            back, _ = expm_krylov(lambda y: hop_apply(lt, r_new, [], y.reshape(su)).ravel(), 1j * dt / 2, u.ravel())
            np.tensordot(cap["nb_site"], back.reshape(su), axes=(-1, 0))
        elif om.to_right and imps != n - 1:
            new = u.reshape(shape[:-1] + [-1])
            l_new = env_update(lt, new, w, "L")
            sv = vt.shape
            back, _ = expm_krylov(lambda y: hop_apply(l_new, rt, [], y.reshape(sv)).ravel(), 1j * dt / 2, vt.ravel())
            np.tensordot(back.reshape(sv), cap["nb_site"], axes=(1, 0))
        return time.perf_counter() - t0, scalars, ""
    cidx = cap["cidx"]
    sites = [dummy] * n
    for i in cidx:
        sites[i] = cap["sites"][i]
    om = osw.Mps(sites, cap["qn"], meta["sigmaqn"], meta["qntot"], cap["qnidx"], cap["to_right"])
    cmo = [mpo[i] for i in cidx]
    qnbigl, qnbigr, qnmat = om.big_qn(cidx)
    mask = get_qn_mask(qnmat, om.qntot)
    guess = np.tensordot(sites[cidx[0]], sites[cidx[1]], axes=1)
    note = ""
    t0 = time.perf_counter()
    if not work.get("hop_only"):
        hdiag = hop_diag(lt, rt, cmo)[mask]
        e, c = davidson(lambda x: hop_apply(lt, rt, cmo, osw._scatter(x, mask))[mask], [guess[mask]],
                        lambda x, e, *a: x / (hdiag - e + 1e-4), max_cycle=100)
        scalars = [float(e)]
        cstruct = osw._scatter(osw._sign_fix(c), mask)
        t_solve = time.perf_counter() - t0
    else:
        # a whole Davidson run and the reference's full-matrices block SVD of a site this size take
        # minutes on the host: one H_eff application is timed and multiplied by the CUDA path's
        # application count at this site; the SVD update is left out of the oracle's time (which makes
        # the reported CPU figure an upper bound of the reference's speed)
        hc = hop_apply(lt, rt, cmo, guess)
        t_solve = (time.perf_counter() - t0) * cap["nhop"]
        scalars = [float(np.vdot(guess, hc).real), float(np.linalg.norm(hc))]
        note = (f"1 H_eff application timed and multiplied by the CUDA path's {cap['nhop']} applications at this "
                f"site, SVD update not timed")
        return t_solve, scalars, note
    t1 = time.perf_counter()
    osw.update_mps(om, cstruct, cidx, qnbigl, qnbigr, work["bond"], 0.0)
    return t_solve + time.perf_counter() - t1, scalars, note


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
class Gpu:
    """Process-wide CUDA / torch.distributed state of the GPU arm."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            import datetime
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank),
                                    timeout=datetime.timedelta(seconds=120))
        from renormalizer_b200 import _lib
        from renormalizer_b200.backend import backend
        _lib.get()
        backend.gemm_path = args.path
        self.lib = _lib

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, nsteps):
        torch = self.torch
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        start.record()
        for _ in range(nsteps):
            fn()
        end.record()
        self.barrier()
        ms = start.elapsed_time(end)
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms


def measure(gpu, args, name, bond, steps, warmup, nsample, do_cpu, do_parity):
    """One workload on the CUDA path: value, e2e, roofline, cpu_baseline, parity_check."""
    import gc
    torch = gpu.torch
    from renormalizer_b200.mpo import Mpo
    from renormalizer_b200.mps import Mps
    from renormalizer_b200.gs import single_sweep
    from renormalizer_b200.lib import Environ
    from renormalizer_b200.configs import CompressConfig, CompressCriteria, EvolveConfig, EvolveMethod

    # independent chains per rank (weak scaling: the path shards over independent sweep jobs)
    work = make_workload(name, bond, args, seed=1234 + gpu.rank)
    meta, mpo_host = work["meta"], work["mpo"]
    mpo = Mpo(mpo_host)

    def fresh_mps(sites, like=None):
        src = meta if like is None else dict(qn=[q.copy() for q in like.qn], sigmaqn=meta["sigmaqn"],
                                             qntot=meta["qntot"], qnidx=like.qnidx, to_right=like.to_right)
        m = Mps(sites, src["qn"], src["sigmaqn"], src["qntot"], src["qnidx"], src["to_right"])
        m.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps)       # the workload's integrator
        m.optimize_config.method = "2site"
        m.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=bond)
        return m

    state = {"mps": fresh_mps(work["sites"])}
    energies = []
    if work["kind"] == "dmrg":
        state["mps"].ensure_right_canonical()
        state["env"] = Environ(state["mps"], mpo, "R")

    def step_device():
        if work["kind"] == "tdvp":
            state["mps"] = state["mps"].evolve(mpo, work["dt"])
        else:
            micro, _, _ = single_sweep(state["mps"], mpo, state["env"], None, 0.0, None)
            energies.append(min(e for e, _ in micro))

    e_start = state["mps"].expectation(mpo) if work["kind"] == "tdvp" else None
    for _ in range(warmup):
        step_device()
    # keep the interpreter's cyclic garbage collector out of the timed region (a collection in the
    # middle of a half sweep shows up as a multi-millisecond host stall)
    gc.collect()
    gc.disable()
    sampler = ClockSampler(gpu.local_rank)
    if gpu.rank == 0:
        sampler.start()
    l0 = gpu.lib.LaunchCounter.total()
    if args.profiler_range:           # `ncu --profile-from-start off`: capture the timed steps only
        torch.cuda.cudart().cudaProfilerStart()
    # long steps (the sub-results): the per-launch GEMM events of the roofline are recorded during the
    # timed steps themselves (two event records per multi-millisecond launch) instead of in an extra step
    inline_roofline = gpu.rank == 0 and not args.no_roofline and name != "sbm_tdvp"
    if inline_roofline:
        gpu.lib.load().rn_profile_begin()
    ms = gpu.timed(step_device, steps)
    roof_raw = None
    if inline_roofline:
        import ctypes
        r_ms, r_fl, r_n = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_long(0)
        gpu.lib.load().rn_profile_end(ctypes.byref(r_ms), ctypes.byref(r_fl), ctypes.byref(r_n))
        roof_raw = (r_ms.value, r_fl.value, r_n.value)
    if args.profiler_range:
        torch.cuda.cudart().cudaProfilerStop()
    launches = gpu.lib.LaunchCounter.total() - l0
    clocks = sampler.stop() if gpu.rank == 0 else None
    gc.enable()
    value = work["sites_per_step"] * steps * gpu.world / (ms * 1e-3)

    # ---- what the timed steps computed: conservation laws of the evolution / variational descent
    drift = {}
    if work["kind"] == "tdvp":
        e_end = state["mps"].expectation(mpo)
        drift = {"norm_error": abs(state["mps"].mp_norm - 1.0),
                 "energy_drift_rel": abs(e_end - e_start) / max(1.0, abs(e_start)),
                 "steps_checked": warmup + steps}
    else:
        drift = {"sweep_energies": energies,
                 "energy_non_increasing": bool(all(b <= a + 1e-9 * max(1.0, abs(a)) for a, b in zip(energies, energies[1:])))}

    # ---- end to end: host (pinned) MPS + MPO in, new MPS out, every step --------------------
    e2e = None
    if not args.no_e2e:
        host_in = [torch.from_numpy(np.ascontiguousarray(s)).pin_memory() for s in state["mps"].to_numpy()]
        shapes_fixed = work["kind"] == "tdvp"
        host_out = [torch.empty_like(h).pin_memory() for h in host_in]
        h2d = sum(h.numel() * h.element_size() for h in host_in) + sum(w.nbytes for w in mpo_host)
        like = {"m": state["mps"]}

        def step_e2e():
            mpo_step = Mpo(mpo_host)                       # MPO site tensors uploaded again
            dev_sites = [h.to("cuda", non_blocking=True) for h in host_in]
            m = fresh_mps(dev_sites, like=like["m"])
            if work["kind"] == "tdvp":
                new = m.evolve(mpo_step, work["dt"])
            else:
                env = Environ(m, mpo_step, "R" if m.to_right else "L")
                single_sweep(m, mpo_step, env, None, 0.0, None)
                new = m
            if shapes_fixed:
                new.store_sites_to_host(host_out)
                torch.cuda.synchronize()
                host_in[:], host_out[:] = host_out[:], host_in[:]   # next step starts from the host result
            else:                                                  # bond dimensions may change in a DMRG sweep
                outs = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in new]
                for o, t in zip(outs, new):
                    o.copy_(t, non_blocking=True)
                torch.cuda.synchronize()
                host_in[:] = outs
            like["m"] = new
        for _ in range(warmup // 2):
            step_e2e()
        ms_e2e = gpu.timed(step_e2e, steps)
        d2h = sum(h.numel() * h.element_size() for h in host_in)
        e2e = {"value": work["sites_per_step"] * steps * gpu.world / (ms_e2e * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}

    out = {"workload": work["name"], "value": value, "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
           "warmup": warmup, "dtype": work["dtype"], "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    if gpu.rank != 0:
        return out, work
    # ---- roofline of the dominant kernel (GEMM of the H_eff chain), CUDA events per launch --
    out["roofline"] = None if args.no_roofline else gemm_roofline(args, step_device, work, roof_raw)
    # ---- CPU baseline + parity: sampled site updates of the current state, captured on the CUDA
    # path and replayed by the oracle (rank 0, N=1) ---------------------------------------------
    out["cpu_baseline"], out["parity_check"] = None, dict(drift)
    if gpu.world == 1 and (do_cpu or do_parity):
        cur = state["mps"]
        if work["kind"] == "tdvp" and (cur.to_right or cur.qnidx != work["nsite"] - 1):
            raise RuntimeError("TDVP state is expected to end a step at the right end, moving left")
        half = work["nsite"] if work["kind"] == "tdvp" else work["nsite"] - 1
        chosen = (sample_sites(work, nsample) if nsample >= half
                  else flop_quantile_sites(work, cur.bond_dims, nsample))
        caps = gpu_capture(work, mpo, cur, chosen)
        threads = blas_threads()
        t_cpu, ratios, notes, worst, count = [], [], set(), 0.0, 0
        for cap in caps:
            tc, scal, note = cpu_replay(work, cap)
            t_cpu.append(tc)
            ratios.append(tc / cap["t_gpu"])
            if note:
                notes.add(note)
            for x, y in zip(scal, cap["scalars"]):
                worst = max(worst, abs(x - y) / max(1.0, abs(x)))
                count += 1
        sites = [c["imps"] for c in caps]
        if len(caps) >= half:                      # every site update of a half sweep: a plain sum
            cpu_value = len(caps) / sum(t_cpu)
            est = "value = sampled updates / their summed oracle time"
        else:
            cpu_value = value / float(np.mean(ratios))
            est = ("sampled at equal shares of the sweep's H_eff FLOPs: value = CUDA value / mean over the "
                   "sampled updates of (oracle seconds / CUDA seconds of the same update, synchronised)")
        what = {"tdvp": "site evolution + QR + environment update + bond evolution",
                "dmrg": "eigensolver + SVD truncation update"}[work["kind"]]
        how = (f"{len(caps)} full-size site updates ({what}) of the state after the timed steps, sites {sites}, inputs "
               f"captured from the CUDA run, {sum(t_cpu):.1f} s of oracle time; {est}; NumPy/BLAS threads = {threads}")
        if notes:
            how += "; eigensolver: " + "; ".join(sorted(notes))
        out["cpu_baseline"] = {"value": cpu_value, "unit": UNIT, "cores": os.cpu_count(), "threads": threads,
                               "kind": "port", "sample": how}
        out["parity_check"].update({
            "what": ("Davidson eigenvalue of every sampled site update" if work["kind"] == "dmrg" and not work.get("hop_only")
                     else "<C|H_eff|C> and |H_eff C| at every sampled site" if work["kind"] == "dmrg"
                     else "<C|H_eff|C> after the Krylov evolution of every sampled site"),
            "sites": sites, "scalars_compared": count, "max_rel_err": worst, "tol": PARITY_TOL,
            "ok": bool(count > 0 and worst <= PARITY_TOL)})
    return out, work


def gemm_roofline(args, step_fn, work, raw=None):
    """Per-launch CUDA events around every GEMM of the H_eff / environment chain (instrumentation
    inside librn_b200.so) over one extra step -- or, when `raw` = (ms, flops, launches) is given,
    recorded during the timed steps -- and the achieved FLOP/s of that kernel."""
    import ctypes
    import torch
    from renormalizer_b200 import _lib
    lib = _lib.get()
    ms, flops, count = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_long(0)
    if raw is None:
        torch.cuda.synchronize()
        lib.rn_profile_begin()
        step_fn()
        torch.cuda.synchronize()
        lib.rn_profile_end(ctypes.byref(ms), ctypes.byref(flops), ctypes.byref(count))
    else:
        ms.value, flops.value, count.value = raw
    if count.value == 0 or ms.value <= 0:
        return None
    achieved = flops.value / (ms.value * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    if args.path == 0:
        peak, src = 40.0, "nominal B200 FP64 tensor (DMMA) peak; MEASURED_PEAKS.json has no FP64 figure"
    else:
        int8 = None
        try:
            int8 = json.load(open(os.path.join(ROOT, "profiles", "r02_int8_peak.json")))["int8_tops_sustained"]
        except (OSError, KeyError, ValueError):
            pass
        if int8:
            peak = int8 / 28
            src = ("FP64-equivalent, of measured: dense int8 tcgen05 rate measured on B200 (profiles/r02_int8_peak.json) "
                   "/ 28 digit products; algorithmic work = 2*m*n*k per launch")
        else:
            peak = peaks.get("bf16_tflops_sustained", 1400.0) * 2 / 28
            src = ("FP64-equivalent: 2 x bf16_tflops_sustained of MEASURED_PEAKS.json (int8 tensor rate) "
                   "/ 28 digit products; algorithmic work = 2*m*n*k per launch")
    traffic, traffic_src = None, None
    if args.path == 1:
        for fn in ("r02_gemm_traffic.json", "r01_gemm_traffic.json"):
            try:
                t = json.load(open(os.path.join(ROOT, "profiles", fn))).get(str(work["bond"]))
                if t:
                    traffic, traffic_src = t["dram_bytes_per_launch"], t["source"]
                    break
            except OSError:
                pass
    return {"bound": "tensor", "kernel": "gemm_tn_f64_kernel" if args.path == 0 else "ozaki_gemm_kernel",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src, "launches": int(count.value),
            "avg_launch_us": ms.value * 1e3 / count.value, "peak_source": src}


def strong_scaling(gpu, args, name, bond, steps, warmup):
    """ONE chain on all N GPUs: the same DMRG sweeps / TDVP steps run (a) on rank 0 alone and (b) with H_eff split
    over the bra-bond rows of L across the ranks and all-gathered over NVLink (parallel.ShardedHop);
    sites/s of both, max over ranks, and the sweep energies of both (they must agree)."""
    from renormalizer_b200 import parallel
    from renormalizer_b200.mpo import Mpo
    from renormalizer_b200.mps import Mps
    from renormalizer_b200.gs import single_sweep
    from renormalizer_b200.lib import Environ
    from renormalizer_b200.configs import CompressConfig, CompressCriteria, EvolveConfig, EvolveMethod
    work = make_workload(name, bond, args, seed=1234)            # the SAME chain on every rank
    # ... bit for bit: the MPO builder and the state generator go through LAPACK on the host, so rank 0's
    # arrays are broadcast (all ranks must take identical decisions, see parallel.ShardedHop)
    torch, dist = gpu.torch, gpu.dist
    for arrs in (work["mpo"], work["sites"]):
        shapes = torch.tensor([x for a in arrs for x in a.shape], dtype=torch.int64, device="cuda")
        ref_shapes = shapes.clone()
        dist.broadcast(ref_shapes, 0)
        if not torch.equal(shapes, ref_shapes):
            raise RuntimeError("workload shapes differ between ranks")
        for i, a in enumerate(arrs):
            t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
            dist.broadcast(torch.view_as_real(t) if t.is_complex() else t, 0)
            arrs[i] = t.cpu().numpy()
    meta = work["meta"]
    mpo = Mpo(work["mpo"])
    res = {}
    for mode in ("single", "sharded"):
        # every rank must take the same decisions (Davidson iteration counts decide how many all-gathers
        # are posted): the null-space completions of svd_qn draw from numpy's global generator
        np.random.seed(20261017)
        parallel.enable_sharded_heff(True if mode == "sharded" else None, min_work=args.shard_min_work)
        m = Mps(work["sites"], meta["qn"], meta["sigmaqn"], meta["qntot"], meta["qnidx"], meta["to_right"])
        m.optimize_config.method = "2site"
        m.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=bond)
        m.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps)
        st = {"m": m}
        if work["kind"] == "dmrg":
            m.ensure_right_canonical()
            env = Environ(m, mpo, "R")
        energies = []

        def step():
            # the single-GPU reference runs on rank 0 alone (the other ranks wait at the barrier), so
            # that it is not slowed by N processes sharing the host
            if mode == "sharded" or gpu.rank == 0:
                if work["kind"] == "dmrg":
                    micro, _, _ = single_sweep(m, mpo, env, None, 0.0, None)
                    energies.append(float(min(e for e, _ in micro)))
                else:
                    st["m"] = st["m"].evolve(mpo, work["dt"])
                    energies.append(float(np.real(st["m"].expectation(mpo))))
        for _ in range(warmup):
            step()
        ms = gpu.timed(step, steps)
        res[mode] = dict(value=work["sites_per_step"] * steps / (ms * 1e-3), ms_per_step=ms / steps, energies=energies)
        if mode == "sharded":
            res["stats"] = parallel.sharded_heff_stats()
    parallel.enable_sharded_heff(None)
    nsweeps = warmup + steps
    de = max([abs(a - b) for a, b in zip(res["single"]["energies"], res["sharded"]["energies"])] or [float("nan")])
    return {"workload": work["name"], "scaling": "strong", "n_gpus": gpu.world, "unit": UNIT,
            "value": res["sharded"]["value"], "value_1gpu_same_run": res["single"]["value"],
            "speedup_vs_1gpu": res["sharded"]["value"] / res["single"]["value"],
            "ms_per_step": res["sharded"]["ms_per_step"], "ms_per_step_1gpu": res["single"]["ms_per_step"],
            "sharding": "H_eff rows of L (bra bond) split over the ranks; Krylov/Davidson algebra, SVD and "
                        "environment update replicated",
            "collective": "one NCCL all-gather of the H_eff result per application",
            "heff_applications_sharded": res["stats"]["applications"],
            "gathered_bytes_per_step": res["stats"]["gathered_bytes"] // nsweeps,
            "energy_max_abs_diff_vs_1gpu": de, "energies": res["sharded"]["energies"]}


def run_ours(args):
    gpu = Gpu(args)
    main_nsample = args.cpu_sites if args.cpu_sites else (21 if args.workload == "sbm_tdvp" else SUB_WORKLOADS[args.workload][3])
    res, work = measure(gpu, args, args.workload, args.bond, args.steps, args.warmup, main_nsample,
                        do_cpu=not args.no_cpu_baseline, do_parity=not args.no_parity)
    subs = []
    if gpu.world == 1 and args.workload == "sbm_tdvp" and not args.no_sub:
        for name, (bond, steps, warmup, nsample) in SUB_WORKLOADS.items():
            if args.only_sub and name not in args.only_sub.split(","):
                continue
            gpu.torch.cuda.empty_cache()
            try:
                sub, _ = measure(gpu, args, name, bond, steps, warmup, nsample,
                                 do_cpu=not args.no_cpu_baseline, do_parity=not args.no_parity)
            except Exception as exc:                     # a failing sub-result must not take the headline line with it
                import traceback
                traceback.print_exc()
                sub = {"workload": name, "error": f"{type(exc).__name__}: {exc}"}
            subs.append(sub)
    strong = []
    if gpu.world > 1 and args.workload == "sbm_tdvp" and not args.no_strong:
        for name in (args.strong.split(",") if args.strong else []):
            bond, steps, warmup, _ = SUB_WORKLOADS[name]
            gpu.torch.cuda.empty_cache()
            strong.append(strong_scaling(gpu, args, name, bond, steps, warmup))
    if gpu.rank == 0:
        out = {
            "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": gpu.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": res["dtype"], "data": "synthetic",
            "config": {"workload": work["name"], "bond_dim": work["bond"], "nsite": work["nsite"],
                       "site_updates_per_step": work["sites_per_step"],
                       "gemm_path": "fp64-dmma" if args.path == 0 else "tcgen05-int8-split",
                       "parallelism": f"independent chains x{gpu.world}",
                       "l2_policy": "working set per step (environments + Krylov stacks) exceeds L2"},
            "clocks": res["clocks"], "e2e": res["e2e"], "gpu_launches": res["gpu_launches"],
            "roofline": res.get("roofline"), "cpu_baseline": res.get("cpu_baseline"),
            "parity_check": res.get("parity_check"), "sub_results": subs, "strong_scaling": strong,
        }
        print(json.dumps(out))
    if gpu.world > 1:
        gpu.dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = blas_threads()
    work = make_workload(args.workload, args.bond, args, seed=1234)
    nsample = args.cpu_sites if args.cpu_sites else (21 if args.workload == "sbm_tdvp" else SUB_WORKLOADS[args.workload][3])
    vals, sample = [], ""
    t_all0 = time.perf_counter()
    for i in range(args.warmup + args.steps):
        # the first warm-up pass is the full sample; later passes repeat it on fewer sites so that the
        # whole run stays within minutes (same estimator, noisier)
        c = cpu_sample(work, nsample if i == args.warmup else max(2, nsample // 4))
        tsum = sum(c["times"].values())
        if i >= args.warmup:
            vals.append(len(c["times"]) / tsum)
        if i == args.warmup:
            sample = (f"per step: site updates spread evenly over the first half sweep (first timed step: {len(c['times'])} "
                      f"updates, sites {c['sites']}, {tsum:.1f} s), the other sites passed over with a QR; "
                      f"NumPy/BLAS threads = {threads}")
            if c["extrapolated"]:
                sample += "; H_eff applications timed and multiplied by 10 per site (no Davidson run)"
    value = float(np.mean(vals))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": (time.perf_counter() - t_all0) * 1e3 / (args.warmup + args.steps),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": work["dtype"], "data": "synthetic",
           "config": {"workload": work["name"], "bond_dim": work["bond"], "nsite": work["nsite"],
                      "site_updates_per_step": work["sites_per_step"]},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "threads": threads,
                            "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sbm_tdvp", choices=["sbm_tdvp"] + list(SUB_WORKLOADS))
    ap.add_argument("--bond", type=int, default=None)
    ap.add_argument("--modes", type=int, default=20)
    ap.add_argument("--mols", type=int, default=20)
    ap.add_argument("--levels", type=int, default=8)
    ap.add_argument("--orbitals", type=int, default=12, help="spatial orbitals of the qc_dmrg workload")
    ap.add_argument("--fmo-modes", type=int, default=2, help="phonon modes per site of the fmo_thermal workload")
    ap.add_argument("--dt", type=float, default=0.05)
    ap.add_argument("--path", type=int, default=1)
    ap.add_argument("--cpu-sites", type=int, default=0, help="site updates in the CPU sample (0: workload default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the M=512 / M=1024 sub-results")
    ap.add_argument("--only-sub", default="", help="comma-separated subset of the sub-results")
    ap.add_argument("--strong", default="holstein_dmrg,fmo_thermal,qc_dmrg",
                    help="N > 1: workloads whose single chain is also run with H_eff split over all GPUs")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--shard-min-work", type=float, default=5.0e8,
                    help="H_eff applications below this many FLOPs stay on one GPU in the strong-scaling leg")
    ap.add_argument("--profiler-range", action="store_true",
                    help="bracket the timed steps with cudaProfilerStart/Stop (ncu launch lists)")
    args = ap.parse_args()
    if args.bond is None:
        args.bond = 256 if args.workload == "sbm_tdvp" else SUB_WORKLOADS[args.workload][0]
    if args.no_cpu_baseline:
        args.no_parity = True
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

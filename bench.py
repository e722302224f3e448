#!/usr/bin/env python
"""Benchmark of the sweep hot path: DMRG/TDVP sweep sites per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload sbm_tdvp|holstein_dmrg] [--bond M] [--path 0|1]

Default workload (BASELINE.json configs[1]): spin-boson model, TDVP-PS time evolution, 20 phonon
modes (8 levels each), bond dimension M=256, complex128.  A "step" is one Mps.evolve call = one
forward and one backward half sweep = 2 * nsite site updates (each: Krylov H_eff applications,
QR, environment update, backward bond evolution).  Prints ONE JSON line (see DESIGN.md for the
keys).  --impl reference times the CPU oracle port of the reference's NumPy path on the host
cores on a bounded sample of the same workload.
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


# --------------------------------------------------------------------------------------------
# workload definition (shared by both arms)
# --------------------------------------------------------------------------------------------
def make_workload(args, seed):
    from renormalizer_b200 import models
    rng = np.random.default_rng(seed)
    if args.workload == "sbm_tdvp":
        nmodes, d = args.modes, args.levels
        omega, g = models.ohmic_modes(nmodes, alpha=0.05, omega_c=20.0)
        w = models.spin_boson_mpo(0.0, 1.0, omega, g, d)
        pdims = [2] + [d] * nmodes
        sites = models.random_mps_sites(pdims, args.bond, rng, dtype=np.complex128)
        n = len(sites)
        qn = [np.zeros((s.shape[0], 1), dtype=int) for s in sites] + [np.zeros((1, 1), dtype=int)]
        sq = [np.zeros((p, 1), dtype=int) for p in pdims]
        meta = dict(qn=qn, sigmaqn=sq, qntot=np.array([0]), qnidx=n - 1, to_right=False)
        name = f"spin-boson TDVP-PS, {nmodes} modes x {d} levels, M={args.bond}, dt={args.dt}"
        return dict(kind="tdvp", mpo=w, sites=sites, meta=meta, name=name, nsite=n,
                    sites_per_step=2 * n)
    if args.workload == "holstein_dmrg":
        nmol, d = args.mols, args.levels
        w = models.holstein_mpo(nmol, d, e0=0.0, j=-0.1, omega=0.2, g=1.0)
        sq = models.holstein_sigmaqn(nmol, d)
        sites, qn = models.random_mps_qn(sq, [1], args.bond, rng)
        n = len(sites)
        meta = dict(qn=qn, sigmaqn=sq, qntot=np.array([1]), qnidx=n - 1, to_right=False)
        name = f"Holstein chain DMRG 2-site, {nmol} mols x {d} levels ({n} sites), M={args.bond}"
        return dict(kind="dmrg", mpo=w, sites=sites, meta=meta, name=name, nsite=n,
                    sites_per_step=n - 1)
    raise SystemExit(f"unknown workload {args.workload}")


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
# CPU oracle arm (bounded sample)
# --------------------------------------------------------------------------------------------
def cpu_sample(work, args, budget_s):
    """Time the oracle port of the reference's NumPy path on a bounded sample of the workload:
    whole site updates of the first half sweep, skipping the cheap boundary sites."""
    from oracle import sweep as osw
    meta = work["meta"]
    om = osw.Mps(work["sites"], meta["qn"], meta["sigmaqn"], meta["qntot"], meta["qnidx"],
                 meta["to_right"])
    mpo = work["mpo"]
    timer = _SiteTimer(budget_s, skip=args.cpu_skip_sites)
    try:
        if work["kind"] == "tdvp":
            _oracle_tdvp_sample(osw, om, mpo, args.dt, timer)
        else:
            _oracle_dmrg_sample(osw, om, mpo, args.bond, timer)
    except _Enough:
        pass
    return timer


class _Enough(Exception):
    pass


class _SiteTimer:
    def __init__(self, budget_s, skip):
        self.budget, self.skip = budget_s, skip
        self.seen = 0
        self.t0 = None
        self.timed_sites = 0
        self.elapsed = 0.0

    def site_done(self):
        self.seen += 1
        now = time.perf_counter()
        if self.seen == self.skip:
            self.t0 = now
            return
        if self.seen > self.skip:
            self.timed_sites = self.seen - self.skip
            self.elapsed = now - self.t0
            if self.elapsed > self.budget:
                raise _Enough


def _oracle_tdvp_sample(osw, mps_in, mpo, dt, timer):
    """oracle.sweep.evolve_tdvp_ps with a per-site hook (same code path, see oracle/sweep.py)."""
    from oracle.contract import hop_apply
    from oracle.krylov import expm_krylov
    from oracle.svdqn import svd_qn
    mps = mps_in.to_complex()
    n = len(mps)
    environ = osw.Environ(mps, mpo)
    if timer.skip == 0:
        timer.t0 = time.perf_counter()
    for _ in range(2):
        for imps in mps.iter_idx_list(full=True):
            system = "L" if mps.to_right else "R"
            l_array, r_array = environ.read("L", imps - 1), environ.read("R", imps + 1)
            shape = list(mps.sites[imps].shape)
            w = mpo[imps]
            mps_t, _ = expm_krylov(lambda y: hop_apply(l_array, r_array, [w], y.reshape(shape)).ravel(),
                                   -1j * dt / 2, mps.sites[imps].ravel())
            mps_t = mps_t.reshape(shape)
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
                back, _ = expm_krylov(lambda y: hop_apply(l_array, r_array, [], y.reshape(su)).ravel(),
                                      1j * dt / 2, u.ravel())
                mps.sites[imps - 1] = np.tensordot(mps.sites[imps - 1], back.reshape(su), axes=(-1, 0))
            elif mps.to_right and imps != n - 1:
                mps.sites[imps] = u.reshape(shape[:-1] + [-1])
                mps.qn[imps + 1] = np.array(qnlset)
                mps.qnidx = imps + 1
                l_array = environ.get_lr("L", imps, mps, mpo, "System")
                sv = vt.shape
                back, _ = expm_krylov(lambda y: hop_apply(l_array, r_array, [], y.reshape(sv)).ravel(),
                                      1j * dt / 2, vt.ravel())
                mps.sites[imps + 1] = np.tensordot(back.reshape(sv), mps.sites[imps + 1], axes=(1, 0))
            else:
                mps.sites[imps] = mps_t
            timer.site_done()
        mps.switch_direction()


def _oracle_dmrg_sample(osw, mps, mpo, bond, timer):
    if mps.qnidx == len(mps) - 1:
        mps.ensure_right_canonical()
        env = "R"
    else:
        mps.ensure_left_canonical()
        env = "L"
    environ = osw.Environ(mps, mpo, env)
    if timer.skip == 0:
        timer.t0 = time.perf_counter()

    class Hook(list):
        def append(self, x):
            timer.site_done()
    while True:
        osw.dmrg_single_sweep(mps, mpo, environ, "2site", bond, 0.0, None, stats=Hook())


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from renormalizer_b200 import _lib
    from renormalizer_b200.backend import backend
    from renormalizer_b200.mpo import Mpo
    from renormalizer_b200.mps import Mps
    from renormalizer_b200.gs import single_sweep
    from renormalizer_b200.lib import Environ
    from renormalizer_b200.configs import CompressConfig, CompressCriteria, EvolveConfig, EvolveMethod
    _lib.get()
    backend.gemm_path = args.path

    # independent chains per rank (weak scaling: the path shards over independent sweep jobs)
    work = make_workload(args, seed=1234 + rank)
    meta = work["meta"]
    mpo_host = work["mpo"]
    mpo = Mpo(mpo_host)

    def fresh_mps(sites):
        m = Mps(sites, meta["qn"], meta["sigmaqn"], meta["qntot"], meta["qnidx"], meta["to_right"])
        m.evolve_config = EvolveConfig(EvolveMethod.tdvp_ps)       # the workload's integrator
        return m

    state = {"mps": fresh_mps(work["sites"])}
    if work["kind"] == "dmrg":
        m = state["mps"]
        m.optimize_config.method = "2site"
        m.compress_config = CompressConfig(CompressCriteria.fixed, max_bonddim=args.bond)
        m.ensure_right_canonical()
        state["env"] = Environ(m, mpo, "R")

    def step_device():
        if work["kind"] == "tdvp":
            state["mps"] = state["mps"].evolve(mpo, args.dt)
        else:
            single_sweep(state["mps"], mpo, state["env"], None, 0.0, None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, nsteps):
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record()
        for _ in range(nsteps):
            fn()
        end.record()
        barrier()
        ms = start.elapsed_time(end)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(args.warmup):
        step_device()
    # keep the interpreter's cyclic garbage collector out of the timed region (a collection in the
    # middle of a half sweep shows up as a multi-millisecond host stall)
    import gc
    gc.collect()
    gc.disable()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _lib.LaunchCounter.total()
    if args.profiler_range:           # `ncu --profile-from-start off`: capture the timed steps only
        torch.cuda.cudart().cudaProfilerStart()
    ms = timed(step_device, args.steps)
    if args.profiler_range:
        torch.cuda.cudart().cudaProfilerStop()
    launches = _lib.LaunchCounter.total() - l0
    clocks = sampler.stop() if rank == 0 else None
    sites_total = work["sites_per_step"] * args.steps * world
    value = sites_total / (ms * 1e-3)

    # ---- end to end: host (pinned) MPS + MPO in, new MPS out, every step --------------------
    e2e = None
    if work["kind"] == "tdvp" and not args.no_e2e:
        host_in = [torch.from_numpy(np.ascontiguousarray(s)).pin_memory() for s in state["mps"].to_numpy()]
        host_out = [torch.empty_like(h).pin_memory() for h in host_in]
        h2d = sum(h.numel() * h.element_size() for h in host_in) + sum(w.nbytes for w in mpo_host)
        d2h = sum(h.numel() * h.element_size() for h in host_out)

        def step_e2e():
            mpo_step = Mpo(mpo_host)                       # MPO site tensors uploaded again
            dev_sites = [h.to("cuda", non_blocking=True) for h in host_in]
            m = fresh_mps(dev_sites)
            m.qn = [q.copy() for q in state["mps"].qn]
            m.qnidx, m.to_right = state["mps"].qnidx, state["mps"].to_right
            new = m.evolve(mpo_step, args.dt)
            new.store_sites_to_host(host_out)
            torch.cuda.synchronize()
            host_in[:], host_out[:] = host_out[:], host_in[:]   # next step starts from the host result
        for _ in range(max(1, args.warmup // 2)):
            step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
        e2e = {"value": work["sites_per_step"] * args.steps * world / (ms_e2e * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}

    # ---- roofline of the dominant kernel (GEMM of the H_eff chain), CUDA events per launch --
    roofline = None
    if rank == 0 and not args.no_roofline:
        roofline = gemm_roofline(args, step_device)

    # ---- CPU baseline (oracle port) on a bounded sample, rank 0 at N=1 only ------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        t = cpu_sample(work, args, args.cpu_budget)
        if t.timed_sites:
            cpu = {"value": t.timed_sites / t.elapsed, "unit": UNIT, "cores": os.cpu_count(),
                   "kind": "port",
                   "sample": f"{t.timed_sites} full-size site updates of the first half sweep "
                             f"(after skipping {args.cpu_skip_sites} boundary sites), {t.elapsed:.1f} s, "
                             f"NumPy/BLAS threads = all host cores"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "c128" if work["kind"] == "tdvp" else "f64",
            "data": "synthetic",
            "config": {"workload": work["name"], "bond_dim": args.bond, "nsite": work["nsite"],
                       "site_updates_per_step": work["sites_per_step"],
                       "gemm_path": "fp64-dmma" if args.path == 0 else "tcgen05-int8-split",
                       "parallelism": f"independent chains x{world}",
                       "l2_policy": "working set per step (environments + Krylov stacks) exceeds L2"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def gemm_roofline(args, step_fn):
    """Re-run one step with per-launch CUDA events around every GEMM of the H_eff / environment
    chain (instrumentation inside librn_b200.so) and report achieved FLOP/s of that kernel."""
    import ctypes
    import torch
    from renormalizer_b200 import _lib
    lib = _lib.get()
    if not hasattr(lib, "rn_profile_begin"):
        return None
    lib.rn_profile_begin.restype = ctypes.c_int
    lib.rn_profile_end.restype = ctypes.c_int
    lib.rn_profile_end.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                   ctypes.POINTER(ctypes.c_long)]
    torch.cuda.synchronize()
    lib.rn_profile_begin()
    step_fn()
    torch.cuda.synchronize()
    ms, flops, count = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_long(0)
    lib.rn_profile_end(ctypes.byref(ms), ctypes.byref(flops), ctypes.byref(count))
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
        peak = peaks.get("bf16_tflops_sustained", 1400.0) * 2 / 28
        src = ("FP64-equivalent, of measured: 2 x bf16_tflops_sustained of MEASURED_PEAKS.json (int8 tensor rate) "
               "/ 28 digit products; algorithmic work = 2*m*n*k per launch")
    traffic, traffic_src = None, None
    if args.path == 1:
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", "r01_gemm_traffic.json"))).get(str(args.bond))
            if t:
                traffic, traffic_src = t["dram_bytes_per_launch"], t["source"]
        except OSError:
            pass
    return {"bound": "tensor", "kernel": "gemm_tn_f64_kernel" if args.path == 0 else "ozaki_gemm_kernel",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src, "launches": int(count.value),
            "avg_launch_us": ms.value * 1e3 / count.value, "peak_source": src}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    work = make_workload(args, seed=1234)
    vals = []
    t_all0 = time.perf_counter()
    sample = ""
    for i in range(args.warmup + args.steps):
        t = cpu_sample(work, args, args.cpu_budget)
        v = t.timed_sites / t.elapsed if t.timed_sites else 0.0
        if i >= args.warmup:
            vals.append(v)
        sample = (f"{t.timed_sites} full-size site updates of the first half sweep per step "
                  f"(after skipping {args.cpu_skip_sites} boundary sites), ~{t.elapsed:.1f} s each")
    value = float(np.mean(vals))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": (time.perf_counter() - t_all0) * 1e3 / (args.warmup + args.steps),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "c128" if work["kind"] == "tdvp" else "f64", "data": "synthetic",
           "config": {"workload": work["name"], "bond_dim": args.bond, "nsite": work["nsite"],
                      "site_updates_per_step": work["sites_per_step"]},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                            "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sbm_tdvp", choices=["sbm_tdvp", "holstein_dmrg"])
    ap.add_argument("--bond", type=int, default=256)
    ap.add_argument("--modes", type=int, default=20)
    ap.add_argument("--mols", type=int, default=20)
    ap.add_argument("--levels", type=int, default=8)
    ap.add_argument("--dt", type=float, default=0.05)
    ap.add_argument("--path", type=int, default=1)
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--cpu-skip-sites", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--profiler-range", action="store_true",
                    help="bracket the timed steps with cudaProfilerStart/Stop (ncu launch lists)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

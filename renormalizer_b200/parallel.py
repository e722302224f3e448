"""Multi-GPU execution of the sweep path: one process per GPU, torch.distributed for the plumbing
(NCCL over NVLink on GPUs, gloo in the CPU tests).

Two modes:

* independent sweep jobs (`run_sharded`): every rank owns a contiguous slice of a job list
  (trajectories of an ensemble, parameter scans, independent initial states), runs its sweeps on
  its own GPU, and only the per-job results (a few scalars each) are gathered at the end -- weak
  scaling, no data-path collective;
* ONE sweep on several GPUs (`enable_sharded_heff` / `ShardedHop`): the effective-Hamiltonian
  application -- the O(M^3) part of every site update -- is split over the bra-bond rows `a` of
  "abc, bdef, lfk, cek -> adl" (renormalizer/mps/hop_expr.py:74-78): rank r holds L[a_r, :, :],
  computes out[a_r, ...] with its own H_eff plan (G1 -> MPO application -> G3 on 1/N of the rows)
  and the slices are all-gathered (contiguous in memory: `a` is the slowest index) into the full
  vector on every rank, one NCCL all-gather per application.  Everything else of the site update
  (Krylov / Davidson vector algebra, SVD truncation, environment update) runs replicated and
  deterministically on every rank, so all ranks take identical decisions and the results are those of
  the single-GPU sweep.  Strong scaling is bounded by the H_eff share of the sweep (large at
  M = 1024 with a wide MPO bond, small at M = 512 with the 5-state Holstein MPO).
"""
import os

import numpy as np


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process -> (0, 1, 0))."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(n_items, rank, world_size):
    """Contiguous, balanced slice [lo, hi) of n_items for `rank` (first n % world ranks get one more)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("invalid rank / world size")
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def init_process_group(backend=None):
    """Initialise torch.distributed from the environment; returns (rank, world_size)."""
    import torch
    import torch.distributed as dist
    rank, world_size, local_rank = world()
    if world_size == 1:
        return rank, world_size
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group(backend)
    return rank, world_size


def run_sharded(jobs, step_fn, result_width):
    """Run step_fn(job) -> sequence of `result_width` floats for this rank's slice of `jobs` and
    return the (len(jobs), result_width) array of all results on every rank (one all-gather of
    the small result table; the sweeps themselves never communicate)."""
    import torch
    import torch.distributed as dist
    rank, world_size, _ = world()
    lo, hi = shard_range(len(jobs), rank, world_size)
    local = np.zeros((hi - lo, result_width), dtype=np.float64)
    for i, job in enumerate(jobs[lo:hi]):
        local[i] = np.asarray(step_fn(job), dtype=np.float64)
    if world_size == 1 or not dist.is_initialized():
        return local
    use_cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    # pad to the largest shard so all_gather sees equal shapes
    width = -(-len(jobs) // world_size)
    buf = torch.zeros((width, result_width), dtype=torch.float64, device=dev)
    buf[:hi - lo] = torch.from_numpy(local).to(dev)
    out = [torch.zeros_like(buf) for _ in range(world_size)]
    dist.all_gather(out, buf)
    rows = []
    for r, t in enumerate(out):
        rlo, rhi = shard_range(len(jobs), r, world_size)
        rows.append(t[:rhi - rlo].cpu().numpy())
    return np.concatenate(rows, axis=0)


# --------------------------------------------------------------------------------------------
# one sweep on several GPUs: H_eff split over the bra-bond rows of L
# --------------------------------------------------------------------------------------------
_heff = {"group": None, "min_work": 5.0e8, "applications": 0, "gathered_bytes": 0}


def enable_sharded_heff(group=True, min_work=5.0e8):
    """Split every H_eff application whose two GEMMs exceed `min_work` FLOPs over the ranks of `group`
    (True: the default process group; None / False: off)."""
    _heff["group"] = None if group in (None, False) else group
    _heff["min_work"] = float(min_work)
    _heff["applications"] = 0
    _heff["gathered_bytes"] = 0


def sharded_heff_stats():
    return dict(applications=_heff["applications"], gathered_bytes=_heff["gathered_bytes"])


def heff_group():
    """The process group H_eff is split over (True: the default group), or None when sharding is off."""
    import torch.distributed as dist
    g = _heff["group"]
    if g is None or not dist.is_available() or not dist.is_initialized():
        return None
    if dist.get_world_size(None if g is True else g) < 2:
        return None
    return g


def heff_flops(lshape, rshape, cshape):
    """2 m n k of the two GEMMs of one application (G1: L.C, G3: T.R)."""
    la, lb, lc = lshape
    rl, rf, rk = rshape
    inner = 1
    for d in cshape[1:-1]:
        inner *= int(d)
    return 2.0 * la * lb * lc * inner * rk + 2.0 * la * inner * rf * rk * rl


class ShardedHop:
    """expr(cstruct) = H_eff . cstruct with the rows `a` of L (and of the result) split over the ranks
    of a process group.  `make_local(l_slice)` returns the rank-local callable (cstruct -> out slice,
    shape (rows, ..., Rl)); on the GPU that is an H_eff plan built on L[lo:hi] (hop_expr.py)."""

    plan = None          # no single-GPU plan: callers fall back to their generic (callable) drivers

    def __init__(self, ltensor, make_local, out_inner_shape, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = None if group is True else group
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.la = int(ltensor.shape[0])
        self.width = -(-self.la // self.world)                 # rows per rank, last ranks may be short / empty
        self.lo = min(self.rank * self.width, self.la)
        self.hi = min(self.lo + self.width, self.la)
        self.inner = tuple(int(x) for x in out_inner_shape)    # (d..., Rl)
        self.local = make_local(ltensor[self.lo:self.hi].contiguous()) if self.hi > self.lo else None
        self.even = self.la == self.width * self.world

    def __call__(self, cstruct):
        import torch
        part = self.local(cstruct) if self.local is not None else None
        ref = part if part is not None else cstruct
        full = torch.empty((self.world * self.width,) + self.inner, dtype=ref.dtype, device=ref.device)
        if self.even:
            send = part.contiguous()
        else:
            send = torch.zeros((self.width,) + self.inner, dtype=ref.dtype, device=ref.device)
            if part is not None:
                send[:self.hi - self.lo] = part
        if full.is_complex():                   # collectives see the interleaved real view
            self.dist.all_gather_into_tensor(torch.view_as_real(full), torch.view_as_real(send), group=self.group)
        else:
            self.dist.all_gather_into_tensor(full, send, group=self.group)
        _heff["applications"] += 1
        _heff["gathered_bytes"] += full.numel() * full.element_size()
        return full if self.even else full[:self.la].contiguous()

    def close(self):
        if self.local is not None and hasattr(self.local, "close"):
            self.local.close()

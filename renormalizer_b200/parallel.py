"""Multi-GPU sharding of independent sweep jobs: one process per GPU, no data-path collective.

The sweep path partitions over independent chains (trajectories of an ensemble, parameter scans,
independent initial states): every rank owns a contiguous slice of the job list, runs its sweeps
on its own GPU, and only the per-job results (a few scalars each) are gathered at the end.
`torch.distributed` provides the plumbing (NCCL on GPUs, gloo in the CPU tests).

Sharding ONE chain into contiguous site segments with a single all-gather of boundary
environments per sweep (real-space parallel DMRG/TDVP) is the next step of DESIGN.md row (e); it
changes the algorithm (inverse-gauge matrices on the segment bonds) and is not done here.
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

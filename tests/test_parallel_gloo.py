"""World-size-2 gloo test of the multi-process sharding logic (CPU, no GPU needed)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from renormalizer_b200 import parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 8, 21, 100):
        for w in (1, 2, 3, 8):
            cover = []
            for r in range(w):
                lo, hi = parallel.shard_range(n, r, w)
                assert 0 <= lo <= hi <= n
                cover += list(range(lo, hi))
            assert cover == list(range(n))
            sizes = [parallel.shard_range(n, r, w)[1] - parallel.shard_range(n, r, w)[0] for r in range(w)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_run_sharded_two_ranks_gloo(tmp_path):
    """Each rank 'sweeps' its own jobs with the CPU oracle's TDVP step; the gathered table must
    equal a single-process run."""
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np
        from renormalizer_b200 import parallel, models
        from oracle import sweep as osw

        def step(seed):
            rng = np.random.default_rng(seed)
            omega, g = models.ohmic_modes(3, alpha=0.3, omega_c=5.0)
            w = models.spin_boson_mpo(0.1, 1.0, omega, g, 3)
            sites = models.random_mps_sites([2, 3, 3, 3], 6, rng, dtype=np.complex128)
            qn = [np.zeros((s.shape[0], 1), dtype=int) for s in sites] + [np.zeros((1, 1), dtype=int)]
            sq = [np.zeros((s.shape[1], 1), dtype=int) for s in sites]
            m = osw.Mps(sites, qn, sq, [0], 3, False)
            m1 = osw.evolve_tdvp_ps(m, w, 0.05)
            return [m1.expectation(w), m1.mp_norm]

        rank, world = parallel.init_process_group("gloo")
        table = parallel.run_sharded([11, 12, 13, 14, 15], step, 2)
        if rank == 0:
            np.save(sys.argv[1], table)
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    """))
    port = _free_port()
    out2 = tmp_path / "two.npy"
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script), str(out2)], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    out1 = tmp_path / "one.npy"
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", OMP_NUM_THREADS="1")
    assert subprocess.run([sys.executable, str(script), str(out1)], env=env, timeout=300).returncode == 0
    a, b = np.load(out1), np.load(out2)
    assert a.shape == (5, 2) and b.shape == (5, 2)
    assert np.abs(a - b).max() < 1e-12


@pytest.mark.parametrize("la", [8, 7])
def test_sharded_heff_two_ranks_gloo(tmp_path, la):
    """ShardedHop on two gloo ranks: each rank applies H_eff with its own rows of L (the oracle's
    contraction stands in for the rank-local CUDA plan), the slices are all-gathered, and every rank
    holds the single-rank result -- for an even split and for a ragged one (7 rows on 2 ranks)."""
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch
        import torch.distributed as dist
        from renormalizer_b200 import parallel
        from oracle.contract import hop_apply
        rank, world = parallel.init_process_group("gloo")
        rng = np.random.default_rng(5)                       # same operands on every rank
        la, w, m, d = {la}, 3, 6, 4
        L = rng.standard_normal((la, w, m)) + 1j * rng.standard_normal((la, w, m))
        R = rng.standard_normal((m, w, m)) + 1j * rng.standard_normal((m, w, m))
        W = rng.standard_normal((w, d, d, w))
        C = rng.standard_normal((m, d, m)) + 1j * rng.standard_normal((m, d, m))
        parallel.enable_sharded_heff(True, min_work=0.0)
        assert parallel.heff_group() is not None
        make_local = lambda l_slice: (lambda c: torch.from_numpy(hop_apply(l_slice.numpy(), R, [W], c.numpy())))
        hop = parallel.ShardedHop(torch.from_numpy(L), make_local, (d, m), group=parallel.heff_group())
        lo, hi = hop.lo, hop.hi
        assert (hi - lo) in (la // 2, la - la // 2, -(-la // 2))
        out = hop(torch.from_numpy(C)).numpy()
        ref = hop_apply(L, R, [W], C)
        assert out.shape == ref.shape
        assert np.abs(out - ref).max() < 1e-13, np.abs(out - ref).max()
        st = parallel.sharded_heff_stats()
        assert st["applications"] == 1 and st["gathered_bytes"] >= ref.nbytes
        # every rank holds the same bits
        t = torch.from_numpy(np.ascontiguousarray(out)).view(torch.float64).clone()
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        mn = t.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        assert torch.equal(mx, mn)
        if rank == 0:
            np.save(sys.argv[1], out)
        dist.destroy_process_group()
    """))
    port = _free_port()
    out = tmp_path / "out.npy"
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script), str(out)], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    assert out.exists()


def test_distributed_block_svds_two_ranks_gloo(tmp_path):
    """One sweep on several GPUs: the quantum-number blocks of a bond are dealt to the ranks, each rank
    factorises its own and the factors are broadcast (svd_qn._economic_svds_distributed).  Host logic
    only: the kernels are replaced by the CPU test double of tests/_host_logic_stub.py."""
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
        import numpy as np, torch
        import torch.distributed as dist
        import _host_logic_stub                                  # ops.svd -> torch.linalg.svd on the CPU
        from renormalizer_b200 import parallel, svd_qn
        rank, world = parallel.init_process_group("gloo")
        parallel.enable_sharded_heff(True, min_work=0.0)
        rng = np.random.default_rng(3)
        blocks = [torch.from_numpy(rng.standard_normal(s)) for s in [(90, 70), (8, 5), (64, 130), (100, 100)]]
        blocks.append(torch.from_numpy(rng.standard_normal((70, 66)) + 1j * rng.standard_normal((70, 66))))
        out = svd_qn._economic_svds(blocks)
        assert len(out) == len(blocks)
        for b, (u, s, vh) in zip(blocks, out):
            assert torch.allclose((u * s.to(u.dtype)) @ vh, b, atol=1e-12)
            for t in (u, s, vh):                                 # the same bits on every rank
                r = torch.view_as_real(t) if t.is_complex() else t
                mx = r.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                mn = r.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
                assert torch.equal(mx, mn)
        if rank == 0:
            open(sys.argv[1], "w").write("ok")
        dist.destroy_process_group()
    """))
    port = _free_port()
    out = tmp_path / "out.txt"
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script), str(out)], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    assert out.read_text() == "ok"


@pytest.mark.parametrize("domain", ["L", "R"])
def test_sharded_environment_update_two_ranks_gloo(tmp_path, domain):
    """contract_one_site split over the ket's new bond across two gloo ranks (host logic with the CPU
    test double for the kernel) equals the single-rank environment update."""
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
        import numpy as np, torch
        import torch.distributed as dist
        import _host_logic_stub
        from renormalizer_b200 import parallel
        from renormalizer_b200.lib import contract_one_site
        from oracle.contract import env_update
        rank, world = parallel.init_process_group("gloo")
        rng = np.random.default_rng(9)
        ea, w, ec, d, mf, mh = 5, 3, 6, 4, 7, 8
        env = rng.standard_normal((ea, w, ec)) + 1j * rng.standard_normal((ea, w, ec))
        mo = rng.standard_normal((w, d, d, w))
        if {domain!r} == "L":
            ket = rng.standard_normal((ec, d, mh)) + 1j * rng.standard_normal((ec, d, mh))
            bra = rng.standard_normal((ea, d, mf)) + 1j * rng.standard_normal((ea, d, mf))
        else:
            ket = rng.standard_normal((mh, d, ec)) + 1j * rng.standard_normal((mh, d, ec))
            bra = rng.standard_normal((mf, d, ea)) + 1j * rng.standard_normal((mf, d, ea))
        ref = env_update(env, ket, mo, {domain!r}, ms_conj=bra.conj())
        parallel.enable_sharded_heff(True, min_work=0.0)
        got = contract_one_site(torch.from_numpy(env), torch.from_numpy(ket), mo, {domain!r},
                                ms_conj=torch.from_numpy(bra).conj()).numpy()
        assert got.shape == ref.shape and np.abs(got - ref).max() < 1e-12
        assert parallel.sharded_heff_stats()["gathered_bytes"] >= ref.nbytes
        if rank == 0:
            open(sys.argv[1], "w").write("ok")
        dist.destroy_process_group()
    """))
    port = _free_port()
    out = tmp_path / "out.txt"
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script), str(out)], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    assert out.read_text() == "ok"

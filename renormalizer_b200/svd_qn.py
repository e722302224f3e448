"""Quantum-number blocked SVD / QR of a centre tensor on the device.

Mirror of renormalizer/mps/svd_qn.py:97-314: same arguments and return tuples as the reference's
svd_qn; the per-block scipy.linalg.svd / qr / rq calls become rn_svd_jacobi / rn_qr / rn_lq.
Quantum numbers stay on the host (tiny integer arrays); matrices never leave the device --
only the singular values are copied back, because the caller's truncation logic needs them.
"""
import numpy as np
import torch

from . import ops
from .backend import asxp


def add_outer(a: np.ndarray, b: np.ndarray):
    """np.add.outer keeping the last (quantum number component) axis; svd_qn.py:302-310."""
    assert a.shape[-1] == b.shape[-1]
    sa, sb = a.shape[:-1], b.shape[:-1]
    return a.reshape(sa + (1,) * len(sb) + (-1,)) + b.reshape((1,) * len(sa) + sb + (-1,))


def get_qn_mask(qnmat: np.ndarray, qntot):
    """svd_qn.py:313-314."""
    return np.all(qnmat == np.array(qntot), axis=-1)


def qn_mask_outer(qnbigl: np.ndarray, qnbigr: np.ndarray, qntot):
    """get_qn_mask(add_outer(qnbigl, qnbigr), qntot) without materialising the int64 outer sum:
    one narrow-integer comparison per quantum-number component."""
    sl, sr = qnbigl.shape[:-1], qnbigr.shape[:-1]
    l = qnbigl.reshape(-1, qnbigl.shape[-1])
    r = np.asarray(qntot).reshape(1, -1) - qnbigr.reshape(-1, qnbigr.shape[-1])
    lo, hi = min(l.min(initial=0), r.min(initial=0)), max(l.max(initial=0), r.max(initial=0))
    dt = np.int8 if -128 <= lo and hi < 128 else np.int32
    l, r = l.astype(dt), r.astype(dt)
    mask = l[:, None, 0] == r[None, :, 0]
    for k in range(1, l.shape[1]):
        mask &= l[:, None, k] == r[None, :, k]
    return mask.reshape(sl + sr)


_rank_cache = {}


def economic_rank(qnbigl: np.ndarray, qnbigr: np.ndarray, qntot):
    """Number of singular vectors svd_qn returns with full_matrices=False: the sum over the
    quantum-number blocks of min(rows, columns).  Cached on the labels (same bond, sweep after sweep)."""
    key = (qnbigl.tobytes(), qnbigr.tobytes(), np.asarray(qntot).tobytes(), qnbigl.shape, qnbigr.shape)
    if key not in _rank_cache:
        if len(_rank_cache) > 512:
            _rank_cache.clear()
        _rank_cache[key] = _economic_rank(qnbigl, qnbigr, qntot)
    return _rank_cache[key]


def _economic_rank(qnbigl, qnbigr, qntot):
    nq = qnbigl.shape[-1]
    lq, lc = np.unique(qnbigl.reshape(-1, nq), axis=0, return_counts=True)
    rq, rc = np.unique(np.asarray(qntot).reshape(1, -1) - qnbigr.reshape(-1, nq), axis=0, return_counts=True)
    right = {tuple(q): c for q, c in zip(rq, rc)}
    return int(sum(min(c, right.get(tuple(q), 0)) for q, c in zip(lq, lc)))


def _distinct_qn(lqn):
    """The reference iterates `set([tuple(t) for t in lqn])` (svd_qn.py:140); a set's layout depends
    only on the sequence of DISTINCT insertions, so inserting the first occurrences in order gives
    the same iteration order without a Python loop over every row."""
    if lqn.shape[0] <= 64:
        return set([tuple(t) for t in lqn])
    _, first = np.unique(lqn, axis=0, return_index=True)
    return set([tuple(t) for t in lqn[np.sort(first)]])


def _idx(a, device):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(device)


def _complete(q, extra):
    """Append `extra` orthonormal columns orthogonal to the orthonormal columns of q.
    Same construction as add_orthonormal_basis (svd_qn.py:52-66): project random vectors and QR."""
    m, n = q.shape
    if extra <= 0:
        return q
    # uniform random vectors as in the reference; drawn on the device from a generator seeded by numpy's
    # global stream, so that np.random.seed still fixes the result (and every rank of a multi-GPU sweep
    # draws the same vectors) without generating and uploading m x extra numbers on the host
    gen = torch.Generator(device=q.device)
    gen.manual_seed(int(np.random.randint(0, 2 ** 31 - 1)))
    a = torch.rand((m, extra), generator=gen, dtype=torch.float64, device=q.device).to(q.dtype)
    for _ in range(2):
        if n > 0:
            a = a - ops.matmul(q, ops.matmul(q.conj().transpose(0, 1).contiguous(), a))
    qa, _ = ops.qr(a)
    return torch.cat([q, qa], dim=1)


def _numerical_rank(sh, shape):
    """Singular values above the cutoff of numpy.linalg.matrix_rank (with a margin for the round-off
    of the two QR factorisations in front of the Jacobi iteration); sh sorted descending (host)."""
    if len(sh) == 0:
        return 0
    tol = sh.max() * max(shape) * np.finfo(float).eps * 8
    return int(np.count_nonzero(sh > tol))


def _orthonormal_null_vectors(u, s, vh, sh=None):
    """One-sided Jacobi delivers the left vectors as (A V)_j / sigma_j: for a (numerically) zero
    singular value that column is zero or normalised round-off, not a unit vector orthogonal to the
    others as LAPACK's would be.  The sweeps do keep such vectors (bond dimension larger than the
    rank), so they are replaced by an orthonormal completion of the well-defined ones, exactly the
    construction the reference uses for its own null-space columns (svd_qn.py:52-66)."""
    k = s.numel()
    if k == 0:
        return u, vh
    if sh is None:
        sh = s.cpu().numpy()
    ngood = _numerical_rank(sh, (u.shape[0], vh.shape[1]))          # s is sorted: the null vectors come last
    if ngood == k:
        return u, vh
    u = _complete(u[:, :ngood].contiguous(), k - ngood)
    vh = _complete(vh[:ngood].conj().transpose(0, 1).contiguous(), k - ngood).conj().transpose(0, 1).contiguous()
    return u, vh


# The quantum-number blocks of one bond are independent (svd_qn.py:140-200 loops over them).  One
# block's Jacobi iteration is a chain of small launches that leaves most of the SMs idle and it
# synchronises with the host once per sweep, so the economic SVDs of the larger blocks run
# concurrently: one host thread and one CUDA stream per block (ctypes releases the GIL during the C
# call).  Everything that draws random numbers (null-space completion) stays on the calling thread,
# in block order, so results do not depend on the scheduling.
SVD_CONCURRENT_MIN = 64          # blocks with min(m, n) below this are not worth a thread
_svd_pool = {"executor": None, "streams": {}}


def _economic_svds(blocks):
    """ops.svd of every block; blocks with min(shape) >= SVD_CONCURRENT_MIN run concurrently."""
    big = [i for i, b in enumerate(blocks) if min(b.shape) >= SVD_CONCURRENT_MIN]
    if len(big) < 2:
        return [ops.svd(b) for b in blocks]
    from . import parallel
    group = parallel.heff_group()
    if group is not None:
        return _economic_svds_distributed(blocks, big, group)
    from concurrent.futures import ThreadPoolExecutor
    nworker = min(len(big), 4)
    if _svd_pool["executor"] is None:
        _svd_pool["executor"] = ThreadPoolExecutor(max_workers=4, thread_name_prefix="rn_svd")
    dev = blocks[0].device
    streams = _svd_pool["streams"].setdefault(dev.index, [torch.cuda.Stream(device=dev) for _ in range(4)])
    main = torch.cuda.current_stream(dev)
    ready = torch.cuda.Event()
    ready.record(main)
    results = [None] * len(blocks)

    def work(slot, idxs):
        torch.cuda.set_device(dev)
        st = streams[slot]
        with torch.cuda.stream(st):
            st.wait_event(ready)
            for i in idxs:
                results[i] = ops.svd(blocks[i])
        done = torch.cuda.Event()
        done.record(st)
        return done

    # largest blocks first, dealt round-robin to the workers
    order = sorted(big, key=lambda i: -blocks[i].shape[0] * blocks[i].shape[1] * min(blocks[i].shape))
    futures = [_svd_pool["executor"].submit(work, w, order[w::nworker]) for w in range(nworker)]
    for i, b in enumerate(blocks):
        if i not in big:
            results[i] = ops.svd(b)
    for f in futures:
        main.wait_event(f.result())
    for i in big:
        for t in results[i]:
            t.record_stream(main)
    return results


def _economic_svds_distributed(blocks, big, group):
    """One sweep on several GPUs (parallel.enable_sharded_heff): the larger blocks are dealt to the ranks
    of the group (largest first, to the least loaded rank), every rank factorises its own and the
    factors are broadcast, so that all ranks continue with the same bits."""
    import torch.distributed as dist
    grp = None if group is True else group
    rank, world = dist.get_rank(grp), dist.get_world_size(grp)
    cost = {i: float(blocks[i].shape[0]) * blocks[i].shape[1] * min(blocks[i].shape) for i in big}
    load, owner = [0.0] * world, {}
    for i in sorted(big, key=lambda i: -cost[i]):
        r = min(range(world), key=lambda q: load[q])
        owner[i] = r
        load[r] += cost[i]
    results = [None] * len(blocks)
    for i, b in enumerate(blocks):
        if i not in owner or owner[i] == rank:
            results[i] = ops.svd(b)
    for i in big:                                           # fixed order on every rank
        m, n = blocks[i].shape
        k = min(m, n)
        if owner[i] == rank:
            u, sv, vh = (t.contiguous() for t in results[i])
        else:
            u = torch.empty((m, k), dtype=blocks[i].dtype, device=blocks[i].device)
            sv = torch.empty((k,), dtype=torch.float64, device=blocks[i].device)
            vh = torch.empty((k, n), dtype=blocks[i].dtype, device=blocks[i].device)
        src = owner[i] if grp is None else dist.get_global_rank(grp, owner[i])
        for t in (u, sv, vh):
            dist.broadcast(torch.view_as_real(t) if t.is_complex() else t, src=src, group=grp)
        results[i] = (u, sv, vh)
    return results


def _block_svd(block, full_matrices, opt_full_matrices, economic=None, sh=None, fix_null=True):
    """optimized_svd (svd_qn.py:13-49): economic Jacobi SVD, completed to the sizes the reference
    returns under full_matrices (all of the null space, or min(m,n) extra vectors when very
    unbalanced).  `economic`: the block's (u, s, vh) when it has been computed already, `sh` its
    singular values on the host; fix_null=False leaves the vectors of (numerically) zero singular
    values as the Jacobi iteration returned them (the caller knows they cannot be selected)."""
    m, n = block.shape
    if not full_matrices:
        opt_full_matrices = False
    opt = opt_full_matrices and not (1 / 3 < m / n < 3)
    u, s, vh = ops.svd(block) if economic is None else economic
    if fix_null or full_matrices:
        u, vh = _orthonormal_null_vectors(u, s, vh, sh)
    if full_matrices:
        k = min(m, n)
        if opt:
            if m < n:
                vh = _complete(vh.conj().transpose(0, 1).contiguous(), k).conj().transpose(0, 1).contiguous()
            else:
                u = _complete(u, k)
        else:
            u = _complete(u, m - k)
            vh = _complete(vh.conj().transpose(0, 1).contiguous(), n - k).conj().transpose(0, 1).contiguous()
    return u, s, vh


_structure_cache = {}


def _block_structure(lqn, rqn, qntot, nl, nr, dev):
    """The quantum-number blocks of a bond matrix: [(ql, qr, row indices, column indices, trivial)] in
    the reference's iteration order (svd_qn.py:176-186), the index arrays already on the device.  The
    structure of a bond repeats whenever its labels do (fixed-basis sweeps, repeated H_eff set-ups), so
    it is cached on the quantum-number labels."""
    key = (lqn.tobytes(), rqn.tobytes(), np.asarray(qntot).tobytes(), lqn.shape, rqn.shape, str(dev))
    hit = _structure_cache.get(key)
    if hit is not None:
        return hit
    out = []
    for ql in _distinct_qn(lqn):
        qr_ = qntot - ql
        rset = np.where(get_qn_mask(rqn, qr_))[0]
        if len(rset) == 0:
            continue
        lset = np.where(get_qn_mask(lqn, ql))[0]
        trivial = len(lset) == nl and len(rset) == nr
        li, ri = (None, None) if trivial else (_idx(lset, dev), _idx(rset, dev))
        out.append((ql, qr_, li, ri, trivial))
    if len(_structure_cache) > 512:
        _structure_cache.clear()
    _structure_cache[key] = out
    return out


def _scatter_rows(indices, block, nrows, trivial):
    if trivial:
        return block
    out = torch.zeros((nrows, block.shape[1]), dtype=block.dtype, device=block.device)
    out.index_copy_(0, indices, block)
    return out


def svd_qn(coef_array, qnbigl: np.ndarray, qnbigr: np.ndarray, qntot: np.ndarray, QR: bool = False,
           system: str = None, full_matrices: bool = True, opt_full_matrices: bool = True, keep_hint: int = None):
    """Block decompose the coefficient tensor by SVD / QR according to the quantum numbers.

    Returns (U, S_u, qnl, V, S_v, qnr) -- or (U, qnl, V, qnr) with QR=True -- exactly like the
    reference (svd_qn.py:97-246): U and V are device tensors whose columns are the new basis
    (V = Vh.T, not conjugated), S_* are NumPy arrays.  With QR=True and system "R" the factor
    returned in U is lower instead of upper triangular (an LQ factorisation); the orthonormal
    factor and the product U @ V.T are the same gauge class as the reference's RQ.

    keep_hint (not in the reference): the largest number of vectors the caller will retain, ranked by
    singular value.  When the blocks hold at least that many vectors with non-zero singular values, the
    vectors of the zero singular values can never be selected and their orthonormal completion
    (random vectors, two projections and a QR per block) is skipped.
    """
    coef_array = asxp(coef_array)
    nl = int(np.prod(qnbigl.shape[:-1]))
    nr = int(np.prod(qnbigr.shape[:-1]))
    mat = coef_array.reshape(nl, nr)
    dev = mat.device
    assert qntot.ndim == 1
    qn_size = len(qntot)
    lqn = qnbigl.reshape(-1, qn_size)
    rqn = qnbigr.reshape(-1, qn_size)

    u_nz, u_z, v_nz, v_z, s_nz, su_z, sv_z = [], [], [], [], [], [], []
    ql_nz, ql_z, qr_nz, qr_z = [], [], [], []
    if QR and not full_matrices and not (lqn.any() or rqn.any() or qntot.any()):
        # no conserved quantum number: one block, nothing to gather or scatter (the general code below
        # produces exactly this; the short cut keeps the host out of the way between two kernels)
        if system == "R":
            bu, bq = ops.qr(mat.contiguous(), lq=True)
            bv = bq.transpose(0, 1)
        elif system == "L":
            bu, br = ops.qr(mat.contiguous())
            bv = br.transpose(0, 1)
        else:
            assert False
        zero = (0,) * qn_size
        dim = min(nl, nr)
        return bu, [zero] * dim, bv, [zero] * dim
    todo = []
    for ql, qr_, li, ri, trivial in _block_structure(lqn, rqn, qntot, nl, nr, dev):
        block = mat if trivial else mat.index_select(0, li).index_select(1, ri).contiguous()
        todo.append((ql, qr_, li, ri, trivial, block))
    economic = _economic_svds([t[-1] for t in todo]) if not QR else [None] * len(todo)
    s_host = [eco[1].cpu().numpy() for eco in economic] if not QR else [None] * len(todo)
    fix_null = True
    if not QR and keep_hint is not None and not full_matrices:
        fix_null = sum(_numerical_rank(sh, t[-1].shape) for sh, t in zip(s_host, todo)) < keep_hint
    for (ql, qr_, li, ri, trivial, block), eco, sh in zip(todo, economic, s_host):
        dim = min(block.shape)
        if not QR:
            bu, bs, bvh = _block_svd(block, full_matrices, opt_full_matrices, economic=eco, sh=sh, fix_null=fix_null)
            s_nz.append(sh)
            bv = bvh.transpose(0, 1)
        else:
            if full_matrices:
                raise NotImplementedError("QR with full_matrices=True is not used on the sweep path")
            if system == "R":
                bu, bq = ops.qr(block, lq=True)
                bv = bq.transpose(0, 1)
            elif system == "L":
                bu, br = ops.qr(block)
                bv = br.transpose(0, 1)
            else:
                assert False
        u_nz.append(_scatter_rows(li, bu[:, :dim].contiguous(), nl, trivial))
        ql_nz += [ql] * dim
        v_nz.append(_scatter_rows(ri, bv[:, :dim].contiguous(), nr, trivial))
        qr_nz += [tuple(qr_)] * dim
        if full_matrices:
            u_z.append(_scatter_rows(li, bu[:, dim:].contiguous(), nl, trivial))
            ql_z += [ql] * (bu.shape[1] - dim)
            su_z.append(np.zeros(bu.shape[1] - dim))
            v_z.append(_scatter_rows(ri, bv[:, dim:].contiguous(), nr, trivial))
            qr_z += [tuple(qr_)] * (bv.shape[1] - dim)
            sv_z.append(np.zeros(bv.shape[1] - dim))
    if len(u_nz) + len(u_z) == 0 or len(v_nz) + len(v_z) == 0:
        raise ValueError("Invalid quantum number")
    u = torch.cat(u_nz + u_z, dim=1) if len(u_nz + u_z) > 1 else (u_nz + u_z)[0]
    v = torch.cat(v_nz + v_z, dim=1) if len(v_nz + v_z) > 1 else (v_nz + v_z)[0]
    qnl_new = ql_nz + ql_z
    qnr_new = qr_nz + qr_z
    if QR:
        return u, qnl_new, v, qnr_new
    su = np.concatenate(s_nz + su_z)
    sv = np.concatenate(s_nz + sv_z)
    if not full_matrices:
        order = np.argsort(su)[::-1]
        oi = _idx(order.copy(), dev)
        u, v = u.index_select(1, oi), v.index_select(1, oi)
        su = sv = su[order]
        qnl_new = np.array(qnl_new)[order].tolist()
        qnr_new = np.array(qnr_new)[order].tolist()
    return u, su, qnl_new, v, sv, qnr_new


def eigh_qn(dm, qnbigl, qnbigr, qntot, system):
    """Block diagonalisation of the averaged reduced density matrix of the multi-state algorithm
    (svd_qn.py:243-302).  The blocks are Hermitian positive semi-definite, so their Jacobi SVD is
    their eigen-decomposition.  Returns (U device tensor, sqrt(eigenvalues) NumPy, new qn list)."""
    assert system in ("L", "R")
    qnbig, comp = (qnbigl, qnbigr) if system == "L" else (qnbigr, qnbigl)
    qn_size = len(qntot)
    localqn = qnbig.reshape(-1, qn_size)
    compqn = comp.reshape(-1, qn_size)
    n = len(localqn)
    dm = asxp(dm).reshape(n, n)
    dev = dm.device
    us, ss, new_qn = [], [], []
    for nl in _distinct_qn(localqn):
        nr = qntot - nl
        if np.sum(get_qn_mask(compqn, nr)) == 0:
            continue
        lset = np.where(get_qn_mask(localqn, nl))[0]
        trivial = len(lset) == n
        li = None if trivial else _idx(lset, dev)
        block = dm if trivial else dm.index_select(0, li).index_select(1, li).contiguous()
        bu, bs, bvh = ops.svd(block)
        bu, _ = _orthonormal_null_vectors(bu, bs, bvh)
        s2 = bs.cpu().numpy()
        s2[s2 < 0] = 0
        ss.append(np.sqrt(s2))
        us.append(_scatter_rows(li, bu.contiguous(), n, trivial))
        new_qn += [nl] * len(lset)
    u = torch.cat(us, dim=1) if len(us) > 1 else us[0]
    return u, np.concatenate(ss), new_qn


def select_basis(vset, sset, qnlist, compset, Mmax, percent=0):
    """Select the retained basis and the complementary tensor (reference lib.py:265-335).
    Index selection runs on the host over the singular values; the column gathers run on device."""
    qnlist = [tuple(qn) for qn in qnlist]
    qnset = set(qnlist)
    pool = {i: (qnlist[i], sset[i]) for i in range(len(qnlist))}

    def block_select(qn, n):
        members = sorted(((i, v) for i, v in pool.items() if v[0] == qn),
                         key=lambda x: x[1][1], reverse=True)
        got = [i for i, _ in members[:min(n, len(members))]]
        for i in got:
            del pool[i]
        return got

    nbasis = min(len(pool), Mmax)
    sidx = []
    if percent != 0:
        nbas_block = int(nbasis * percent / len(qnset))
        for iqn in qnset:
            sidx += block_select(iqn, nbas_block)
    nbasis = nbasis - len(sidx)
    ranked = sorted(pool.items(), key=lambda x: x[1][1], reverse=True)
    sidx += [i for i, _ in ranked[:nbasis]]
    assert len(sidx) == len(set(sidx))
    mpsdim = len(sidx)
    dev = vset.device
    sel = _idx(np.array(sidx), dev)
    ms = vset.index_select(1, sel)
    compmps = None
    if compset is not None:
        ncomp = compset.shape[1]
        inside = np.array([i < ncomp for i in sidx])
        scale = np.where(inside, np.asarray(sset)[sidx], 0.0)
        safe = _idx(np.where(inside, np.array(sidx), 0), dev)
        compmps = compset.index_select(1, safe) * torch.from_numpy(scale).to(dev).to(compset.dtype)
    mpsqn = [qnlist[i] for i in sidx]
    return ms, mpsdim, np.array(mpsqn), compmps

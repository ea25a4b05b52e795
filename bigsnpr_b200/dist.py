"""Multi-GPU host layer: SNP columns sharded over the ranks of one node.

One process per GPU (torchrun), ``torch.distributed`` for the plumbing.  The only exchange on the path is the
sum of the n-vector of partial products after a column-sharded X.y (and, inside the SVD, after every
A (A^T x)): one all-reduce of n doubles per product -- NCCL over NVLink on the GPUs, gloo in the CPU tests.
Xt.y needs no collective: every rank owns a disjoint slice of the result.

The reduction logic is written against a small "local operator" protocol so the same code runs over the GPU
engine (``LocalGpu``) and, in the CPU tests, over any stand-in with the same two methods.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def shard_bounds(m: int, world: int, rank: int):
    """Contiguous column range [begin, end) of `rank`: the first (m % world) ranks get one more column."""
    base, rem = divmod(int(m), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class ShardedMatVec:
    """X~ = [X~_0 | X~_1 | ...] by columns.  `local` implements prodvec(x_local) -> partial (n,) and
    cprodvec(y) -> (m_local,) on torch tensors living on the device of the process group's backend."""

    def __init__(self, local, m_total: int, group=None):
        import torch.distributed as dist

        self.local, self.group = local, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.begin, self.end = shard_bounds(m_total, self.world, self.rank)
        self.m_total = m_total

    def prodvec(self, x_local):
        """X~ x, with x given as this rank's slice; returns the full n-vector on every rank."""
        import torch.distributed as dist

        out = self.local.prodvec(x_local)
        if self.world > 1:
            dist.all_reduce(out, group=self.group)
        return out

    def cprodvec(self, y):
        """t(X~) y: this rank's slice of the m-vector (no collective)."""
        return self.local.cprodvec(y)

    def cprodvec_gathered(self, y):
        """t(X~) y gathered on every rank (for callers that want the whole vector)."""
        import torch
        import torch.distributed as dist

        mine = self.local.cprodvec(y)
        if self.world == 1:
            return mine
        sizes = [e - b for b, e in (shard_bounds(self.m_total, self.world, r) for r in range(self.world))]
        pad = torch.zeros(max(sizes), dtype=mine.dtype, device=mine.device)
        pad[: mine.numel()] = mine
        parts = [torch.empty_like(pad) for _ in sizes]
        dist.all_gather(parts, pad, group=self.group)
        return torch.cat([t[:k] for t, k in zip(parts, sizes)])


class LocalGpu:
    """Local operator over a bigsnpr_b200 View: device-resident torch vectors, no host copies."""

    def __init__(self, view, device):
        import torch

        self.view, self.device = view, device
        self.stream = torch.cuda.current_stream(device).cuda_stream

    def prodvec(self, x):
        import torch

        out = torch.empty(self.view.nr, dtype=torch.float64, device=self.device)
        self.view.prodvec_dev(x.data_ptr(), out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream or 0)
        return out

    def cprodvec(self, y):
        import torch

        out = torch.empty(self.view.nc, dtype=torch.float64, device=self.device)
        self.view.cprodvec_dev(y.data_ptr(), out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream or 0)
        return out


class Comm:
    """This rank's end of the library's own communicator (bsg_comm): the ranks of one node exchange data through each
    other's HBM over NVLink (CUDA IPC mappings).  The only thing that crosses the host is the 64-byte IPC handle, gathered
    here once with torch.distributed; afterwards the X.y epilogue kernel itself sums the shards' partial vectors."""

    def __init__(self, max_elems, device=None, group=None):
        import torch
        import torch.distributed as dist

        from . import _lib

        self._c = None
        L = _lib.lib()
        world, rank = _world(group)
        dev = torch.cuda.current_device() if device is None else int(device)
        c = C.c_void_p()
        mine = np.zeros(64, dtype=np.uint8)
        _lib.check(L.bsg_comm_create(rank, world, dev, int(max_elems), C.byref(c), mine.ctypes.data_as(_lib.c_u8_p)))
        self._c, self.rank, self.world = c, rank, world
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, mine.tobytes(), group=group)
            allh = np.frombuffer(b"".join(handles), dtype=np.uint8).copy()
            _lib.check(L.bsg_comm_connect(c, allh.ctypes.data_as(_lib.c_u8_p)))
            dist.barrier(group=group)  # every rank has mapped every region before the first collective touches one

    def prodvec_allreduce(self, view, x_ptr, out_ptr, stream=0):
        """X~ x over this rank's shard, summed over the ranks inside the epilogue kernel; out = the full n-vector."""
        from . import _lib

        _lib.check(_lib.lib().bsg_view_prodvec_allreduce_dev(view._v, self._c, int(x_ptr), int(out_ptr), int(stream) or None))

    def allreduce(self, buf_ptr, count, stream=0):
        from . import _lib

        _lib.check(_lib.lib().bsg_comm_allreduce_dev(self._c, int(buf_ptr), int(count), int(stream) or None))

    def check(self):
        from . import _lib

        _lib.check(_lib.lib().bsg_comm_check(self._c))

    def close(self):
        if self._c is not None:
            from . import _lib

            _lib.lib().bsg_comm_destroy(self._c)
            self._c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def randomsvd_comm(obj_bed, comm, m_total, k=10, tol=1e-4, maxit=1000):
    """bed_randomSVD on a column-sharded matrix through the library's communicator: every rank passes its shard; u (n x k)
    and d are replicated (bit-identical on all ranks), v holds this rank's rows.  No host synchronisation inside the
    iteration except one read of the (ncv x ncv) projected matrix per restart."""
    from . import _lib

    L = _lib.lib()
    n, m_loc = obj_bed.nrow, obj_bed.ncol
    d = np.empty(k)
    u = np.empty((k, n))
    v = np.empty((k, m_loc))
    c_out, s_out = np.empty(m_loc), np.empty(m_loc)
    niter, nops = C.c_int(0), C.c_int(0)
    pd = lambda a: a.ctypes.data_as(_lib.c_dbl_p)  # noqa: E731
    _lib.check(L.bsg_randomsvd_comm(obj_bed._h, comm._c, None, n, None, m_loc, None, None, int(m_total), int(k), float(tol),
                                    int(maxit), pd(d), pd(u), pd(v), pd(c_out), pd(s_out), C.byref(niter), C.byref(nops)))
    return {"d": d, "u": u.T, "v": v.T, "niter": niter.value, "nops": nops.value, "center": c_out, "scale": s_out}


def randomsvd_sharded(obj_bed, m_total, k=10, tol=1e-4, maxit=1000, group=None):
    """Same decomposition with the caller's collective: the Lanczos iteration is the library's (bsg_randomsvd_ex); the
    reduce callback all-reduces the n-vector of partial products with torch.distributed (NCCL, or gloo in CPU tests of
    the host logic).  Kept as the baseline the communicator form is measured against."""
    import torch
    import torch.distributed as dist

    from . import _lib

    L = _lib.lib()
    n, m_loc = obj_bed.nrow, obj_bed.ncol
    dev = torch.device("cuda", torch.cuda.current_device())
    z = torch.zeros(n, dtype=torch.float64, device=dev)

    @C.CFUNCTYPE(None, C.c_void_p)
    def reduce_cb(_ctx):
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(z, group=group)
        torch.cuda.synchronize(dev)

    d = np.empty(k)
    u = np.empty((k, n))
    v = np.empty((k, m_loc))
    c_out, s_out = np.empty(m_loc), np.empty(m_loc)
    niter, nops = C.c_int(0), C.c_int(0)
    pd = lambda a: a.ctypes.data_as(_lib.c_dbl_p)  # noqa: E731
    _lib.check(L.bsg_randomsvd_ex(obj_bed._h, None, n, None, m_loc, None, None, int(k), float(tol), int(maxit), pd(d),
                                  pd(u), pd(v), pd(c_out), pd(s_out), C.byref(niter), C.byref(nops), z.data_ptr(),
                                  C.cast(reduce_cb, C.c_void_p), None, int(m_total)))
    return {"d": d, "u": u.T, "v": v.T, "niter": niter.value, "nops": nops.value, "center": c_out, "scale": s_out}


# ---------------------------------------------------------------------------------------------------------
# The other rows of SURVEY.md section 8e.  Everything below is host-side composition: the per-shard work is one
# C-ABI call on the rank's own handle; what crosses ranks is stated per function.
# ---------------------------------------------------------------------------------------------------------
def _world(group=None):
    import torch.distributed as dist

    if not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def _coll_device(group=None):
    """Device collectives run on: the current CUDA device under NCCL, the CPU under gloo."""
    import torch
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def gather_columns(local, m_total, group=None):
    """Concatenate per-rank column slices (last axis = this rank's columns, shard_bounds order) on every rank.
    Used for colstats / column counts / Xt.y: results are disjoint slices, so this is a gather, not a reduction."""
    import torch
    import torch.distributed as dist

    world, _ = _world(group)
    local = np.ascontiguousarray(local)
    if world == 1:
        return local
    dev = _coll_device(group)
    lead = local.shape[:-1]
    sizes = [e - b for b, e in (shard_bounds(m_total, world, r) for r in range(world))]
    width = max(sizes)  # shards differ by at most one column: pad to equal size, trim after the gather
    mine = torch.zeros((width,) + lead, dtype=torch.from_numpy(local).dtype, device=dev)
    mine[: local.shape[-1]] = torch.from_numpy(np.moveaxis(local, -1, 0).copy()).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    out = torch.cat([t[:k] for t, k in zip(parts, sizes)])
    return np.moveaxis(out.cpu().numpy(), 0, -1)


def sum_over_ranks(local, group=None):
    """Element-wise sum of equally-shaped arrays (by-row counts: 4 x n int32 partial counts per column shard)."""
    import torch
    import torch.distributed as dist

    world, _ = _world(group)
    if world == 1:
        return np.asarray(local)
    t = torch.from_numpy(np.ascontiguousarray(local)).to(_coll_device(group))
    dist.all_reduce(t, group=group)
    return t.cpu().numpy()


def sharded_colstats(local_stats, m_total, group=None):
    """bed_colstats over column shards: every field is per column -> gather."""
    return {k: gather_columns(np.asarray(v), m_total, group) for k, v in local_stats.items()}


def sharded_counts(local_counts, m_total, byrow=False, group=None):
    """bed_counts over column shards.  By column: gather of the 4 x m_local blocks.  By row: each rank counted
    its own columns for every sample -> all-reduce (sum) of 4 x n integers."""
    if byrow:
        return sum_over_ranks(np.asarray(local_counts, dtype=np.int64), group).astype(np.int32)
    return gather_columns(np.asarray(local_counts), m_total, group)


def tcrossprod_sharded(obj_bed, center, scale, ind_row=None, group=None):
    """GRM over column shards: K = sum_g X~_g X~_g^T.  Each rank accumulates its K_g on the device
    (bsg_tcrossprod_dev, no host copy) and one all-reduce of n^2 doubles sums them in place over NVLink.
    Returns the n x n torch tensor on this rank's GPU (replicated)."""
    import torch
    import torch.distributed as dist

    from . import _lib

    L = _lib.lib()
    n = obj_bed.nrow if ind_row is None else len(ind_row)
    dev = torch.device("cuda", torch.cuda.current_device())
    K = torch.empty((n, n), dtype=torch.float64, device=dev)
    center = np.ascontiguousarray(center, dtype=np.float64)
    scale = np.ascontiguousarray(scale, dtype=np.float64)
    ir = None if ind_row is None else np.ascontiguousarray(ind_row, dtype=np.int32)
    _lib.check(L.bsg_tcrossprod_dev(obj_bed._h, None if ir is None else ir.ctypes.data_as(_lib.c_int_p), n, None,
                                    obj_bed.ncol, center.ctypes.data_as(_lib.c_dbl_p),
                                    scale.ctypes.data_as(_lib.c_dbl_p), K.data_ptr()))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(K, group=group)
    return K


def halo_bounds(pos, size_bp, begin, end, right=False):
    """Columns a shard [begin, end) has to see for the windowed pair statistics.

    The reference pairs j < j0 iff pos[j] >= pos[j0] - size (src/corr.cpp:44-46, src/ld-scores.cpp:36-38).
    Left halo (correlation columns begin..end-1 need their earlier partners): lo = first j with
    pos[j] >= pos[begin] - size.  Right halo (LD scores also collect the pairs in which a column is the EARLIER
    member): hi = one past the last j0 with pos[j0] - size <= pos[end-1].  Both use the same floating-point
    expression as the pair test, so the shard sees exactly the reference's pairs."""
    pos = np.asarray(pos, dtype=np.float64)
    if end <= begin:
        return begin, end
    lo = int(np.searchsorted(pos, pos[begin] - size_bp, side="left"))
    lo = min(lo, begin)
    hi = end
    if right:
        hi = max(end, int(np.searchsorted(pos - size_bp, pos[end - 1], side="right")))
    return lo, hi


def cor_sharded(cor_fn, pos, size_bp, m_total, group=None, gather=True):
    """Windowed correlation matrix over column shards, no data-path collective.

    `cor_fn(lo, hi)` returns the CSC pieces (p, i, x) of corMat restricted to global columns [lo, hi) (local
    0-based row indices).  The rank computes [lo, end) with the left halo, keeps the columns it owns and
    re-bases the row indices; with gather=True the pieces are concatenated on every rank (host objects, like the
    reference's list of columns)."""
    import torch.distributed as dist

    world, rank = _world(group)
    b, e = shard_bounds(m_total, world, rank)
    lo, _ = halo_bounds(pos, size_bp, b, e)
    p, i, x = cor_fn(lo, e)
    k0 = b - lo
    s0, s1 = int(p[k0]), int(p[-1])
    p_own = np.asarray(p[k0:], dtype=np.int64) - s0
    i_own = np.asarray(i[s0:s1], dtype=np.int64) + lo
    x_own = np.asarray(x[s0:s1], dtype=np.float64)
    if not gather or world == 1:
        return p_own, i_own, x_own
    pieces = [None] * world
    dist.all_gather_object(pieces, (p_own, i_own, x_own), group=group)
    P, I, X, off = [np.zeros(1, dtype=np.int64)], [], [], 0
    for pp, ii, xx in pieces:
        P.append(pp[1:] + off)
        off += int(pp[-1])
        I.append(ii)
        X.append(xx)
    return np.concatenate(P), np.concatenate(I), np.concatenate(X)


def ld_scores_sharded(ld_fn, pos, size_bp, m_total, group=None):
    """LD scores over column shards.  `ld_fn(lo, hi)` returns the scores of global columns [lo, hi) computed on
    that range alone.  With halos on both sides every pair a shard's column belongs to is inside the range, so the
    kept middle part is complete and the result is a gather (instead of the all-reduce of boundary terms)."""
    world, rank = _world(group)
    b, e = shard_bounds(m_total, world, rank)
    lo, hi = halo_bounds(pos, size_bp, b, e, right=True)
    ld = np.asarray(ld_fn(lo, hi), dtype=np.float64)
    return gather_columns(ld[b - lo:e - lo], m_total, group)


# ---------------------------------------------------------------------------------------------------------
# Arbitrary `ind.col` over column shards ("staging" row): the caller's global 1-based column
# multiset is bucketed by owning rank once per call; every rank then works on its own local indices.
# ---------------------------------------------------------------------------------------------------------
def bucket_columns(ind_col, m_total, world):
    """For every entry of the global 1-based `ind_col` (any order, duplicates allowed): the owning rank and the
    1-based index inside that rank's shard (shard_bounds).  Returns (owner, local) int arrays of ind_col's length."""
    ind_col = np.asarray(ind_col, dtype=np.int64)
    if ind_col.size and (ind_col.min() < 1 or ind_col.max() > m_total):
        raise IndexError("Tested subscript out of bounds (column not in 1..%d)." % m_total)
    begins = np.array([shard_bounds(m_total, world, r)[0] for r in range(world)], dtype=np.int64)
    owner = np.searchsorted(begins, ind_col - 1, side="right") - 1
    return owner.astype(np.int32), (ind_col - begins[owner]).astype(np.int32)


def prodvec_selected(local_prodvec, ind_col, x, m_total, group=None):
    """X~[, ind_col] x over column shards.  `local_prodvec(local_ind_col, x_part)` returns this rank's partial
    n-vector (a torch tensor on the collective's device) for its own columns of the multiset; the partials are
    summed by one all-reduce.  Ranks that own none of the columns contribute zeros."""
    import torch.distributed as dist

    world, rank = _world(group)
    owner, local = bucket_columns(ind_col, m_total, world)
    mine = owner == rank
    out = local_prodvec(local[mine], np.asarray(x, dtype=np.float64)[mine])
    if world > 1:
        dist.all_reduce(out, group=group)
    return out


def cprodvec_selected(local_cprodvec, ind_col, y, m_total, group=None):
    """t(X~[, ind_col]) y over column shards, in the caller's order.  `local_cprodvec(local_ind_col, y)` returns the
    values of this rank's columns of the multiset (numpy); every rank fills its positions of a zero vector and one
    all-reduce (a sum of disjoint supports, i.e. a gather) assembles the result."""
    world, rank = _world(group)
    owner, local = bucket_columns(ind_col, m_total, world)
    mine = owner == rank
    full = np.zeros(len(owner), dtype=np.float64)
    if mine.any():
        full[mine] = np.asarray(local_cprodvec(local[mine], y), dtype=np.float64)
    return sum_over_ranks(full, group)

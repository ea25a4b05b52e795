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
        sizes = [shard_bounds(self.m_total, self.world, r) for r in range(self.world)]
        parts = [torch.empty(e - b, dtype=mine.dtype, device=mine.device) for b, e in sizes]
        dist.all_gather(parts, mine, group=self.group)
        return torch.cat(parts)


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


def randomsvd_sharded(obj_bed, m_total, k=10, tol=1e-4, maxit=1000, group=None):
    """bed_randomSVD on a column-sharded matrix: every rank passes its shard handle; u (n x k) and d are
    replicated, v holds this rank's rows.  The Lanczos iteration is the library's (bsg_randomsvd_ex); the
    reduce callback all-reduces the n-vector of partial products with torch.distributed."""
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

"""World-size-2 gloo test of the multi-GPU host logic (bigsnpr_b200/dist.py) on CPU: column sharding, the
all-reduce after a sharded X.y, the collective-free Xt.y.  The local operator is a CPU stand-in built on the
oracle (tests may use it); the GPU run uses the same ShardedMatVec over libbsgpu views."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition():
    from bigsnpr_b200.dist import shard_bounds

    for m, w in ((10, 3), (1_100_000, 8), (7, 8), (500_000, 1)):
        cuts = [shard_bounds(m, w, r) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == m
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
        sizes = [e - b for b, e in cuts]
        assert max(sizes) - min(sizes) <= 1


class _OracleLocal:
    def __init__(self, ref, obed, cols, center, scale):
        self.ref, self.o, self.cols, self.c, self.s = ref, obed, cols, center, scale

    def prodvec(self, x):
        r = self.ref.bed_prodVec(self.o, x.numpy(), ind_col=self.cols, center=self.c, scale=self.s)
        return torch.from_numpy(r)

    def cprodvec(self, y):
        r = self.ref.bed_cprodVec(self.o, y.numpy(), ind_col=self.cols, center=self.c, scale=self.s)
        return torch.from_numpy(r)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bigsnpr_b200.dist import ShardedMatVec, shard_bounds
        from oracle import ref

        o = ref.OracleBed(os.path.join(ROOT, "tests", "golden", "example-missing.bed"))
        sc = ref.bed_scaleBinom(o)
        b, e = shard_bounds(o.ncol, world, rank)
        cols = np.arange(b + 1, e + 1, dtype=np.int32)
        op = ShardedMatVec(_OracleLocal(ref, o, cols, sc["center"][b:e], sc["scale"][b:e]), o.ncol)
        rng = np.random.default_rng(3)
        x, y = rng.normal(size=o.ncol), rng.normal(size=o.nrow)
        full = op.prodvec(torch.from_numpy(x[b:e].copy()))
        want = ref.bed_prodVec(o, x, center=sc["center"], scale=sc["scale"])
        e1 = float(np.max(np.abs(full.numpy() - want)) / np.max(np.abs(want)))
        mine = op.cprodvec(torch.from_numpy(y))
        wantc = ref.bed_cprodVec(o, y, center=sc["center"], scale=sc["scale"])
        e2 = float(np.max(np.abs(mine.numpy() - wantc[b:e])))
        allc = op.cprodvec_gathered(torch.from_numpy(y))
        e3 = float(np.max(np.abs(allc.numpy() - wantc)))
        q.put((rank, e1, e2, e3))
    finally:
        dist.destroy_process_group()


def test_sharded_matvec_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e1, e2, e3 in res:
        assert e1 < 1e-12 and e2 == 0.0 and e3 == 0.0, (rank, e1, e2, e3)

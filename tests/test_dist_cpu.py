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


# ---- the other section-8e rows: stats / counts / GRM / windowed correlation / LD scores over column shards ----
def _worker_rows(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bigsnpr_b200 import dist as D
        from oracle import ref

        o = ref.OracleBed(os.path.join(ROOT, "tests", "golden", "example-missing.bed"))
        m = o.ncol  # 500: not a multiple of the world size 3, shards differ by one column
        pos = np.cumsum(np.random.default_rng(5).integers(1, 4000, size=m)).astype(np.float64)
        b, e = D.shard_bounds(m, world, rank)
        cols = lambda lo, hi: np.arange(lo + 1, hi + 1, dtype=np.int32)  # noqa: E731
        errs = {}

        rows = np.arange(1, o.nrow + 1, dtype=np.int32)
        st = D.sharded_colstats(ref.bed_colstats(o, rows, cols(b, e)), m)
        want = ref.bed_colstats(o, rows, cols(0, m))
        errs["colstats"] = max(float(np.max(np.abs(np.asarray(st[k], dtype=float) - np.asarray(want[k], dtype=float))))
                               for k in want)
        cc = D.sharded_counts(ref.bed_counts(o, ind_col=cols(b, e)), m)
        errs["col_counts"] = int(np.max(np.abs(cc - ref.bed_counts(o, ind_col=cols(0, m)))))
        rc = D.sharded_counts(ref.bed_counts(o, ind_col=cols(b, e), byrow=True), m, byrow=True)
        errs["row_counts"] = int(np.max(np.abs(rc - ref.bed_counts(o, ind_col=cols(0, m), byrow=True))))

        # GRM: sum of the shards' K (the GPU path all-reduces the device tensor; same reduction here on CPU)
        Kl, _, _ = ref.bed_tcrossprodSelf(o, ind_col=cols(b, e))
        K = D.sum_over_ranks(Kl)
        Kw, _, _ = ref.bed_tcrossprodSelf(o, ind_col=cols(0, m))
        errs["grm"] = float(np.max(np.abs(K - Kw)) / np.max(np.abs(Kw)))

        size_kb = 30.0
        cor_fn = lambda lo, hi: ref.cor0(o, ind_col=cols(lo, hi), size=size_kb, alpha=0.5, infos_pos=pos[lo:hi])  # noqa: E731
        p, i, x = D.cor_sharded(cor_fn, pos, size_kb * 1000.0, m)
        pw, iw, xw = ref.cor0(o, ind_col=cols(0, m), size=size_kb, alpha=0.5, infos_pos=pos)
        same = np.array_equal(p, pw) and np.array_equal(i, iw) and np.array_equal(x, xw, equal_nan=True)
        errs["cor_identical"] = bool(same)
        errs["cor_nnz"] = int(pw[-1])

        ld_fn = lambda lo, hi: ref.ld0(o, ind_col=cols(lo, hi), size=size_kb, infos_pos=pos[lo:hi])  # noqa: E731
        ld = D.ld_scores_sharded(ld_fn, pos, size_kb * 1000.0, m)
        ldw = ref.ld0(o, ind_col=cols(0, m), size=size_kb, infos_pos=pos)
        errs["ld"] = float(np.nanmax(np.abs(ld - ldw)))
        lo, hi = D.halo_bounds(pos, size_kb * 1000.0, b, e, right=True)
        errs["halo"] = (b - lo, hi - e)

        # arbitrary global ind.col (unsorted, with duplicates) bucketed by owner
        rng = np.random.default_rng(11)
        gcols = rng.integers(1, m + 1, size=333)
        xs, ys = rng.normal(size=gcols.size), rng.normal(size=o.nrow)
        sc = ref.bed_scaleBinom(o)

        def lp(loc, xpart):
            gl = (loc + b).astype(np.int32)  # local 1-based -> global 1-based
            if gl.size == 0:
                return torch.zeros(o.nrow, dtype=torch.float64)
            return torch.from_numpy(ref.bed_prodVec(o, xpart, ind_col=gl, center=sc["center"][gl - 1], scale=sc["scale"][gl - 1]))

        def lc(loc, yv):
            gl = (loc + b).astype(np.int32)
            return ref.bed_cprodVec(o, yv, ind_col=gl, center=sc["center"][gl - 1], scale=sc["scale"][gl - 1])

        got = D.prodvec_selected(lp, gcols, xs, m).numpy()
        want = ref.bed_prodVec(o, xs, ind_col=gcols.astype(np.int32), center=sc["center"][gcols - 1], scale=sc["scale"][gcols - 1])
        errs["sel_prod"] = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
        gotc = D.cprodvec_selected(lc, gcols, ys, m)
        wantc = ref.bed_cprodVec(o, ys, ind_col=gcols.astype(np.int32), center=sc["center"][gcols - 1], scale=sc["scale"][gcols - 1])
        errs["sel_cprod"] = float(np.max(np.abs(gotc - wantc)))
        q.put((rank, errs))
    finally:
        dist.destroy_process_group()


def test_bucket_columns_by_owner():
    from bigsnpr_b200.dist import bucket_columns, shard_bounds

    m, world = 1001, 4
    cols = np.array([1, 1001, 251, 250, 252, 500, 501, 751, 1, 1000])
    owner, local = bucket_columns(cols, m, world)
    for c, r, l in zip(cols, owner, local):
        b, e = shard_bounds(m, world, r)
        assert b < c <= e and l == c - b
    with pytest.raises(IndexError):
        bucket_columns([0], m, world)
    with pytest.raises(IndexError):
        bucket_columns([m + 1], m, world)


def test_halo_bounds_cover_the_reference_window():
    from bigsnpr_b200.dist import halo_bounds

    rng = np.random.default_rng(2)
    pos = np.cumsum(rng.integers(0, 50, size=400)).astype(np.float64)
    size = 333.0
    for b, e in ((0, 100), (100, 250), (250, 400), (399, 400)):
        lo, hi = halo_bounds(pos, size, b, e, right=True)
        for j0 in range(b, e):  # every partner the reference pairs with a shard column is inside [lo, hi)
            part = [j for j in range(400) if (j < j0 and pos[j] >= pos[j0] - size) or (j > j0 and pos[j0] >= pos[j] - size)]
            assert all(lo <= j < hi for j in part)
        assert lo == 0 or not (pos[lo - 1] >= pos[b] - size)
        assert hi == 400 or not (pos[e - 1] >= pos[hi] - size)


def test_sharded_rows_gloo_world3():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 3
    procs = [ctx.Process(target=_worker_rows, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e in res:
        assert e["colstats"] == 0.0 and e["col_counts"] == 0 and e["row_counts"] == 0, (rank, e)
        assert e["grm"] < 1e-12, (rank, e)
        assert e["cor_identical"] and e["cor_nnz"] > 500, (rank, e)
        assert e["ld"] < 1e-10, (rank, e)
        assert e["sel_prod"] < 1e-12 and e["sel_cprod"] == 0.0, (rank, e)
    assert any(e["halo"][0] > 0 for _, e in res) and any(e["halo"][1] > 0 for _, e in res)

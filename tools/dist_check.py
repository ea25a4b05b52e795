#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, 1 rank per GPU): column-sharded bed_prodVec / bed_cprodVec / bed_randomSVD
against the single-GPU result on rank 0, then the other sharded rows (stats, counts, GRM, cor, LD scores).
Prints one JSON line per part on rank 0."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bigsnpr_b200 as B  # noqa: E402
from bigsnpr_b200 import dist as D  # noqa: E402
from bigsnpr_b200.dist import LocalGpu, ShardedMatVec, randomsvd_sharded, shard_bounds  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n, m, seed, k = 20000, 64000, 77, 10
    b, e = shard_bounds(m, world, rank)
    g = B.Bed.synthetic(n, e - b, seed=seed, na_rate=0.01, col_offset=b, device=local)
    sc = B.bed_scaleBinom(g)
    view = B.View(g, center=sc["center"], scale=sc["scale"])
    op = ShardedMatVec(LocalGpu(view, dev), m)
    rng = np.random.default_rng(5)
    x, y = rng.normal(size=m), rng.normal(size=n)
    Ax = op.prodvec(torch.from_numpy(x[b:e].copy()).to(dev))
    Aty = op.cprodvec_gathered(torch.from_numpy(y).to(dev))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    svd = randomsvd_sharded(g, m, k=k)
    t_svd = time.perf_counter() - t0
    # ---- the library's own communicator (NVLink peer memory, reduction fused into the X.y epilogue) on the DEFAULT stream
    comm = D.Comm(n, device=local)
    xd = torch.from_numpy(x[b:e].copy()).to(dev)
    outc = torch.empty(n, dtype=torch.float64, device=dev)
    comm.prodvec_allreduce(view, xd.data_ptr(), outc.data_ptr(), 0)
    torch.cuda.synchronize()
    comm_vs_nccl = float((outc - Ax).abs().max() / Ax.abs().max())
    gath = [torch.empty_like(outc) for _ in range(world)]
    dist.all_gather(gath, outc)
    comm_same_bits = all(bool(torch.equal(gath[0], t)) for t in gath)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tms = {}
    for name in ("comm", "nccl"):
        for it in range(25):
            if it == 5:
                ev0.record()
            if name == "comm":
                comm.prodvec_allreduce(view, xd.data_ptr(), outc.data_ptr(), 0)
            else:
                view.prodvec_dev(xd.data_ptr(), outc.data_ptr(), 0)
                dist.all_reduce(outc)
        ev1.record()
        torch.cuda.synchronize()
        tms[name] = ev0.elapsed_time(ev1) / 20
    t0 = time.perf_counter()
    svdc = D.randomsvd_comm(g, comm, m, k=k)
    t_svdc = time.perf_counter() - t0
    comm.check()
    res = {}
    if rank == 0:
        gf = B.Bed.synthetic(n, m, seed=seed, na_rate=0.01, device=local)
        scf = B.bed_scaleBinom(gf)
        vf = B.View(gf, center=scf["center"], scale=scf["scale"])
        Ax1, Aty1 = vf.prodvec(x), vf.cprodvec(y)
        t0 = time.perf_counter()
        svd1 = B.bed_randomSVD(gf, k=k)
        t1 = time.perf_counter() - t0
        res = {"world": world, "prodvec_max_rel": float(np.max(np.abs(Ax.cpu().numpy() - Ax1)) / np.max(np.abs(Ax1))),
               "cprodvec_max_rel": float(np.max(np.abs(Aty.cpu().numpy() - Aty1)) / np.max(np.abs(Aty1))),
               "svd_d_max_rel": float(np.max(np.abs(svd["d"] - svd1["d"]) / svd1["d"])),
               "svd_u_min_abs_corr": float(np.min(np.abs(np.sum(svd["u"] * svd1["u"], axis=0)))),
               "svd_sharded_s": t_svd, "svd_single_s": t1, "nops": svd["nops"], "nops_single": svd1["nops"],
               "comm_prodvec_vs_nccl_max_rel": comm_vs_nccl, "comm_same_bits_on_all_ranks": comm_same_bits,
               "prodvec_step_ms_comm": tms["comm"], "prodvec_step_ms_nccl": tms["nccl"],
               "svd_comm_d_max_rel": float(np.max(np.abs(svdc["d"] - svd1["d"]) / svd1["d"])),
               "svd_comm_u_min_abs_corr": float(np.min(np.abs(np.sum(svdc["u"] * svd1["u"], axis=0)))),
               "svd_comm_s": t_svdc, "nops_comm": svdc["nops"]}
        print(json.dumps(res), flush=True)
        ok = res["prodvec_max_rel"] < 1e-12 and res["cprodvec_max_rel"] < 1e-12 and res["svd_d_max_rel"] < 1e-7 \
            and res["svd_u_min_abs_corr"] > 1 - 1e-6 and res["comm_prodvec_vs_nccl_max_rel"] < 1e-13 \
            and res["comm_same_bits_on_all_ranks"] and res["svd_comm_d_max_rel"] < 1e-7 \
            and res["svd_comm_u_min_abs_corr"] > 1 - 1e-6
        if not ok:
            print("DIST CHECK FAILED", flush=True)
    dist.barrier()
    comm.close()
    rows_check(rank, world, local)
    dist.barrier()
    dist.destroy_process_group()


def rows_check(rank, world, local):
    """The other section-8e rows on the GPUs: stats / counts (gather, all-reduce), GRM (device all-reduce of n^2
    doubles), windowed correlation and LD scores (halo columns, no data-path collective) against one GPU."""
    n, m, seed, size_kb = 6000, 30001, 91, 150.0
    pos = 1000.0 * np.arange(1, m + 1)
    b, e = shard_bounds(m, world, rank)
    lo, hi = D.halo_bounds(pos, size_kb * 1000.0, b, e, right=True)
    g = B.Bed.synthetic(n, hi - lo, seed=seed, na_rate=0.01, col_offset=lo, device=local)  # shard + halos
    own = np.arange(b - lo + 1, e - lo + 1, dtype=np.int32)  # 1-based local indices of the owned columns
    t = {}
    t0 = time.perf_counter()
    st = D.sharded_colstats(B.bed_colstats(g, ind_col=own), m)
    cc = D.sharded_counts(B.bed_counts(g, ind_col=own), m)
    rc = D.sharded_counts(B.bed_counts(g, ind_col=own, byrow=True), m, byrow=True)
    t["stats_counts_s"] = time.perf_counter() - t0
    sc = B.bed_scaleBinom(g, ind_col=own)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gs = B.Bed.synthetic(n, e - b, seed=seed, na_rate=0.01, col_offset=b, device=local)
    t0 = time.perf_counter()
    K = D.tcrossprod_sharded(gs, sc["center"], sc["scale"])
    torch.cuda.synchronize()
    t["grm_sharded_s"] = time.perf_counter() - t0
    K = K.cpu().numpy()
    gs.close()

    def cor_fn(l, h):
        return B.bed_cor(g, ind_col=np.arange(l - lo + 1, h - lo + 1, dtype=np.int32), size=size_kb, alpha=0.05,
                         infos_pos=pos[l:h])

    def ld_fn(l, h):
        return B.bed_ld_scores(g, ind_col=np.arange(l - lo + 1, h - lo + 1, dtype=np.int32), size=size_kb,
                               infos_pos=pos[l:h])

    t0 = time.perf_counter()
    p, i, x = D.cor_sharded(cor_fn, pos, size_kb * 1000.0, m)
    t["cor_sharded_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    ld = D.ld_scores_sharded(ld_fn, pos, size_kb * 1000.0, m)
    t["ld_sharded_s"] = time.perf_counter() - t0
    if rank == 0:
        gf = B.Bed.synthetic(n, m, seed=seed, na_rate=0.01, device=local)
        st1, cc1, rc1 = B.bed_colstats(gf), B.bed_counts(gf), B.bed_counts(gf, byrow=True)
        t0 = time.perf_counter()
        K1, _, _ = B.bed_tcrossprodSelf(gf)
        t["grm_single_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        p1, i1, x1 = B.bed_cor(gf, size=size_kb, alpha=0.05, infos_pos=pos)
        t["cor_single_s"] = time.perf_counter() - t0
        ld1 = B.bed_ld_scores(gf, size=size_kb, infos_pos=pos)
        res = {"rows_check": True, "world": world, "n": n, "m": m,
               "colstats_identical": all(np.array_equal(np.asarray(st[k]), np.asarray(st1[k])) for k in st1),
               "col_counts_identical": bool(np.array_equal(cc, cc1)),
               "row_counts_identical": bool(np.array_equal(rc, rc1)),
               "grm_max_rel": float(np.max(np.abs(K - K1)) / np.max(np.abs(K1))),
               "cor_identical": bool(np.array_equal(p, p1) and np.array_equal(i, i1) and np.array_equal(x, x1, equal_nan=True)),
               "cor_nnz": int(p1[-1]), "ld_max_abs": float(np.max(np.abs(ld - ld1))), "halo": [int(b - lo), int(hi - e)]}
        res.update({k: round(v, 4) for k, v in t.items()})
        print(json.dumps(res), flush=True)
        ok = res["colstats_identical"] and res["col_counts_identical"] and res["row_counts_identical"] and \
            res["grm_max_rel"] < 1e-12 and res["cor_identical"] and res["ld_max_abs"] < 1e-9
        if not ok:
            print("DIST ROWS CHECK FAILED", flush=True)


if __name__ == "__main__":
    main()

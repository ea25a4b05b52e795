#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, 1 rank per GPU): column-sharded bed_prodVec / bed_cprodVec / bed_randomSVD
against the single-GPU result on rank 0.  Prints one JSON line on rank 0."""
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
               "svd_sharded_s": t_svd, "svd_single_s": t1, "nops": svd["nops"], "nops_single": svd1["nops"]}
        print(json.dumps(res), flush=True)
        ok = res["prodvec_max_rel"] < 1e-12 and res["cprodvec_max_rel"] < 1e-12 and res["svd_d_max_rel"] < 1e-7 \
            and res["svd_u_min_abs_corr"] > 1 - 1e-6
        if not ok:
            print("DIST CHECK FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

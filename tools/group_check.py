#!/usr/bin/env python
"""Single-process multi-GPU check of the C-ABI group entry points (bsg_group_*): one host process drives every visible GPU
(the shape an R session has).  Results of the sharded calls against ONE GPU holding the whole matrix, then timings of
bed_prodVec / bed_randomSVD / bed_tcrossprodSelf at a configs[1]-sized shape.  One JSON line per part.

    python tools/group_check.py [--ndev N] [--big]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bigsnpr_b200 as B  # noqa: E402
from bigsnpr_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ndev", type=int, default=0)
    ap.add_argument("--big", action="store_true")
    args = ap.parse_args()
    L = _lib.lib()
    ndev = args.ndev or L.bsg_device_count()
    devs = list(range(ndev))
    rng = np.random.default_rng(5)
    ok = True

    # ---- correctness: sharded vs one GPU, missing values, multisets
    n, m, seed = 20011, 64007, 77
    grp = B.Group.synthetic(n, m, devs, seed=seed, na_rate=0.01)
    one = B.Bed.synthetic(n, m, seed=seed, na_rate=0.01, device=0)
    sc = B.bed_scaleBinom(one)
    scg = grp.scaleBinom()
    x, y = rng.normal(size=m), rng.normal(size=n)
    a, a1 = grp.prodVec(x, center=sc["center"], scale=sc["scale"]), B.bed_prodVec(one, x, center=sc["center"], scale=sc["scale"])
    b, b1 = grp.cprodVec(y, center=sc["center"], scale=sc["scale"]), B.bed_cprodVec(one, y, center=sc["center"], scale=sc["scale"])
    ir = rng.integers(1, n + 1, size=5000).astype(np.int32)
    ic = rng.integers(1, m + 1, size=7000).astype(np.int32)
    xs, ys = rng.normal(size=ic.size), rng.normal(size=ir.size)
    c, c1 = grp.prodVec(xs, ind_row=ir, ind_col=ic), B.bed_prodVec(one, xs, ind_row=ir, ind_col=ic)
    d, d1 = grp.cprodVec(ys, ind_row=ir, ind_col=ic), B.bed_cprodVec(one, ys, ind_row=ir, ind_col=ic)
    a_again = grp.prodVec(x, center=sc["center"], scale=sc["scale"])
    rel = lambda u, v: float(np.max(np.abs(u - v)) / np.max(np.abs(v)))  # noqa: E731
    res = {"part": "products", "ndev": ndev, "scaling_identical": bool(np.array_equal(scg["center"], sc["center"]) and np.array_equal(scg["scale"], sc["scale"])),
           "prodvec": rel(a, a1), "cprodvec": rel(b, b1), "prodvec_multiset": rel(c, c1), "cprodvec_multiset": rel(d, d1),
           "repeat_bit_equal": bool(np.array_equal(a, a_again))}
    print(json.dumps(res), flush=True)
    ok &= res["scaling_identical"] and max(res["prodvec"], res["cprodvec"], res["prodvec_multiset"], res["cprodvec_multiset"]) < 1e-12 and res["repeat_bit_equal"]

    t0 = time.perf_counter()
    sg = grp.randomSVD(k=10)
    tg = time.perf_counter() - t0
    t0 = time.perf_counter()
    s1 = B.bed_randomSVD(one, k=10)
    t1 = time.perf_counter() - t0
    sub = np.sort(rng.choice(m, 30000, replace=False)).astype(np.int32) + 1
    sgs, s1s = grp.randomSVD(ind_col=sub, k=5), B.bed_randomSVD(one, ind_col=sub, k=5)
    res = {"part": "svd", "d_max_rel": float(np.max(np.abs(sg["d"] - s1["d"]) / s1["d"])),
           "u_min_abs_corr": float(np.min(np.abs(np.sum(sg["u"] * s1["u"], axis=0)))),
           "v_min_abs_corr": float(np.min(np.abs(np.sum(sg["v"] * s1["v"], axis=0)))),
           "subset_d_max_rel": float(np.max(np.abs(sgs["d"] - s1s["d"]) / s1s["d"])),
           "subset_v_min_abs_corr": float(np.min(np.abs(np.sum(sgs["v"] * s1s["v"], axis=0)))),
           "nops": sg["nops"], "nops_one": s1["nops"], "group_s": tg, "one_s": t1}
    print(json.dumps(res), flush=True)
    ok &= res["d_max_rel"] < 1e-7 and res["u_min_abs_corr"] > 1 - 1e-6 and res["v_min_abs_corr"] > 1 - 1e-6 and res["subset_d_max_rel"] < 1e-7
    grp.close()
    one.close()

    n2, m2 = 6000, 30001
    grp = B.Group.synthetic(n2, m2, devs, seed=91, na_rate=0.01)
    one = B.Bed.synthetic(n2, m2, seed=91, na_rate=0.01, device=0)
    sc = B.bed_scaleBinom(one)
    t0 = time.perf_counter()
    Kg = grp.tcrossprodSelf(sc["center"], sc["scale"])
    tg = time.perf_counter() - t0
    t0 = time.perf_counter()
    K1, _, _ = B.bed_tcrossprodSelf(one)
    t1 = time.perf_counter() - t0
    res = {"part": "grm", "max_rel": rel(Kg, K1), "symmetric": bool(np.array_equal(Kg, Kg.T)), "group_s": tg, "one_s": t1}
    print(json.dumps(res), flush=True)
    ok &= res["max_rel"] < 1e-12
    grp.close()
    one.close()

    if args.big:
        # ---- timings at a configs[1]-sized shape (weak: 500,000 columns per device) and the configs[3] GRM
        n, mper = 50_000, 500_000
        grp = B.Group.synthetic(n, mper * ndev, devs, seed=20250925)
        sc = grp.scaleBinom()
        x = rng.normal(size=mper * ndev)
        for _ in range(3):
            grp.prodVec(x, center=sc["center"], scale=sc["scale"])
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            grp.prodVec(x, center=sc["center"], scale=sc["scale"])
        tp = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        sv = grp.randomSVD(k=10)
        ts = time.perf_counter() - t0
        print(json.dumps({"part": "timing_cfg2_weak", "ndev": ndev, "n": n, "m": mper * ndev,
                          "group_prodvec_host_call_ms": tp * 1e3, "genotypes_per_s_host_call": n * mper * ndev / tp,
                          "randomsvd_k10_s": ts, "nops": sv["nops"]}), flush=True)
        grp.close()
        n, m = 10_000, 1_000_000
        grp = B.Group.synthetic(n, m, devs, seed=20250927)
        sc = grp.scaleBinom()
        grp.tcrossprodSelf(sc["center"], sc["scale"])
        t0 = time.perf_counter()
        grp.tcrossprodSelf(sc["center"], sc["scale"])
        tk = time.perf_counter() - t0
        print(json.dumps({"part": "timing_cfg4_grm", "ndev": ndev, "n": n, "m": m, "tcrossprod_s": tk}), flush=True)
        grp.close()
    print("GROUP CHECK " + ("OK" if ok else "FAILED"), flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

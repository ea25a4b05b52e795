#!/usr/bin/env python
"""Timings of the post-SVD rows on BASELINE.json configs[1]'s shape: PCA projection
(prod_and_rowSumsSq, K = 10) and pcadapt's multLinReg (K = 10), with a CPU-oracle sample for scale.
Prints one JSON line per op.  Not the driver's bench."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bigsnpr_b200 as B  # noqa: E402


def timeit(f, reps=3):
    f()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = f()
    return (time.perf_counter() - t0) / reps, r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=50000)
    ap.add_argument("--m", type=int, default=500000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--cpu-cols", type=int, default=4000)
    a = ap.parse_args()
    rng = np.random.default_rng(1)
    for na_rate in (0.0, 0.01):
        g = B.Bed.synthetic(a.n, a.m, seed=20250925, na_rate=na_rate)
        sc = B.bed_scaleBinom(g)
        ir, ic = np.arange(1, a.n + 1, dtype=np.int32), np.arange(1, a.m + 1, dtype=np.int32)
        V = np.asfortranarray(rng.normal(size=(a.m, a.k)) / np.sqrt(a.m))
        U = np.asfortranarray(np.linalg.qr(rng.normal(size=(a.n, a.k)))[0])
        L = B._lib.lib()
        n0 = L.bsg_launch_count()
        t, (XV, rss) = timeit(lambda: B.prod_and_rowSumsSq(g, ir, ic, sc["center"], sc["scale"], V))
        launches = (L.bsg_launch_count() - n0) // 4
        passes = a.k + 1 + (1 if na_rate > 0 else 0)
        print(json.dumps({"op": "prod_and_rowSumsSq", "n": a.n, "m": a.m, "K": a.k, "na_rate": na_rate, "seconds": t,
                          "matrix_passes": passes, "packed_GBps": passes * a.m * ((a.n + 3) // 4) / t / 1e9,
                          "genotypes_per_s": float(a.n) * a.m / t, "launches": launches}), flush=True)
        t2, ts = timeit(lambda: B.multLinReg(g, ir, ic, U))
        passes = a.k * (2 if na_rate > 0 else 1)
        print(json.dumps({"op": "multLinReg", "n": a.n, "m": a.m, "K": a.k, "na_rate": na_rate, "seconds": t2,
                          "matrix_passes": passes, "packed_GBps": passes * a.m * ((a.n + 3) // 4) / t2 / 1e9,
                          "genotypes_per_s": float(a.n) * a.m / t2}), flush=True)
        if na_rate == 0.0:
            from oracle import ref

            o = ref.synth_bed(a.n, a.cpu_cols, seed=20250925)
            icc = np.arange(1, a.cpu_cols + 1, dtype=np.int32)
            t0 = time.perf_counter()
            XVo, rsso = ref.prod_and_rowSumsSq(o, ir, icc, sc["center"][:a.cpu_cols], sc["scale"][:a.cpu_cols],
                                               V[:a.cpu_cols])
            tc = time.perf_counter() - t0
            t0 = time.perf_counter()
            tso = ref.multLinReg(o, ir, icc, U, ncores=ref.max_threads())
            tm = time.perf_counter() - t0
            XVs, rsss = B.prod_and_rowSumsSq(g, ir, icc, sc["center"][:a.cpu_cols], sc["scale"][:a.cpu_cols], V[:a.cpu_cols])
            print(json.dumps({"op": "cpu_oracle_sample", "cols": a.cpu_cols, "prod_and_rowSumsSq_s": tc,
                              "prod_and_rowSumsSq_threads": 1, "prod_full_extrapolated_s": tc * a.m / a.cpu_cols,
                              "multLinReg_s": tm, "multLinReg_threads": ref.max_threads(),
                              "multLinReg_full_extrapolated_s": tm * a.m / a.cpu_cols,
                              "xv_max_rel_err": float(np.max(np.abs(XVs - XVo)) / np.max(np.abs(XVo))),
                              "rss_max_rel_err": float(np.max(np.abs(rsss - rsso) / rsso)),
                              "tscore_max_abs_err": float(np.nanmax(np.abs(ts[:a.cpu_cols] - tso)))}), flush=True)
        g.close()


if __name__ == "__main__":
    main()

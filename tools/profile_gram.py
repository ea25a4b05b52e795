#!/usr/bin/env python
"""Driver for ncu captures / timings of the dense-contraction rows: bed_ld_scores, bed_cor, bed_tcrossprodSelf on synthetic
shapes.  Prints wall seconds per call.

    ncu --set full --clock-control none --import-source on -k regex:k_gramt -c 1 -o gpurun_out/prof \
        python tools/profile_gram.py --op grm --n 10000 --m 100000 --reps 1
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bigsnpr_b200 as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--op", choices=("ld", "cor", "grm", "clump"), default="ld")
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--m", type=int, default=20000)
    ap.add_argument("--size", type=int, default=500)
    ap.add_argument("--na-rate", type=float, default=0.0)
    ap.add_argument("--ld-rho", type=float, default=0.9)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    g = B.Bed.synthetic(a.n, a.m, seed=20250927, na_rate=a.na_rate, ld_rho=a.ld_rho if a.op != "grm" else 0.0)
    sc = B.bed_scaleBinom(g)
    times = []
    for _ in range(a.reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if a.op == "ld":
            B.bed_ld_scores(g, size=a.size)
        elif a.op == "cor":
            B.bed_cor(g, size=a.size)
        elif a.op == "clump":
            import numpy as np

            B.bed_clumping(g, infos_chr=np.ones(a.m, dtype=int), infos_pos=1000.0 * np.arange(1, a.m + 1))
        else:
            B.bed_tcrossprodSelf(g, fun_scaling=lambda *x, **k: sc)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    print(json.dumps({"op": a.op, "n": a.n, "m": a.m, "na_rate": a.na_rate, "seconds": times}))


if __name__ == "__main__":
    main()

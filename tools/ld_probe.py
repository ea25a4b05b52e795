#!/usr/bin/env python
"""Repeated bed_ld_scores / bed_cor calls on configs[2] with per-call wall times (host overheads, pool behaviour)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bigsnpr_b200 as B  # noqa: E402


def t(f):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = f()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, r


g = B.Bed.synthetic(100000, 200000, seed=20250927, layouts=B.LAYOUT_SNP_MAJOR, ld_rho=0.9)
for k in range(4):
    print("ld", k, "%.4f" % t(lambda: B.bed_ld_scores(g, size=500))[0], flush=True)
for k in range(2):
    dt, r = t(lambda: B.bed_cor(g, size=500))
    print("cor", k, "%.4f" % dt, flush=True)
    del r
for k in range(3):
    print("ld after cor", k, "%.4f" % t(lambda: B.bed_ld_scores(g, size=500))[0], flush=True)
dt, r = t(lambda: B.bed_cor(g, size=500))
print("cor keep", "%.4f" % dt, flush=True)
for k in range(2):
    print("ld with cor result alive", k, "%.4f" % t(lambda: B.bed_ld_scores(g, size=500))[0], flush=True)

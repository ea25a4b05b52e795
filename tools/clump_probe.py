#!/usr/bin/env python
"""Repeated bed_clumping calls on configs[2] after a bed_cor on another handle (allocator state), per-call wall times."""
import os
import sys
import time

import numpy as np

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


m = 200000
chrom, pos = np.ones(m, dtype=int), 1000.0 * np.arange(1, m + 1)
g = B.Bed.synthetic(100000, m, seed=20250927, layouts=B.LAYOUT_SNP_MAJOR, ld_rho=0.9)
dt, r = t(lambda: B.bed_cor(g, size=500))
print("cor on g %.4f" % dt, flush=True)
g.close()
gc = B.Bed.synthetic(100000, m, seed=20250929, ld_rho=0.9)
print("layouts of gc", gc.layouts, "free GB %.1f" % (torch.cuda.mem_get_info()[0] / 1e9), flush=True)
for k in range(4):
    dt, keep = t(lambda: B.bed_clumping(gc, infos_chr=chrom, infos_pos=pos))
    print("clump", k, "%.4f" % dt, keep.size, "free GB %.1f" % (torch.cuda.mem_get_info()[0] / 1e9), flush=True)
del r
for k in range(2):
    dt, keep = t(lambda: B.bed_clumping(gc, infos_chr=chrom, infos_pos=pos))
    print("clump after dropping the cor result", k, "%.4f" % dt, flush=True)

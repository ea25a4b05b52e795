#!/usr/bin/env python
"""Small driver for ncu captures of the matvec kernels: stages a synthetic matrix of the given shape and runs a few
bed_prodVec / bed_cprodVec on device-resident vectors.

    ncu --set full --clock-control none --import-source on -k regex:k_pmv -s 4 -c 2 -o gpurun_out/prof \
        python tools/profile_pmv.py --n 487000 --m 137500 --layout snp
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bigsnpr_b200 as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=50000)
    ap.add_argument("--m", type=int, default=500000)
    ap.add_argument("--na-rate", type=float, default=0.0)
    ap.add_argument("--layout", choices=("snp", "both"), default="snp")
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--side", choices=("x", "xt", "both"), default="both")
    a = ap.parse_args()
    lay = B.LAYOUT_SNP_MAJOR if a.layout == "snp" else (B.LAYOUT_SNP_MAJOR | B.LAYOUT_SAMPLE_MAJOR)
    g = B.Bed.synthetic(a.n, a.m, seed=20250928, na_rate=a.na_rate, layouts=lay)
    sc = B.bed_scaleBinom(g)
    v = B.View(g, center=sc["center"], scale=sc["scale"])
    dev = torch.device("cuda", 0)
    x = torch.randn(a.m, dtype=torch.float64, device=dev)
    y = torch.randn(a.n, dtype=torch.float64, device=dev)
    ox = torch.empty(a.n, dtype=torch.float64, device=dev)
    oy = torch.empty(a.m, dtype=torch.float64, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    import ctypes as C

    from bigsnpr_b200 import _lib

    L = _lib.lib()
    L.bsg_set_kernel_timing(1)
    for _ in range(a.reps):
        if a.side in ("x", "both"):
            v.prodvec_dev(x.data_ptr(), ox.data_ptr(), s)
        if a.side in ("xt", "both"):
            v.cprodvec_dev(y.data_ptr(), oy.data_ptr(), s)
    torch.cuda.synchronize()
    cnt, tot = C.c_int(0), C.c_double(0)
    L.bsg_kernel_time_stats(C.byref(cnt), C.byref(tot))
    alg = ((a.n + 3) // 4) * a.m
    ms = tot.value / max(cnt.value, 1)
    print("done", float(ox.sum()), float(oy.sum()), "avg kernel ms %.4f over %d launches -> %.1f GB/s" % (ms, cnt.value, alg / ms / 1e6))


if __name__ == "__main__":
    main()

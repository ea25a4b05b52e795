import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bigsnpr_b200 as B
g = B.Bed.synthetic(10000, 100000, seed=20250928, na_rate=float(os.environ.get("NA","0")))
sc = B.bed_scaleBinom(g)
fun = lambda *a, **k: sc
for _ in range(2):
    torch.cuda.synchronize(); t=time.perf_counter(); K,c,s = B.bed_tcrossprodSelf(g, fun_scaling=fun); torch.cuda.synchronize(); print("total", time.perf_counter()-t)

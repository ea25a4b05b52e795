#!/usr/bin/env python
"""Randomised parity sweep on the GPU: random shapes (ragged n, column counts around the kernels' step sizes), NA
rates, layouts, index multisets and scalings; every product / statistic against the CPU oracle.  Prints one JSON
summary line; exit code 1 on the first mismatch.  Not part of the driver's test run (use: python tools/stress.py)."""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bigsnpr_b200 as B  # noqa: E402
from oracle import ref  # noqa: E402


def close(a, b, scale, tol=1e-11):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    if a.shape != b.shape:
        return False, "shape %s vs %s" % (a.shape, b.shape)
    if a.size == 0:
        return True, 0.0
    err = float(np.max(np.abs(a - b) / np.maximum(scale, 1e-300)))
    return err < tol, err


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=12345)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    tmp = tempfile.mkdtemp()
    worst = {}
    for case in range(a.cases):
        n = int(rng.choice([1, 2, 3, 4, 5, 63, 64, 65, 127, 129, 255, 257, 511, 513, 1000, 2049, 3001]))
        m = int(rng.choice([1, 2, 31, 32, 33, 63, 65, 191, 193, 255, 257, 1023, 1025, 2500, 4097]))
        na = float(rng.choice([0.0, 0.0, 0.003, 0.05, 0.4]))
        G = rng.integers(0, 3, size=(n, m)).astype(np.uint8)
        if na > 0:
            G[rng.uniform(size=G.shape) < na] = 3
        path = ref.write_bed(os.path.join(tmp, "s%d.bed" % case), G)
        o = ref.OracleBed(path)
        layouts = int(rng.choice([B.LAYOUT_SNP_MAJOR, B.LAYOUT_SNP_MAJOR | B.LAYOUT_SAMPLE_MAJOR]))
        g = B.Bed(path, layouts=layouts)
        mode = int(rng.integers(0, 3))
        if mode == 0:
            ir, ic = np.arange(1, n + 1), np.arange(1, m + 1)
        elif mode == 1:
            ir = np.sort(rng.choice(n, max(1, n // 2), replace=False)) + 1
            ic = rng.permutation(m)[: max(1, (2 * m) // 3)] + 1
        else:
            ir, ic = rng.integers(1, n + 1, size=n + 3), rng.integers(1, m + 1, size=m + 5)
        ir, ic = ir.astype(np.int32), ic.astype(np.int32)
        y_col, y_row = rng.normal(size=ic.size) * 10.0 ** rng.integers(-3, 4), rng.normal(size=ir.size)
        scaled = bool(rng.integers(0, 2))
        c = rng.normal(size=ic.size) if scaled else None
        s = rng.uniform(0.3, 2.0, size=ic.size) if scaled else None
        tag = dict(case=case, n=n, m=m, na=na, layouts=layouts, mode=mode, scaled=scaled)
        sc1 = np.max(np.abs(y_col / (s if scaled else 1.0))) * ic.size * 3
        sc2 = np.max(np.abs(y_row)) * ir.size * 3 / (np.min(s) if scaled else 1.0)
        checks = {
            "prodvec": close(B.bed_prodVec(g, y_col, ir, ic, c, s), ref.bed_prodVec(o, y_col, ir, ic, c, s), sc1),
            "cprodvec": close(B.bed_cprodVec(g, y_row, ir, ic, c, s), ref.bed_cprodVec(o, y_row, ir, ic, c, s), sc2),
        }
        cc, cr = B.bed_counts(g, ir, ic), B.bed_counts(g, ir, ic, byrow=True)
        checks["col_counts"] = (bool(np.array_equal(cc, ref.bed_counts(o, ir, ic))), 0.0)
        checks["row_counts"] = (bool(np.array_equal(cr, ref.bed_counts(o, ir, ic, byrow=True))), 0.0)
        checks["readbina2"] = (bool(np.array_equal(B.readbina2(g, ir, ic), ref.read_bed(o, ir, ic, na_val=3))), 0.0)
        cs = (c if scaled else np.zeros(ic.size)), (s if scaled else np.ones(ic.size))
        V = rng.normal(size=(ic.size, 2))
        XV, rss = B.prod_and_rowSumsSq(g, ir, ic, cs[0], cs[1], V)
        XVo, rsso = ref.prod_and_rowSumsSq(o, ir, ic, cs[0], cs[1], V)
        vscale = np.max(np.abs(V)) * ic.size * 3 / np.min(cs[1]) * (1.0 + np.max(np.abs(cs[0])))
        checks["XV"] = close(XV, XVo, max(np.max(np.abs(XVo)), 1e-3 * vscale))
        checks["rowSumsSq"] = close(rss, rsso, np.max(np.abs(rsso)) + 1e-300, tol=1e-12)
        if ir.size >= 3:
            U = np.linalg.qr(rng.normal(size=(ir.size, 2)))[0]
            t, to = B.multLinReg(g, ir, ic, U), ref.multLinReg(o, ir, ic, U)
            okm = ~np.isnan(to) & ~np.isnan(t) & (np.abs(to) < 1e6)
            checks["multLinReg"] = close(t[okm], to[okm], 1.0 + np.abs(to[okm]), tol=1e-7)
        if case % 3 == 0 and m >= 3:
            icc = np.sort(np.unique(ic)).astype(np.int32)  # window ops want sorted positions
            irr = np.unique(ir).astype(np.int32)
            pos = np.cumsum(rng.integers(1, 3000, size=icc.size)).astype(np.float64)
            kw = dict(size=float(rng.choice([1.0, 20.0, 500.0])), infos_pos=pos)
            if irr.size > 3:
                import warnings

                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    p1, i1, x1 = B.bed_cor(g, irr, icc, alpha=0.3, **kw)
                po, io, xo = ref.cor0(o, irr, icc, alpha=0.3, **kw)
                checks["cor"] = (bool(np.array_equal(p1, po) and np.array_equal(i1, io)
                                      and np.array_equal(x1, xo, equal_nan=True)), 0.0)
                ld, ldo = B.bed_ld_scores(g, irr, icc, **kw), ref.ld0(o, irr, icc, **kw)
                okl = ~np.isnan(ldo)
                checks["ld"] = (bool(np.array_equal(np.isnan(ld), np.isnan(ldo))) and
                                close(ld[okl], ldo[okl], np.abs(ldo[okl]) + 1e-300, tol=1e-10)[0], 0.0)
            if n >= 8 and n <= 600:
                st = ref.bed_colstats(o, irr, icc)
                good = (st["denoX"] > 0)
                if good.sum() >= 2:
                    icg = icc[good]
                    cen = st["sumX"][good] / st["nb_nona_col"][good]
                    sca = np.sqrt(st["denoX"][good])
                    K = np.empty((irr.size, irr.size))
                    B._lib.check(B._lib.lib().bsg_tcrossprod(g._h, irr.ctypes.data_as(B._lib.c_int_p), irr.size,
                                                           icg.ctypes.data_as(B._lib.c_int_p), icg.size,
                                                           cen.ctypes.data_as(B._lib.c_dbl_p), sca.ctypes.data_as(B._lib.c_dbl_p),
                                                           K.ctypes.data_as(B._lib.c_dbl_p)))
                    X = ref.read_bed_scaled(o, irr, icg, cen, sca)
                    Ko = X @ X.T
                    checks["grm"] = close(K, Ko, np.max(np.abs(Ko)) + 1e-300, tol=1e-10)
        for k, (ok, err) in checks.items():
            if not ok:
                print(json.dumps({"FAILED": k, "err": err, **tag}), flush=True)
                return 1
            if isinstance(err, float):
                worst[k] = max(worst.get(k, 0.0), err)
        g.close()
    print(json.dumps({"cases": a.cases, "seed": a.seed, "worst_relative_error": worst, "status": "all equal"}), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""Pins the CPU oracle (oracle/) to the reference's own fixtures and relational tests.

Golden inputs (copied from /root/reference): inst/extdata/example.bed (517 x 4542, no NA),
inst/extdata/example-missing.bed (200 x 500, 2788 NA), tests/testthat/testdata/example.ld (PLINK --r2).
"""
import os

import numpy as np
import pytest


def test_decode_known_counts(oracle, obed, obed_na):
    # example.bed counts 0/1/2 = 1,212,929 / 891,930 / 243,355, no NA;
    # example-missing.bed has 2,788 NAs.
    D = oracle.decode_dense(obed)
    assert D.shape == (517, 4542)
    assert np.bincount(D.ravel(), minlength=4).tolist() == [1212929, 891930, 243355, 0]
    Dn = oracle.decode_dense(obed_na)
    assert Dn.shape == (200, 500)
    assert int((Dn == 3).sum()) == 2788


def test_get_code_matches_bit_definition(oracle):
    # R/utils.R:21-31 builds the same table from bits: geno = !b1 + !b2, NA if (!b1 == 0 and !b2 == 1)
    code = oracle.getCode()
    for b in range(256):
        for s in range(4):
            c = (b >> (2 * s)) & 3
            assert code[s, b] == {0: 2, 1: 3, 2: 1, 3: 0}[c]
    inv = oracle.getInverseCode()
    G = np.array([[0], [1], [2], [3]])
    assert oracle.write_bed_bytes(G)[0, 0] == inv[0, 1, 2, 3]


def test_file_validation(oracle, tmp_path, golden_dir):
    # src/bed-acc-xptr.cpp:21-34
    good = os.path.join(golden_dir, "example.bed")
    with pytest.raises(oracle.OracleError, match="n or p does not match"):
        oracle.OracleBed(good, 517, 4541)
    raw = bytearray(open(good, "rb").read())
    bad = tmp_path / "bad.bed"
    raw2 = bytearray(raw); raw2[0] = 0
    bad.write_bytes(raw2)
    with pytest.raises(oracle.OracleError, match="not a binary PED"):
        oracle.OracleBed(str(bad), 517, 4542)
    raw3 = bytearray(raw); raw3[2] = 0
    bad.write_bytes(raw3)
    with pytest.raises(oracle.OracleError, match="Variant-major"):
        oracle.OracleBed(str(bad), 517, 4542)


def test_cor_matches_plink_golden(oracle, obed, golden_dir):
    # tests/testthat/test-2-corr.R:14-58 : r^2 from corMat(thr = sqrt(0.2)) == PLINK's example.ld,
    # same sparsity pattern, values to 1e-6 (PLINK prints 6 significant digits).
    rows = [l.split() for l in open(os.path.join(golden_dir, "example.ld"))][1:]
    a = np.array([int(r[2][3:]) for r in rows])
    b = np.array([int(r[5][3:]) for r in rows])
    r2 = np.array([float(r[6]) for r in rows])
    assert len(rows) == 1431 and np.all(a < b)
    thr = np.full(obed.nrow, np.sqrt(0.2))
    pos = np.arange(1, obed.ncol + 1, dtype=float)
    for size in (7, 200, 5000):
        p, i, x = oracle.corMat(obed, obed.rows_along(), obed.cols_along(), size, thr, pos,
                                fill_diag=False, ncores=oracle.max_threads())
        j = np.repeat(np.arange(obed.ncol), np.diff(p))
        keep = (b - a) <= size
        got = {(ii, jj): v * v for ii, jj, v in zip(i.tolist(), j.tolist(), x.tolist())}
        want = {(ii, jj): v for ii, jj, v in zip(a[keep].tolist(), b[keep].tolist(), r2[keep].tolist())}
        assert set(got) == set(want)
        err = max(abs(got[k] - want[k]) for k in want)
        assert err < 1e-6
        # rows ascending within each column (rev() in src/corr.cpp:90-92)
        for c in range(obed.ncol):
            seg = i[p[c]:p[c + 1]]
            assert np.all(np.diff(seg) > 0)


def test_prodvec_equals_dense(oracle, obed_na, rng):
    # tests/testthat/test-5-bed-prod-vec.R:18-41
    N, M = obed_na.nrow, obed_na.ncol
    for rep in range(10):
        n, m = rng.integers(1, N + 1), rng.integers(1, M + 1)
        ind_row = rng.choice(N, n, replace=False) + 1
        ind_col = rng.choice(M, m, replace=False) + 1
        for center, scale in ((np.zeros(m), np.ones(m)), (rng.normal(size=m), rng.uniform(size=m))):
            X = oracle.read_bed_scaled(obed_na, ind_row, ind_col, center, scale)
            y_col, y_row = rng.normal(size=m), rng.normal(size=n)
            for nc in (1, 3):
                np.testing.assert_allclose(
                    oracle.bed_prodVec(obed_na, y_col, ind_row, ind_col, center, scale, ncores=nc),
                    X @ y_col, rtol=1e-10, atol=1e-10)
                np.testing.assert_allclose(
                    oracle.bed_cprodVec(obed_na, y_row, ind_row, ind_col, center, scale, ncores=nc),
                    X.T @ y_row, rtol=1e-10, atol=1e-10)
    # dimension errors (tests/testthat/test-5-bed-prod-vec.R:43-50)
    with pytest.raises(oracle.OracleError, match="Incompatibility between dimensions"):
        oracle.bed_prodVec(obed_na, rng.normal(size=21), np.arange(1, 22), np.arange(1, 12))
    with pytest.raises(oracle.OracleError, match="Incompatibility between dimensions"):
        oracle.bed_cprodVec(obed_na, rng.normal(size=11), np.arange(1, 22), np.arange(1, 12))


def test_counts_colstats_scaling(oracle, obed_na, rng):
    # tests/testthat/test-2-bed-clumping-SVD.R:95-136 (relational): counts vs dense decode,
    # scaleBinom center == 2 * af.
    D = oracle.decode_dense(obed_na)
    ind_row = rng.choice(obed_na.nrow, 120, replace=False) + 1
    ind_col = rng.choice(obed_na.ncol, 300, replace=False) + 1
    sub = D[np.ix_(ind_row - 1, ind_col - 1)]
    want_col = np.stack([(sub == k).sum(0) for k in range(4)])
    want_row = np.stack([(sub == k).sum(1) for k in range(4)])
    for nc in (1, 2):
        assert np.array_equal(oracle.bed_counts(obed_na, ind_row, ind_col, ncores=nc), want_col)
        assert np.array_equal(oracle.bed_counts(obed_na, ind_row, ind_col, byrow=True, ncores=nc), want_row)
    st = oracle.bed_colstats(obed_na, ind_row, ind_col)
    subf = np.where(sub == 3, np.nan, sub.astype(float))
    np.testing.assert_array_equal(st["sumX"], np.nansum(subf, 0))
    np.testing.assert_array_equal(st["nb_nona_col"], (sub != 3).sum(0))
    sc = oracle.bed_scaleBinom(obed_na, ind_row, ind_col)
    maf = oracle.bed_MAF(obed_na, ind_row, ind_col, ncores=2)
    np.testing.assert_array_equal(sc["center"], 2 * maf["af"])
    # read_bed with replacement indices (tests/testthat/test-1-readBed.R:71-87)
    ir = rng.integers(1, obed_na.nrow + 1, 50)
    ic = rng.integers(1, obed_na.ncol + 1, 60)
    got = oracle.read_bed(obed_na, ir, ic, na_val=3)
    assert np.array_equal(got, D[np.ix_(ir - 1, ic - 1)])


def test_cor_pairwise_complete_and_ld(oracle, rng, tmp_path):
    # tests/testthat/test-2-corr.R:62-116,148-171 and test-2-ld-scores.R:15-30,54-64
    N, M = 500, 100
    G = rng.integers(0, 4, size=(N, M))
    fbm = oracle.OracleFBM(G.astype(np.uint8))
    bed = oracle.OracleBed(oracle.write_bed(str(tmp_path / "fake.bed"), G))
    ind_row = rng.choice(N, N // 2, replace=False) + 1
    ind_col = np.sort(rng.choice(M, M // 2, replace=False)) + 1
    size = 30
    p, i, x = oracle.cor0(fbm, ind_row, ind_col, size=size)
    p2, i2, x2 = oracle.cor0(bed, ind_row, ind_col, size=size, ncores=2)
    assert np.array_equal(p, p2) and np.array_equal(i, i2) and np.array_equal(x, x2)
    sub = G[np.ix_(ind_row - 1, ind_col - 1)].astype(float)
    sub[sub == 3] = np.nan
    m = ind_col.size
    j = np.repeat(np.arange(m), np.diff(p))
    for ii, jj, v in zip(i, j, x):
        if ii == jj:
            assert v == 1.0
            continue
        ok = ~np.isnan(sub[:, ii]) & ~np.isnan(sub[:, jj])
        r = np.corrcoef(sub[ok, ii], sub[ok, jj])[0, 1]
        assert abs(r - v) < 1e-12
    # ld == colSums(corr^2) of the symmetric matrix
    sym = np.zeros((m, m)); sym[i, j] = x; sym = sym + sym.T - np.diag(np.diag(sym))
    ld = oracle.ld0(fbm, ind_row, ind_col, size=size)
    np.testing.assert_allclose(ld, (sym ** 2).sum(0), rtol=1e-12)
    np.testing.assert_allclose(oracle.ld0(bed, ind_row, ind_col, size=size, ncores=2), ld, rtol=1e-12)
    assert np.all(oracle.ld0(fbm, size=0.5) == 1.0)
    # alpha threshold keeps exactly the pairs significant at level alpha
    alpha = 0.07
    p3, i3, x3 = oracle.cor0(fbm, ind_row, ind_col, size=size, alpha=alpha, fill_diag=False)
    from scipy import stats
    j3 = np.repeat(np.arange(m), np.diff(p3))
    kept = set(zip(i3.tolist(), j3.tolist()))
    for jj in range(m):
        for ii in range(max(0, jj - size), jj):
            ok = ~np.isnan(sub[:, ii]) & ~np.isnan(sub[:, jj])
            k = ok.sum()
            r = np.corrcoef(sub[ok, ii], sub[ok, jj])[0, 1]
            t = r * np.sqrt((k - 2) / (1 - r * r))
            pval = 2 * stats.t.sf(abs(t), k - 2)
            assert ((ii, jj) in kept) == (pval < alpha)


def test_cor_nan_on_zero_variance(oracle):
    # tests/testthat/test-2-corr.R:163-171
    rng = np.random.default_rng(3)
    G = rng.integers(0, 3, size=(10, 10)); G[:, 0] = 0
    fbm = oracle.OracleFBM(G.astype(np.uint8))
    p, i, x = oracle.cor0(fbm)
    j = np.repeat(np.arange(10), np.diff(p))
    assert np.isnan(x[(i == 0) & (j > 0)]).all() and p[-1] == 55


def test_grm_matches_svd(oracle, obed_na):
    # tests/testthat/test-2-bed-clumping-SVD.R:72-79: sqrt(eigen(K)[1:10]) == svd$d
    ind_col = np.arange(1, obed_na.ncol + 1, 3, dtype=np.int32)
    K, c, s = oracle.bed_tcrossprodSelf(obed_na, ind_col=ind_col, block_size=37)
    ev = np.linalg.eigvalsh(K)[::-1][:10]
    svd = oracle.bed_randomSVD(obed_na, ind_col=ind_col, k=10)
    np.testing.assert_allclose(np.sqrt(ev), svd["d"], rtol=1e-10)


def test_clumping_relational(oracle, obed, obed_na):
    # tests/testthat/test-2-bed-clumping-SVD.R:34-39,47-49: rescaling positions and window together changes nothing,
    # excluded variants never come back, pruning is monotone in the threshold; kept variants are pairwise below thr.
    chrom, pos = oracle.read_bim(obed.bedfile)
    k = oracle.bed_clumping(obed)
    assert len(k) == 4270 and k.min() >= 1 and k.max() <= obed.ncol and np.all(np.diff(k) > 0)
    assert np.array_equal(oracle.bed_clumping(obed, infos_chr=chrom, infos_pos=pos * 1e6, size=500 * 1e6), k)
    assert np.array_equal(oracle.bed_clumping(obed, infos_chr=chrom, infos_pos=pos / 1e6, size=500 / 1e6), k)
    assert oracle.bed_clumping(obed, exclude=np.arange(1, 101)).min() > 100
    assert len(oracle.bed_clumping(obed, thr_r2=0.05)) < len(k) < len(oracle.bed_clumping(obed, thr_r2=0.8))
    # no pair of kept variants inside the window exceeds the threshold (r2 of the scaled dot product)
    kn = oracle.bed_clumping(obed_na, thr_r2=0.3)
    st = oracle.bed_colstats(obed_na, obed_na.rows_along(), kn)
    X = oracle.read_bed_scaled(obed_na, obed_na.rows_along(), kn, st["sumX"] / st["nb_nona_col"], np.sqrt(st["denoX"]))
    R2 = (X.T @ X) ** 2
    chrom_n, pos_n = oracle.read_bim(obed_na.bedfile)
    same = chrom_n[kn - 1][:, None] == chrom_n[kn - 1][None, :]
    near = np.abs(pos_n[kn - 1][:, None] - pos_n[kn - 1][None, :]) <= (100 / 0.3) * 1000
    off = ~np.eye(len(kn), dtype=bool)
    assert np.all(R2[same & near & off] <= 0.3)


def test_snp_clumping_fbm_twin(oracle, obed):
    # tests/testthat/test-2-bed-clumping-SVD.R:34-36,47-48: on a file without missing values snp_clumping (FBM.code256,
    # src/clumping.cpp) and bed_clumping (src/clumping-bed.cpp) keep the same variants; kept pairs are below thr
    chrom, pos = oracle.read_bim(obed.bedfile)
    G = oracle.read_bed(obed, obed.rows_along(), obed.cols_along(), na_val=3).astype(np.uint8)
    fbm = oracle.OracleFBM(G)
    k_bed = oracle.bed_clumping(obed)
    k_fbm = oracle.snp_clumping(fbm, chrom, infos_pos=pos)
    assert np.array_equal(k_fbm, k_bed)
    ir = np.arange(1, obed.nrow + 1, 2).astype(np.int32)
    k2 = oracle.snp_clumping(fbm, chrom, ind_row=ir, thr_r2=0.1, size=50)  # index-based window
    assert 0 < len(k2) < obed.ncol
    Gs = G[ir - 1][:, k2 - 1].astype(float)
    with np.errstate(all="ignore"):
        R2 = np.corrcoef(Gs.T) ** 2
    idx = np.arange(len(k2))
    near = (np.abs((k2[:, None] - k2[None, :])) <= 50) & (chrom[k2 - 1][:, None] == chrom[k2 - 1][None, :])
    off = idx[:, None] != idx[None, :]
    assert np.nanmax(R2[near & off]) <= 0.1 + 1e-12


def test_snp_clumping_against_reference_rds_golden(oracle, obed, golden_dir):
    # tests/testthat/test-6-PRS.R:25-31: snp_clumping(G, S = abs(gwas$score), size = 250, infos.pos) against the kept
    # indices stored in testdata/clumping.rds, `expect_gt(mean(ind.keep %in% ind.keep2), 0.98)`.  The priority vector is
    # recovered from testdata/pval.rds (the same test pins predict(gwas, log10 = FALSE) to it, :19-22): the p-value is a
    # decreasing function of |score| and clumping only uses the ORDER of S, so S = -pval ranks the variants like
    # abs(gwas$score).  Both vectors were parsed from the reference's RDS files by tests/golden/make_rds_golden.py.
    gold = np.load(os.path.join(golden_dir, "prs_clumping.npz"))
    pval, keep2 = gold["pval"], gold["keep"]
    assert pval.size == obed.ncol == 4542 and keep2.size == 4390 and np.unique(keep2).size == 4390
    chrom, pos = oracle.read_bim(obed.bedfile)
    G = oracle.read_bed(obed, obed.rows_along(), obed.cols_along(), na_val=3).astype(np.uint8)
    keep = oracle.snp_clumping(oracle.OracleFBM(G), chrom, S=-pval, size=250, infos_pos=pos)
    frac = np.mean(np.isin(keep, keep2))
    assert frac > 0.98, frac                      # the reference's own bar
    assert abs(keep.size - keep2.size) <= 0.02 * keep2.size
    # bed twin on the same file (no missing values): identical indices (test-2-bed-clumping-SVD.R:47-48)
    assert np.array_equal(oracle.bed_clumping(obed, S=-pval, size=250), keep)


def test_projection_and_pcadapt_oracle_against_numpy(oracle, obed_na):
    # src/bed-fun.cpp:103-133 and src/multLinReg.cpp:8-60 restated in the oracle, pinned on independent NumPy / SciPy
    # formulations: X~ V and row sums of squares of the dense scaled matrix; t-score = slope / stderr of the simple
    # regression of the genotype on u over the samples where the genotype is present (tests/testthat/test-4-pcadapt.R
    # compares with the pcadapt package, which computes the same statistic).
    from scipy import stats

    rng = np.random.default_rng(4)
    n, m = obed_na.nrow, obed_na.ncol
    ir = np.sort(rng.choice(n, 150, replace=False)) + 1
    ic = rng.choice(m, 120, replace=False) + 1
    sc = oracle.bed_scaleBinom(obed_na, ir, ic)
    V = rng.normal(size=(ic.size, 3))
    XV, rss = oracle.prod_and_rowSumsSq(obed_na, ir, ic, sc["center"], sc["scale"], V)
    X = oracle.read_bed_scaled(obed_na, ir, ic, sc["center"], sc["scale"])
    np.testing.assert_allclose(XV, X @ V, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(rss, (X ** 2).sum(1), rtol=1e-12)
    U = np.linalg.qr(rng.normal(size=(ir.size, 2)))[0]
    t = oracle.multLinReg(obed_na, ir, ic, U, ncores=2)
    G = oracle.read_bed(obed_na, ir, ic, na_val=3).astype(float)
    for j in rng.choice(ic.size, 12, replace=False):
        ok = G[:, j] != 3
        for k in range(2):
            if np.ptp(G[ok, j]) == 0:
                assert np.isnan(t[j, k])
                continue
            r = stats.linregress(U[ok, k], G[ok, j])
            np.testing.assert_allclose(t[j, k], r.slope / r.stderr, rtol=1e-9)

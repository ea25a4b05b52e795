"""GPU parity at BASELINE.json's sizes: the CUDA path (through the C ABI) against the CPU ORACLE -- not against
itself -- on the shapes whose code paths do not exist at fixture size:

* configs[1] in full (50,000 x 500,000): one bed_pMatVec4 and one bed_cpMatVec4 of the oracle's literal port on all host
  threads, both HBM layouts, counts and binomial scaling bit for bit;
* 487,000 x 4,096 and 4,096 x 600,000: every line longer than 262,144 codes is k-split (bsg_pmv.cu MAX_CHUNKS_PER_ITEM),
  which is what every configs[4] cprodVec (n = 487,000) and every configs[1] prodVec (m = 500,000) runs; with and without
  1 % missing values, SNP-major copy alone and both copies;
* a configs[2]-shaped slice (100,000 x 2,000, 500-SNP window, LD-structured data): bed_cor, bed_ld_scores, bed_clumping;
* a configs[3]-shaped slice (10,000 x 20,000): bed_tcrossprodSelf.

The relational form is the reference's own (tests/testthat/test-5-bed-prod-vec.R:18-41: products == dense decode %*% vector,
default and random center / scale); the oracle's loops are that dense product restated literally.  Tolerances are the
north_star's: bit-exact for counts / indices, 1e-6 relative for floating point -- the tests ask for much less
(1e-11 of the vector scale for the products, 1e-10 for r, 1e-8 for K: 28-bit weights).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED_CFG2 = 20250924 + 1


@pytest.fixture(scope="module")
def B():
    import bigsnpr_b200 as b
    from bigsnpr_b200 import build

    build.build()
    return b


def _relerr(got, want, scale):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape
    return float(np.max(np.abs(got - want)) / scale)


def _free_gb():
    import torch

    return torch.cuda.mem_get_info()[0] / 1e9


def _products_vs_oracle(B, oracle, g, o, rng, tol=1e-11, random_scaling=True):
    """bed_prodVec / bed_cprodVec == oracle, with binomial scaling and (reference test :32-39) random center / scale."""
    n, m = o.nrow, o.ncol
    nt = oracle.max_threads()
    sc = B.bed_scaleBinom(g)
    sco = oracle.bed_scaleBinom(o, ncores=nt)
    assert np.array_equal(sc["center"], sco["center"]) and np.array_equal(sc["scale"], sco["scale"])
    y_col, y_row = rng.normal(size=m), rng.normal(size=n)
    worst = 0.0
    scalings = [(sc["center"], sc["scale"])]
    if random_scaling:
        scalings.append((rng.normal(size=m), rng.uniform(0.05, 1.0, size=m)))
    for center, scale in scalings:
        a = B.bed_prodVec(g, y_col, center=center, scale=scale)
        a0 = oracle.bed_prodVec(o, y_col, center=center, scale=scale, ncores=nt)
        b = B.bed_cprodVec(g, y_row, center=center, scale=scale)
        b0 = oracle.bed_cprodVec(o, y_row, center=center, scale=scale, ncores=nt)
        # scale of the sums: every term is bounded by |y| * max(|g - c|) / s
        sa = np.max(np.abs(y_col) * (3 + np.abs(center)) / scale) * np.sqrt(m)
        sb = np.max(np.abs(y_row)) * np.sqrt(n) * np.max((3 + np.abs(center)) / scale)
        ea, eb = _relerr(a, a0, sa), _relerr(b, b0, sb)
        assert ea < tol and eb < tol, (ea, eb)
        # and in the plain sense of the reference's expect_equal (relative to the result's own size)
        assert _relerr(a, a0, np.max(np.abs(a0))) < 1e-9 and _relerr(b, b0, np.max(np.abs(b0))) < 1e-9
        worst = max(worst, ea, eb)
    return worst


def test_cfg2_full_size_vs_oracle(B, oracle, rng):
    """configs[1] in full: 50,000 x 500,000 (6.25 GB packed).  The oracle's generator is the bit-exact twin of the device
    generator, so both sides see the same matrix; ~1-2 s per oracle product on 64 threads."""
    if _free_gb() < 30:
        pytest.skip("needs ~14 GB of HBM")
    n, m = 50_000, 500_000
    o = oracle.synth_bed(n, m, seed=SEED_CFG2)
    nt = oracle.max_threads()
    for layouts in (B.LAYOUT_SNP_MAJOR, B.LAYOUT_SNP_MAJOR | B.LAYOUT_SAMPLE_MAJOR):
        g = B.Bed.synthetic(n, m, seed=SEED_CFG2, layouts=layouts)
        assert g.layouts == layouts
        if layouts == B.LAYOUT_SNP_MAJOR:
            assert np.array_equal(B.bed_counts(g), oracle.bed_col_counts_cpp(o, o.rows_along(), o.cols_along(), nt))
        _products_vs_oracle(B, oracle, g, o, rng, random_scaling=(layouts == B.LAYOUT_SNP_MAJOR))
        g.close()


@pytest.mark.parametrize("shape", [(487_000, 4_096), (4_096, 600_000)])
@pytest.mark.parametrize("na_rate", [0.0, 0.01])
def test_ksplit_shapes_vs_oracle(B, oracle, rng, shape, na_rate):
    """Lines longer than 262,144 codes are split (k-split >= 2) on the side whose contraction is long: Xt.y for
    n = 487,000 (configs[4]'s sample count), X.y for m = 600,000.  Both HBM layouts, with and without missing values."""
    n, m = shape
    o = oracle.synth_bed(n, m, seed=77, na_rate=na_rate)
    for layouts in (B.LAYOUT_SNP_MAJOR, B.LAYOUT_SNP_MAJOR | B.LAYOUT_SAMPLE_MAJOR):
        g = B.Bed.synthetic(n, m, seed=77, na_rate=na_rate, layouts=layouts)
        assert g.has_na == (na_rate > 0)
        _products_vs_oracle(B, oracle, g, o, rng)
        # multiset indices on the long side (duplicates scatter-add in integers)
        ir = rng.integers(1, n + 1, size=min(n, 3000)).astype(np.int32)
        ic = rng.integers(1, m + 1, size=min(m, 3000)).astype(np.int32)
        y = rng.normal(size=ic.size)
        nt = oracle.max_threads()
        a = B.bed_prodVec(g, y, ind_row=ir, ind_col=ic)
        a0 = oracle.bed_prodVec(o, y, ind_row=ir, ind_col=ic, ncores=nt)
        assert _relerr(a, a0, np.max(np.abs(a0)) + 1) < 1e-11
        yr = rng.normal(size=ir.size)
        b = B.bed_cprodVec(g, yr, ind_row=ir, ind_col=ic)
        b0 = oracle.bed_cprodVec(o, yr, ind_row=ir, ind_col=ic, ncores=nt)
        assert _relerr(b, b0, np.max(np.abs(b0)) + 1) < 1e-11
        g.close()


@pytest.mark.parametrize("na_rate", [0.0, 0.005])
def test_cfg3_slice_cor_ld_clumping_vs_oracle(B, oracle, na_rate):
    """configs[2]-shaped slice: 100,000 samples x 2,000 SNPs, 500-SNP window, LD-structured synthetic data (blocks of 50
    correlated SNPs) so thresholds and pruning are exercised.  r is compared value by value with the same sparsity pattern,
    LD scores to 1e-10, clumping indices exactly."""
    n, m = 100_000, 2_000
    kw = dict(seed=31, na_rate=na_rate, ld_rho=0.9, ld_block=50)
    o = oracle.synth_bed(n, m, **kw)
    g = B.Bed.synthetic(n, m, **kw)
    nt = oracle.max_threads()
    pos = 1000.0 * np.arange(1, m + 1)
    for thr_r2 in (0.0, 0.2):
        p, i, x = B.bed_cor(g, size=500, thr_r2=thr_r2, infos_pos=pos)
        p0, i0, x0 = oracle.cor0(o, size=500, thr_r2=thr_r2, infos_pos=pos, ncores=nt)
        assert np.array_equal(p, p0) and np.array_equal(i, i0)
        assert np.allclose(x, x0, rtol=0, atol=1e-12)
        if thr_r2 > 0:
            assert 0 < x.size < 0.5 * m * 500  # the threshold really prunes on this data
    assert np.mean(np.abs(x) > 0.3) > 0.01
    ld = B.bed_ld_scores(g, size=500, infos_pos=pos)
    ld0 = oracle.ld0(o, size=500, infos_pos=pos, ncores=nt)
    assert np.max(np.abs(ld - ld0) / ld0) < 1e-10 and np.max(ld0) > 3
    # clumping on the first 500 SNPs, +-100 SNP window (the oracle's sweep is single-threaded by construction)
    sub = np.arange(1, 501, dtype=np.int32)
    excl = np.arange(501, m + 1)
    chrom = np.ones(m, dtype=int)
    k = B.bed_clumping(g, thr_r2=0.2, size=100, exclude=excl, infos_chr=chrom, infos_pos=pos)
    k0 = oracle.bed_clumping(o, thr_r2=0.2, size=100, exclude=excl, infos_chr=chrom, infos_pos=pos)
    assert np.array_equal(k, k0)
    assert 20 < k.size < sub.size  # pruning happened
    g.close()


def test_cfg4_slice_grm_vs_oracle(B, oracle):
    """configs[3]-shaped slice: 10,000 samples x 20,000 SNPs (78 row tiles x 157 chunks: multi-tile bands and several
    accumulation passes of the Gram kernel).  K against the oracle's block loop (decode + fp64 GEMM)."""
    n, m = 10_000, 20_000
    for na_rate in (0.0, 0.01):
        o = oracle.synth_bed(n, m, seed=41, na_rate=na_rate)
        g = B.Bed.synthetic(n, m, seed=41, na_rate=na_rate)
        K, c, s = B.bed_tcrossprodSelf(g)
        K0, c0, s0 = oracle.bed_tcrossprodSelf(o, block_size=2000)
        assert np.array_equal(c, c0) and np.array_equal(s, s0)
        err = np.max(np.abs(K - K0)) / np.max(np.abs(K0))
        assert err < 1e-8, err  # 28-bit weights (4 base-128 digit slices), exact integer Gram per slice
        assert np.array_equal(K, K.T)
        g.close()


def test_products_on_the_default_stream_are_ordered(B):
    """a NULL stream means the legacy default stream.  The product is enqueued between two torch
    operations on torch's default stream with no synchronisation in between; repeated with fresh inputs it must always see
    the input written just before and be seen by the reduction enqueued just after."""
    import torch

    n, m = 20_000, 40_000
    g = B.Bed.synthetic(n, m, seed=3)
    sc = B.bed_scaleBinom(g)
    v = B.View(g, center=sc["center"], scale=sc["scale"])
    dev = torch.device("cuda", 0)
    base = torch.randn(m, dtype=torch.float64, device=dev)
    ref_out = torch.empty(n, dtype=torch.float64, device=dev)
    v.prodvec_dev(base.data_ptr(), ref_out.data_ptr(), 0)
    torch.cuda.synchronize()
    want = float(ref_out.sum())
    x = torch.zeros(m, dtype=torch.float64, device=dev)
    out = torch.zeros(n, dtype=torch.float64, device=dev)
    for it in range(30):
        big = torch.randn(8_000_000, device=dev).sum()  # keeps the default stream busy before the input is written
        x.copy_(base * (it + 1))                       # produced on torch's default stream ...
        v.prodvec_dev(x.data_ptr(), out.data_ptr(), 0)  # ... consumed by the library on stream NULL
        s = out.sum() / (it + 1)                        # ... and reduced by torch right after
        out.zero_()
        assert abs(float(s) - want) <= 1e-9 * abs(want) + 1e-6, it
        del big
    v.close()
    g.close()

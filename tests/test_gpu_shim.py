"""The R side of the boundary, executed: r_shim/bigsnpr_shim.c is compiled, linked against libbsgpu and a minimal stand-in
for R's C API (tests/stubs/minir.c: vectors, environments, external pointers, .Call by registered name, Rf_error as a
catchable error) and its `.Call` entry points are driven with the objects the reference's R wrappers pass
(R/RcppExports.R:4-78): `bed` environments with `$address`, FBM.code256 environments with `$backingfile` / `$code256`,
the 1 x m integer FBM `keep` of the clumping routines.  Results against the oracle.  This is what a maintainer's
`R CMD SHLIB` build of the shim does with real R; R itself is absent from the image."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
SEXP = C.c_void_p


class MiniR:
    def __init__(self, so):
        L = C.CDLL(so)
        for name, res, args in [
            ("minir_int_vec", SEXP, [C.POINTER(C.c_int), C.c_int]), ("minir_real_vec", SEXP, [C.POINTER(C.c_double), C.c_int]),
            ("minir_real_mat", SEXP, [C.POINTER(C.c_double), C.c_int, C.c_int]), ("minir_lgl", SEXP, [C.c_int]),
            ("minir_str", SEXP, [C.c_char_p]), ("minir_env_new", SEXP, []), ("minir_env_set", None, [SEXP, C.c_char_p, SEXP]),
            ("minir_env_get", SEXP, [SEXP, C.c_char_p]), ("minir_nil", SEXP, []), ("minir_type", C.c_int, [SEXP]),
            ("minir_len", C.c_long, [SEXP]), ("minir_data", C.c_void_p, [SEXP]), ("minir_list_get", SEXP, [SEXP, C.c_int]),
            ("minir_list_by_name", SEXP, [SEXP, C.c_char_p]), ("minir_nrow", C.c_int, [SEXP]), ("minir_ncol", C.c_int, [SEXP]),
            ("minir_last_error", C.c_char_p, []), ("minir_last_warning", C.c_char_p, []), ("minir_warning_count", C.c_int, []),
            ("minir_protect_depth", C.c_int, []), ("minir_run_finalizers", None, []),
            ("minir_dot_call", SEXP, [C.c_char_p, C.c_int, C.POINTER(SEXP)]), ("R_init_bigsnpr_hotpath", None, [C.c_void_p]),
        ]:
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        self.L = L
        L.R_init_bigsnpr_hotpath(None)

    # --- values in
    def ints(self, a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        return self.L.minir_int_vec(a.ctypes.data_as(C.POINTER(C.c_int)), a.size)

    def reals(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        return self.L.minir_real_vec(a.ctypes.data_as(C.POINTER(C.c_double)), a.size)

    def mat(self, a):
        a = np.asfortranarray(a, dtype=np.float64)
        return self.L.minir_real_mat(a.ctypes.data_as(C.POINTER(C.c_double)), a.shape[0], a.shape[1])

    def env(self, **fields):
        e = self.L.minir_env_new()
        for k, v in fields.items():
            self.L.minir_env_set(e, k.encode(), v)
        return e

    def s(self, txt):
        return self.L.minir_str(os.fsencode(txt))

    # --- .Call
    def call(self, name, *args):
        arr = (SEXP * len(args))(*args)
        res = self.L.minir_dot_call(name.encode(), len(args), arr)
        if not res:
            raise RuntimeError(self.L.minir_last_error().decode())
        assert self.L.minir_protect_depth() == 0, "unbalanced PROTECT in " + name
        return res

    # --- values out
    def vec(self, sx):
        n, t = self.L.minir_len(sx), self.L.minir_type(sx)
        if n == 0:
            return np.zeros(0)
        ct = {13: C.c_int, 10: C.c_int, 14: C.c_double, 24: C.c_ubyte}[t]
        a = np.ctypeslib.as_array(C.cast(self.L.minir_data(sx), C.POINTER(ct)), shape=(n,)).copy()
        nr, nc = self.L.minir_nrow(sx), self.L.minir_ncol(sx)
        return a.reshape(nc, nr).T if nr else a

    def named(self, sx, name):
        return self.L.minir_list_by_name(sx, name.encode())


@pytest.fixture(scope="module")
def R(tmp_path_factory):
    from tests.test_abi import build_shim_with_minir

    return MiniR(build_shim_with_minir(tmp_path_factory.mktemp("shim")))


def _bed_env(R, path, n, m):
    xp = R.call("_bigsnpr_bedXPtr", R.s(path), R.ints([n]), R.ints([m]))
    return R.env(address=xp, bedfile=R.s(path), nrow=R.ints([n]), ncol=R.ints([m]))


def test_shim_bed_entry_points(R, oracle, obed_na, rng):
    """bedXPtr, bed_pMatVec4 / bed_cpMatVec4, bed_colstats (+ warning), counts, read_bed(_scaled), prod_and_rowSumsSq,
    multLinReg, corMat, ld_scores and the two collapsed calls through the registered .Call names."""
    path = os.path.join(GOLDEN, "example-missing.bed")
    o, n, m = obed_na, obed_na.nrow, obed_na.ncol
    with pytest.raises(RuntimeError, match="n or p does not match the dimensions of the file."):
        R.call("_bigsnpr_bedXPtr", R.s(path), R.ints([n]), R.ints([m - 1]))
    bed = _bed_env(R, path, n, m)
    ir = (rng.choice(n, 150, replace=False) + 1).astype(np.int32)
    ic = (rng.choice(m, 400, replace=False) + 1).astype(np.int32)
    one = R.ints([1])
    st = R.call("_bigsnpr_bed_colstats", bed, R.ints(ir), R.ints(ic), one)
    sto = oracle.bed_colstats(o, ir, ic)
    for k in ("sumX", "denoX", "nb_nona_col"):
        assert np.array_equal(R.vec(R.named(st, k)), sto[k]), k
    sc = oracle.bed_scaleBinom(o, ir, ic)
    x, y = rng.normal(size=ic.size), rng.normal(size=ir.size)
    a = R.vec(R.call("_bigsnpr_bed_pMatVec4", bed, R.ints(ir), R.ints(ic), R.reals(sc["center"]), R.reals(sc["scale"]), R.reals(x), one))
    b = R.vec(R.call("_bigsnpr_bed_cpMatVec4", bed, R.ints(ir), R.ints(ic), R.reals(sc["center"]), R.reals(sc["scale"]), R.reals(y), one))
    a0 = oracle.bed_pMatVec4(o, ir, ic, sc["center"], sc["scale"], x)
    b0 = oracle.bed_cpMatVec4(o, ir, ic, sc["center"], sc["scale"], y)
    assert np.max(np.abs(a - a0)) < 1e-10 * np.max(np.abs(a0)) and np.max(np.abs(b - b0)) < 1e-10 * np.max(np.abs(b0))
    with pytest.raises(RuntimeError, match="Incompatibility between dimensions."):  # tests/testthat/test-5-bed-prod-vec.R:44-50
        R.call("_bigsnpr_bed_pMatVec4", bed, R.ints(ir), R.ints(ic), R.reals(sc["center"][:-1]), R.reals(sc["scale"]), R.reals(x), one)
    cc = R.vec(R.call("_bigsnpr_bed_col_counts_cpp", bed, R.ints(ir), R.ints(ic), one))
    rc = R.vec(R.call("_bigsnpr_bed_row_counts_cpp", bed, R.ints(ir), R.ints(ic), one))
    assert np.array_equal(cc, oracle.bed_col_counts_cpp(o, ir, ic)) and np.array_equal(rc, oracle.bed_row_counts_cpp(o, ir, ic))
    dense = R.vec(R.call("_bigsnpr_read_bed", bed, R.ints(ir[:50]), R.ints(ic[:60])))
    assert np.array_equal(dense, oracle.read_bed(o, ir[:50], ic[:60]))
    ds = R.vec(R.call("_bigsnpr_read_bed_scaled", bed, R.ints(ir[:50]), R.ints(ic[:60]), R.reals(sc["center"][:60]), R.reals(sc["scale"][:60])))
    assert np.array_equal(ds, oracle.read_bed_scaled(o, ir[:50], ic[:60], sc["center"][:60], sc["scale"][:60]))
    V = rng.normal(size=(ic.size, 3))
    pr = R.call("_bigsnpr_prod_and_rowSumsSq", bed, R.ints(ir), R.ints(ic), R.reals(sc["center"]), R.reals(sc["scale"]), R.mat(V))
    XV0, rss0 = oracle.prod_and_rowSumsSq(o, ir, ic, sc["center"], sc["scale"], V)
    assert np.allclose(R.vec(R.L.minir_list_get(pr, 0)), XV0, rtol=0, atol=1e-8 * np.max(np.abs(XV0)))  # two columns per pass, 30-bit
    assert np.allclose(R.vec(R.L.minir_list_get(pr, 1)), rss0, rtol=1e-11)
    U = np.linalg.qr(rng.normal(size=(ir.size, 2)))[0]
    ts = R.vec(R.call("_bigsnpr_multLinReg", bed, R.ints(ir), R.ints(ic), R.mat(U), one))
    ts0 = oracle.multLinReg(o, ir, ic, U)
    assert np.array_equal(np.isnan(ts), np.isnan(ts0)) and np.allclose(ts[~np.isnan(ts0)], ts0[~np.isnan(ts0)], rtol=1e-6, atol=1e-7)
    # corMat: list of m lists {i, x}; ld_scores
    ics = np.sort(ic[:200])
    pos = 1000.0 * ics
    thr = oracle.cor_thresholds(ir.size, 1.0, 0.04)
    lst = R.call("_bigsnpr_corMat", bed, R.ints(ir), R.ints(ics), R.reals([50e3]), R.reals(thr), R.reals(pos), R.L.minir_lgl(1), one)
    p0, i0, x0 = oracle.corMat(o, ir, ics, 50e3, thr, pos, True)
    for j in (0, 7, 199):
        el = R.L.minir_list_get(lst, j)
        assert np.array_equal(R.vec(R.named(el, "i")), i0[p0[j]:p0[j + 1]])
        assert np.array_equal(R.vec(R.named(el, "x")), x0[p0[j]:p0[j + 1]], equal_nan=True)
    ld = R.vec(R.call("_bigsnpr_ld_scores", bed, R.ints(ir), R.ints(ics), R.reals([50e3]), R.reals(pos), one))
    assert np.allclose(ld, oracle.ld_scores(o, ir, ics, 50e3, pos), rtol=1e-12)
    # collapsed calls: GRM and SVD
    alli = np.arange(1, n + 1, dtype=np.int32)
    sub = np.arange(1, m + 1, 5, dtype=np.int32)
    sc2 = oracle.bed_scaleBinom(o, alli, sub)
    K = R.vec(R.call("_bigsnpr_bed_tcrossprod_gpu", bed, R.ints(alli), R.ints(sub), R.reals(sc2["center"]), R.reals(sc2["scale"])))
    K0, _, _ = oracle.bed_tcrossprodSelf(o, ind_col=sub)
    assert np.max(np.abs(K - K0)) < 1e-8 * np.max(np.abs(K0))
    sv = R.call("_bigsnpr_bed_randomSVD_gpu", bed, R.ints(alli), R.ints(sub), R.L.minir_nil(), R.L.minir_nil(), R.ints([5]), R.reals([1e-4]))
    d = R.vec(R.named(sv, "d"))
    assert np.max(np.abs(d - np.sqrt(np.linalg.eigvalsh(K0)[::-1][:5])) / d) < 1e-7  # tests/testthat/test-2-bed-clumping-SVD.R:76-78
    # a one-device group through the same table (multi-GPU entry points of the R side)
    grp = R.call("_bigsnpr_bed_group_gpu", R.s(path), R.ints([n]), R.ints([m]), R.ints([0]))
    ag = R.vec(R.call("_bigsnpr_group_pMatVec4_gpu", grp, R.ints(ir), R.ints(ic), R.reals(sc["center"]), R.reals(sc["scale"]), R.reals(x), R.L.minir_lgl(0)))
    bg = R.vec(R.call("_bigsnpr_group_pMatVec4_gpu", grp, R.ints(ir), R.ints(ic), R.reals(sc["center"]), R.reals(sc["scale"]), R.reals(y), R.L.minir_lgl(1)))
    assert np.array_equal(ag, a) and np.array_equal(bg, b)
    Kg = R.vec(R.call("_bigsnpr_group_tcrossprod_gpu", grp, R.ints(alli), R.ints(sub), R.reals(sc2["center"]), R.reals(sc2["scale"])))
    assert np.max(np.abs(Kg - K0)) < 1e-8 * np.max(np.abs(K0))
    with pytest.raises(RuntimeError, match="Unknown object type."):  # src/corr.cpp:124
        R.call("_bigsnpr_ld_scores", R.env(nrow=R.ints([1])), R.ints(ir), R.ints(ics), R.reals([50e3]), R.reals(pos), one)
    R.L.minir_run_finalizers()  # what R's GC does with the external pointers


def test_shim_fbm_entry_points(R, oracle, obed, rng, tmp_path):
    """The entry points that take an FBM (`snp_clumping`, `snp_writeBed`, `snp_pcadapt` crashed by design in the
    round-1 shim): the FBM.code256 is an environment with $backingfile / $nrow / $ncol / $code256, `keep` is a 1 x m integer
    FBM written in place, readbina2 fills a raw FBM in place."""
    o = obed
    n, m = o.nrow, o.ncol
    G = oracle.decode_dense(o).astype(np.uint8)  # no missing value in example.bed
    bk = tmp_path / "geno.bk"
    np.asfortranarray(G).T.tofile(bk)  # column-major n x m bytes
    code = np.full(256, np.nan)
    code[:3] = [0, 1, 2]
    fbm = R.env(backingfile=R.s(str(bk)), nrow=R.ints([n]), ncol=R.ints([m]), code256=R.reals(code))
    of = oracle.OracleFBM(G)
    ir = np.arange(1, n + 1, dtype=np.int32)
    ic = np.arange(1, 801, dtype=np.int32)
    one = R.ints([1])
    st = R.call("_bigsnpr_snp_colstats", fbm, R.ints(ir), R.ints(ic), one)
    st0 = oracle.snp_colstats(of, ir, ic)
    sumX, denoX = R.vec(R.named(st, "sumX")), R.vec(R.named(st, "denoX"))
    assert np.array_equal(sumX, st0["sumX"]) and np.array_equal(denoX, st0["denoX"])
    # clumping_chr on the FBM: keep is an integer FBM initialised to -1 (R/clumping.R:115)
    af = sumX / (2 * n)
    S = np.minimum(af, 1 - af)
    ordv = (np.argsort(-S, kind="stable") + 1).astype(np.int32)
    rank = np.empty_like(ordv)
    rank[ordv - 1] = np.arange(1, ordv.size + 1)
    pos = np.arange(1, ic.size + 1, dtype=np.float64)
    kb = tmp_path / "keep.bk"
    np.full(ic.size, -1, dtype=np.int32).tofile(kb)
    keep = R.env(backingfile=R.s(str(kb)), nrow=R.ints([1]), ncol=R.ints([ic.size]))
    R.call("_bigsnpr_clumping_chr", fbm, keep, R.ints(ir), R.ints(ic), R.ints(ordv), R.ints(rank), R.reals(pos), R.reals(sumX),
           R.reals(denoX), R.reals([50.0]), R.reals([0.2]), one)
    got = np.fromfile(kb, dtype=np.int32)
    want = oracle.clumping_chr(of, ir, ic, ordv, rank, pos, sumX, denoX, 50.0, 0.2)
    assert np.array_equal(got, want) and set(np.unique(got)) <= {0, 1} and 0 < got.sum() < got.size
    # bed_clumping_chr with the same kind of `keep`
    path = os.path.join(GOLDEN, "example.bed")
    bed = _bed_env(R, path, n, m)
    stb = oracle.bed_colstats(o, ir, ic)
    center, scale = stb["sumX"] / stb["nb_nona_col"], np.sqrt(stb["denoX"])
    np.full(ic.size, -1, dtype=np.int32).tofile(kb)
    R.call("_bigsnpr_bed_clumping_chr", bed, keep, R.ints(ir), R.ints(ic), R.reals(center), R.reals(scale), R.ints(ordv), R.ints(rank),
           R.reals(1000.0 * pos), R.reals([50e3]), R.reals([0.2]), one)
    want_b = oracle.bed_clumping_chr(o, ir, ic, center, scale, ordv, rank, 1000.0 * pos, 50e3, 0.2)
    assert np.array_equal(np.fromfile(kb, dtype=np.int32), want_b)
    # writebina from the FBM (snp_writeBed) -> bytes of the original file; readbina2 back into a raw FBM
    outbed = tmp_path / "out.bed"
    R.call("_bigsnpr_writebina", R.s(str(outbed)), fbm, R.ints(np.zeros(1, dtype=np.int32)), R.ints(ir), R.ints(np.arange(1, m + 1)))
    raw = np.frombuffer(outbed.read_bytes(), dtype=np.uint8)
    assert bytes(raw[:3]) == bytes([108, 27, 1])
    # the reference's writer pads the last byte of a column with genotype 0 (code 11), PLINK pads with 00: compare with the
    # oracle's restatement of src/write-plink.cpp:29-47, and decode back to the same genotypes
    assert np.array_equal(raw[3:].reshape(m, -1), oracle.write_bed_bytes(G))
    assert np.array_equal(oracle.decode_dense(oracle.OracleBed.from_packed(raw[3:], n, m)), G)
    sub_r, sub_c = ir[::3], np.arange(1, m + 1, 7, dtype=np.int32)
    rb = tmp_path / "read.bk"
    np.zeros(sub_r.size * sub_c.size, dtype=np.uint8).tofile(rb)
    dst = R.env(backingfile=R.s(str(rb)), nrow=R.ints([sub_r.size]), ncol=R.ints([sub_c.size]))
    R.call("_bigsnpr_readbina2", dst, bed, R.ints(sub_r), R.ints(sub_c), one)
    back = np.fromfile(rb, dtype=np.uint8).reshape(sub_c.size, sub_r.size).T
    assert np.array_equal(back, G[np.ix_(sub_r - 1, sub_c - 1)])
    # multLinReg and corMat dispatch on the FBM (src/multLinReg.cpp:72-78, src/corr.cpp:113-118)
    U = np.linalg.qr(rng.normal(size=(n, 2)))[0]
    ts = R.vec(R.call("_bigsnpr_multLinReg", fbm, R.ints(ir), R.ints(ic), R.mat(U), one))
    ts0 = oracle.multLinReg(of, ir, ic, U)
    assert np.array_equal(np.isnan(ts), np.isnan(ts0)) and np.allclose(ts[~np.isnan(ts0)], ts0[~np.isnan(ts0)], rtol=1e-6, atol=1e-7)
    ld = R.vec(R.call("_bigsnpr_ld_scores", fbm, R.ints(ir), R.ints(ic), R.reals([100.0]), R.reals(pos), one))
    assert np.allclose(ld, oracle.ld_scores(of, ir, ic, 100.0, pos), rtol=1e-12)
    R.L.minir_run_finalizers()

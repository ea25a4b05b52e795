/*
 * bigsnpr_shim.c -- the R side of the drop-in boundary.
 *
 * Replaces, inside the bigsnpr package, the generated src/RcppExports.cpp entries of the hot path and the
 * C++ files behind them (src/bed-acc-xptr.cpp, src/bed-prod-vec.cpp, src/bed-fun.cpp, src/bed-mat-acc.cpp,
 * src/corr.cpp, src/ld-scores.cpp, src/colstats.cpp).  Every function keeps the registered name and arity
 * of the reference (src/RcppExports.cpp:597-640), pulls plain pointers out of the SEXPs and calls libbsgpu
 * (include/bsgpu.h).  R's own wrappers (R/RcppExports.R, R/bed-mult-vec.R, R/binom-scaling.R, R/corr.R,
 * R/ld-scores.R) run unchanged.  `ncores` is accepted and ignored.
 *
 * Build (inside the package, needs R headers -- not available in the CUDA build image, so this file is
 * not compiled there):   PKG_LIBS = -L<dir> -lbsgpu     PKG_CPPFLAGS = -I<repo>/include
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

#include "bsgpu.h"

static void chk(int rc) {
  if (rc) Rf_error("%s", bsg_last_error()); /* same texts as the reference, e.g. "Incompatibility between dimensions." */
}

/* obj$name for RefClass objects / environments (active bindings are evaluated, R/bed-class.R:105-110) */
static SEXP field(SEXP obj, const char *name) {
  SEXP call = PROTECT(Rf_lang3(Rf_install("$"), obj, Rf_mkString(name)));
  SEXP val = Rf_eval(call, R_GlobalEnv);
  UNPROTECT(1);
  return val;
}

static void bed_finalizer(SEXP xp) {
  bsg_bed *h = (bsg_bed *)R_ExternalPtrAddr(xp);
  if (h) bsg_close(h);
  R_ClearExternalPtr(xp);
}

static int gpu_device(void) {
  /* options(bigsnpr.gpu.device = k), default 0 */
  SEXP o = Rf_GetOption1(Rf_install("bigsnpr.gpu.device"));
  return (o == R_NilValue) ? 0 : Rf_asInteger(o);
}

/* _bigsnpr_bedXPtr(path, n, p): src/bed-acc-xptr.cpp:40-55.  Validation + staging to HBM. */
SEXP _bigsnpr_bedXPtr(SEXP path, SEXP n, SEXP p) {
  bsg_bed *h = NULL;
  int m = Rf_asInteger(p);
  chk(bsg_open_bed(CHAR(STRING_ELT(path, 0)), Rf_asInteger(n), m, 0, m, gpu_device(), BSG_LAYOUT_AUTO, &h));
  SEXP xp = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(xp, bed_finalizer, TRUE);
  UNPROTECT(1);
  return xp;
}

static bsg_bed *handle_of(SEXP obj_bed) {
  SEXP xp = field(obj_bed, "address"); /* lazily re-opens in PSOCK workers, R/bed-class.R:187-192 */
  bsg_bed *h = (bsg_bed *)R_ExternalPtrAddr(xp);
  if (!h) Rf_error("external pointer is not valid");
  return h;
}

/* _bigsnpr_bed_pMatVec4: src/bed-prod-vec.cpp:15-54 */
SEXP _bigsnpr_bed_pMatVec4(SEXP obj_bed, SEXP ind_row, SEXP ind_col, SEXP center, SEXP scale, SEXP x, SEXP ncores) {
  bsg_bed *h = handle_of(obj_bed);
  int nr = LENGTH(ind_row), nc = LENGTH(ind_col);
  if (LENGTH(center) != nc || LENGTH(scale) != nc) Rf_error("Incompatibility between dimensions.");
  SEXP out = PROTECT(Rf_allocVector(REALSXP, nr));
  chk(bsg_prodvec(h, INTEGER(ind_row), nr, INTEGER(ind_col), nc, REAL(center), REAL(scale), REAL(x), REAL(out)));
  UNPROTECT(1);
  return out;
}

/* _bigsnpr_bed_cpMatVec4: src/bed-prod-vec.cpp:59-97 */
SEXP _bigsnpr_bed_cpMatVec4(SEXP obj_bed, SEXP ind_row, SEXP ind_col, SEXP center, SEXP scale, SEXP x, SEXP ncores) {
  bsg_bed *h = handle_of(obj_bed);
  int nr = LENGTH(ind_row), nc = LENGTH(ind_col);
  if (LENGTH(center) != nc || LENGTH(scale) != nc) Rf_error("Incompatibility between dimensions.");
  SEXP out = PROTECT(Rf_allocVector(REALSXP, nc));
  chk(bsg_cprodvec(h, INTEGER(ind_row), nr, INTEGER(ind_col), nc, REAL(center), REAL(scale), REAL(x), REAL(out)));
  UNPROTECT(1);
  return out;
}

/* _bigsnpr_bed_colstats: src/bed-fun.cpp:9-46 -> list(sumX, denoX, nb_nona_col) + the >50% warning */
SEXP _bigsnpr_bed_colstats(SEXP obj_bed, SEXP ind_row, SEXP ind_col, SEXP ncores) {
  bsg_bed *h = handle_of(obj_bed);
  int nr = LENGTH(ind_row), nc = LENGTH(ind_col), n_bad = 0;
  SEXP sumX = PROTECT(Rf_allocVector(REALSXP, nc)), denoX = PROTECT(Rf_allocVector(REALSXP, nc));
  SEXP nona = PROTECT(Rf_allocVector(INTSXP, nc));
  chk(bsg_colstats(h, INTEGER(ind_row), nr, INTEGER(ind_col), nc, REAL(sumX), REAL(denoX), INTEGER(nona), &n_bad));
  if (n_bad > 0) Rf_warning("%d variants have >50%% missing values.", n_bad);
  SEXP res = PROTECT(Rf_allocVector(VECSXP, 3)), nm = PROTECT(Rf_allocVector(STRSXP, 3));
  SET_VECTOR_ELT(res, 0, sumX); SET_VECTOR_ELT(res, 1, denoX); SET_VECTOR_ELT(res, 2, nona);
  SET_STRING_ELT(nm, 0, Rf_mkChar("sumX")); SET_STRING_ELT(nm, 1, Rf_mkChar("denoX"));
  SET_STRING_ELT(nm, 2, Rf_mkChar("nb_nona_col"));
  Rf_setAttrib(res, R_NamesSymbol, nm);
  UNPROTECT(5);
  return res;
}

/* _bigsnpr_bed_col_counts_cpp / _bigsnpr_bed_row_counts_cpp: src/bed-fun.cpp:51-69, :72-98 */
static SEXP counts(SEXP obj_bed, SEXP ind_row, SEXP ind_col, int byrow) {
  bsg_bed *h = handle_of(obj_bed);
  int nr = LENGTH(ind_row), nc = LENGTH(ind_col), k = byrow ? nr : nc;
  SEXP res = PROTECT(Rf_allocMatrix(INTSXP, 4, k));
  chk((byrow ? bsg_row_counts : bsg_col_counts)(h, INTEGER(ind_row), nr, INTEGER(ind_col), nc, INTEGER(res)));
  UNPROTECT(1);
  return res;
}
SEXP _bigsnpr_bed_col_counts_cpp(SEXP o, SEXP r, SEXP c, SEXP ncores) { return counts(o, r, c, 0); }
SEXP _bigsnpr_bed_row_counts_cpp(SEXP o, SEXP r, SEXP c, SEXP ncores) { return counts(o, r, c, 1); }

/* _bigsnpr_read_bed: src/bed-mat-acc.cpp:8-26 */
SEXP _bigsnpr_read_bed(SEXP obj_bed, SEXP ind_row, SEXP ind_col) {
  bsg_bed *h = handle_of(obj_bed);
  int nr = LENGTH(ind_row), nc = LENGTH(ind_col);
  SEXP res = PROTECT(Rf_allocMatrix(INTSXP, nr, nc));
  chk(bsg_read_bed(h, INTEGER(ind_row), nr, INTEGER(ind_col), nc, NA_INTEGER, INTEGER(res)));
  UNPROTECT(1);
  return res;
}

/* _bigsnpr_read_bed_scaled: src/bed-mat-acc.cpp:30-49 */
SEXP _bigsnpr_read_bed_scaled(SEXP obj_bed, SEXP ind_row, SEXP ind_col, SEXP center, SEXP scale) {
  bsg_bed *h = handle_of(obj_bed);
  int nr = LENGTH(ind_row), nc = LENGTH(ind_col);
  if (LENGTH(center) != nc || LENGTH(scale) != nc) Rf_error("Incompatibility between dimensions.");
  SEXP res = PROTECT(Rf_allocMatrix(REALSXP, nr, nc));
  chk(bsg_read_bed_scaled(h, INTEGER(ind_row), nr, INTEGER(ind_col), nc, REAL(center), REAL(scale), REAL(res)));
  UNPROTECT(1);
  return res;
}

/* FBM.code256 objects (snp_cor / snp_ld_scores / snp_colstats): the raw n x m bytes live in the memory-mapped
 * backing file; a handle is opened once per object and cached in the environment (field ".bsg"). */
static bsg_bed *fbm_handle_of(SEXP obj) {
  SEXP cached = Rf_findVarInFrame(obj, Rf_install(".bsg"));
  if (cached != R_UnboundValue && TYPEOF(cached) == EXTPTRSXP && R_ExternalPtrAddr(cached))
    return (bsg_bed *)R_ExternalPtrAddr(cached);
  int n = Rf_asInteger(field(obj, "nrow")), m = Rf_asInteger(field(obj, "ncol"));
  SEXP code = PROTECT(Rf_coerceVector(field(obj, "code256"), REALSXP));
  /* G[] as raw: the shim reads the bytes through R so it does not depend on bigstatsr's C++ classes */
  SEXP call = PROTECT(Rf_lang2(Rf_install("as.raw.FBM.bytes"), obj)); /* helper exported by the R glue, INTEGRATION.md */
  SEXP bytes = PROTECT(Rf_eval(call, R_GlobalEnv));
  bsg_bed *h = NULL;
  chk(bsg_open_fbm256(RAW(bytes), n, m, REAL(code), gpu_device(), BSG_LAYOUT_SNP_MAJOR, &h));
  SEXP xp = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(xp, bed_finalizer, TRUE);
  Rf_defineVar(Rf_install(".bsg"), xp, obj);
  UNPROTECT(4);
  return h;
}

static int has_field(SEXP obj, const char *name) { return Rf_findVarInFrame(obj, Rf_install(name)) != R_UnboundValue; }

/* dispatch of src/corr.cpp:113-125: "code256" -> FBM, "bedfile" -> bed, else "Unknown object type." */
static bsg_bed *any_handle(SEXP obj) {
  if (has_field(obj, "code256")) return fbm_handle_of(obj);
  if (has_field(obj, "bedfile")) return handle_of(obj);
  Rf_error("Unknown object type.");
  return NULL;
}

/* _bigsnpr_snp_colstats: src/colstats.cpp:8-35 */
SEXP _bigsnpr_snp_colstats(SEXP BM, SEXP rowInd, SEXP colInd, SEXP ncores) {
  bsg_bed *h = fbm_handle_of(BM);
  int nr = LENGTH(rowInd), nc = LENGTH(colInd);
  SEXP sumX = PROTECT(Rf_allocVector(REALSXP, nc)), denoX = PROTECT(Rf_allocVector(REALSXP, nc));
  chk(bsg_snp_colstats(h, INTEGER(rowInd), nr, INTEGER(colInd), nc, REAL(sumX), REAL(denoX)));
  SEXP res = PROTECT(Rf_allocVector(VECSXP, 2)), nm = PROTECT(Rf_allocVector(STRSXP, 2));
  SET_VECTOR_ELT(res, 0, sumX); SET_VECTOR_ELT(res, 1, denoX);
  SET_STRING_ELT(nm, 0, Rf_mkChar("sumX")); SET_STRING_ELT(nm, 1, Rf_mkChar("denoX"));
  Rf_setAttrib(res, R_NamesSymbol, nm);
  UNPROTECT(4);
  return res;
}

/* _bigsnpr_corMat: src/corr.cpp:102-126 -> list of m lists {i, x} (R/corr.R:43-47 assembles the dsCMatrix) */
SEXP _bigsnpr_corMat(SEXP obj, SEXP rowInd, SEXP colInd, SEXP size, SEXP thr, SEXP pos, SEXP fill_diag, SEXP ncores) {
  int nr = LENGTH(rowInd), nc = LENGTH(colInd);
  if (LENGTH(pos) != nc) Rf_error("Incompatibility between dimensions.");
  bsg_bed *h = any_handle(obj);
  int64_t *p = (int64_t *)R_alloc((size_t)nc + 1, sizeof(int64_t));
  int *ci = NULL;
  double *cx = NULL;
  chk(bsg_cor(h, INTEGER(rowInd), nr, INTEGER(colInd), nc, Rf_asReal(size), REAL(thr), REAL(pos), Rf_asLogical(fill_diag),
              p, &ci, &cx));
  SEXP res = PROTECT(Rf_allocVector(VECSXP, nc));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, 2));
  SET_STRING_ELT(nm, 0, Rf_mkChar("i")); SET_STRING_ELT(nm, 1, Rf_mkChar("x"));
  for (int j = 0; j < nc; j++) {
    int len = (int)(p[j + 1] - p[j]);
    SEXP el = PROTECT(Rf_allocVector(VECSXP, 2)), vi = PROTECT(Rf_allocVector(INTSXP, len));
    SEXP vx = PROTECT(Rf_allocVector(REALSXP, len));
    for (int k = 0; k < len; k++) { INTEGER(vi)[k] = ci[p[j] + k]; REAL(vx)[k] = cx[p[j] + k]; }
    SET_VECTOR_ELT(el, 0, vi); SET_VECTOR_ELT(el, 1, vx);
    Rf_setAttrib(el, R_NamesSymbol, nm);
    SET_VECTOR_ELT(res, j, el);
    UNPROTECT(3);
  }
  bsg_free(ci); bsg_free(cx);
  UNPROTECT(2);
  return res;
}

/* _bigsnpr_ld_scores: src/ld-scores.cpp:83-105 */
SEXP _bigsnpr_ld_scores(SEXP obj, SEXP rowInd, SEXP colInd, SEXP size, SEXP pos, SEXP ncores) {
  int nr = LENGTH(rowInd), nc = LENGTH(colInd);
  if (LENGTH(pos) != nc) Rf_error("Incompatibility between dimensions.");
  bsg_bed *h = any_handle(obj);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, nc));
  chk(bsg_ld_scores(h, INTEGER(rowInd), nr, INTEGER(colInd), nc, Rf_asReal(size), REAL(pos), REAL(out)));
  UNPROTECT(1);
  return out;
}

/* _bigsnpr_bed_clumping_chr: src/clumping-bed.cpp:11-91 (12 arguments, keep is the integer FBM BM2 written in place;
 * rankInd is implied by ordInd).  BM2$address_rw is bigstatsr's XPtr<FBM_RW>: the shim writes through as.integer
 * storage obtained from R -- here via the helper `fbm_int_ptr`. */
extern int *fbm_int_ptr(SEXP BM2); /* provided by the package glue: pointer to the mmap'ed int matrix of an FBM */
SEXP _bigsnpr_bed_clumping_chr(SEXP obj_bed, SEXP BM2, SEXP ind_row, SEXP ind_col, SEXP center, SEXP scale, SEXP ordInd,
                               SEXP rankInd, SEXP pos, SEXP size, SEXP thr, SEXP ncores) {
  bsg_bed *h = handle_of(obj_bed);
  int nr = LENGTH(ind_row), nc = LENGTH(ind_col);
  chk(bsg_clumping_chr(h, INTEGER(ind_row), nr, INTEGER(ind_col), nc, REAL(center), REAL(scale), INTEGER(ordInd), REAL(pos),
                       Rf_asReal(size), Rf_asReal(thr), fbm_int_ptr(BM2)));
  return R_NilValue;
}

/* _bigsnpr_clumping_chr: src/clumping.cpp:10-91 (12 arguments; BM is the FBM.code256 environment) */
SEXP _bigsnpr_clumping_chr(SEXP BM, SEXP BM2, SEXP rowInd, SEXP colInd, SEXP ordInd, SEXP rankInd, SEXP pos, SEXP sumX,
                           SEXP denoX, SEXP size, SEXP thr, SEXP ncores) {
  bsg_bed *h = handle_of(BM);
  int nr = LENGTH(rowInd), nc = LENGTH(colInd);
  chk(bsg_clumping_chr_fbm(h, INTEGER(rowInd), nr, INTEGER(colInd), nc, REAL(sumX), REAL(denoX), INTEGER(ordInd), REAL(pos),
                           Rf_asReal(size), Rf_asReal(thr), fbm_int_ptr(BM2)));
  return R_NilValue;
}

/* _bigsnpr_readbina2: src/read-plink.cpp:61-80 (5 arguments; BM is the destination FBM.code256, filled in place) */
extern unsigned char *fbm_raw_ptr(SEXP BM); /* package glue: pointer to the mmap'ed raw matrix of an FBM */
SEXP _bigsnpr_readbina2(SEXP BM, SEXP obj_bed, SEXP ind_row, SEXP ind_col, SEXP ncores) {
  bsg_bed *h = handle_of(obj_bed);
  chk(bsg_readbina2(h, INTEGER(ind_row), LENGTH(ind_row), INTEGER(ind_col), LENGTH(ind_col), fbm_raw_ptr(BM)));
  return R_NilValue;
}

/* _bigsnpr_writebina: src/write-plink.cpp:13-52 (5 arguments; `tab` = getInverseCode() is implied by the library) */
SEXP _bigsnpr_writebina(SEXP filename, SEXP BM, SEXP tab, SEXP rowInd, SEXP colInd) {
  bsg_bed *h = handle_of(BM);
  chk(bsg_writebina(h, CHAR(STRING_ELT(filename, 0)), INTEGER(rowInd), LENGTH(rowInd), INTEGER(colInd), LENGTH(colInd)));
  return R_NilValue;
}

/* _bigsnpr_prod_and_rowSumsSq: src/bed-fun.cpp:103-133 (6 arguments) -> list(XV, rowSumsSq) */
SEXP _bigsnpr_prod_and_rowSumsSq(SEXP obj_bed, SEXP ind_row, SEXP ind_col, SEXP center, SEXP scale, SEXP V) {
  bsg_bed *h = handle_of(obj_bed);
  int nr = LENGTH(ind_row), nc = LENGTH(ind_col), K = Rf_ncols(V);
  if (LENGTH(center) != nc || LENGTH(scale) != nc || Rf_nrows(V) != nc) Rf_error("Incompatibility between dimensions.");
  SEXP XV = PROTECT(Rf_allocMatrix(REALSXP, nr, K)), rss = PROTECT(Rf_allocVector(REALSXP, nr));
  chk(bsg_prod_and_rowsumssq(h, INTEGER(ind_row), nr, INTEGER(ind_col), nc, REAL(center), REAL(scale), REAL(V), K, REAL(XV),
                             REAL(rss)));
  SEXP res = PROTECT(Rf_allocVector(VECSXP, 2));
  SET_VECTOR_ELT(res, 0, XV);
  SET_VECTOR_ELT(res, 1, rss);
  UNPROTECT(3);
  return res;
}

/* _bigsnpr_multLinReg: src/multLinReg.cpp:64-88 (5 arguments; `obj` is a bed or an FBM.code256 environment) */
SEXP _bigsnpr_multLinReg(SEXP obj, SEXP ind_row, SEXP ind_col, SEXP U, SEXP ncores) {
  bsg_bed *h = handle_of(obj);
  int nr = LENGTH(ind_row), nc = LENGTH(ind_col), K = Rf_ncols(U);
  if (Rf_nrows(U) != nr) Rf_error("Incompatibility between dimensions.");
  SEXP t = PROTECT(Rf_allocMatrix(REALSXP, nc, K));
  chk(bsg_multlinreg(h, INTEGER(ind_row), nr, INTEGER(ind_col), nc, REAL(U), K, REAL(t)));
  for (R_xlen_t i = 0; i < XLENGTH(t); i++)
    if (ISNAN(REAL(t)[i])) REAL(t)[i] = NA_REAL; /* the library writes NaN where the reference writes NA_REAL */
  UNPROTECT(1);
  return t;
}

/* new entry points: R/bed-tcrossprodSelf.R's block loop and R/autoSVD.R's bed_randomSVD collapse to one call each */
SEXP _bigsnpr_bed_tcrossprod_gpu(SEXP obj_bed, SEXP ind_row, SEXP ind_col, SEXP center, SEXP scale) {
  bsg_bed *h = handle_of(obj_bed);
  int nr = LENGTH(ind_row), nc = LENGTH(ind_col);
  SEXP K = PROTECT(Rf_allocMatrix(REALSXP, nr, nr));
  chk(bsg_tcrossprod(h, INTEGER(ind_row), nr, INTEGER(ind_col), nc, REAL(center), REAL(scale), REAL(K)));
  UNPROTECT(1);
  return K;
}

SEXP _bigsnpr_bed_randomSVD_gpu(SEXP obj_bed, SEXP ind_row, SEXP ind_col, SEXP center, SEXP scale, SEXP k, SEXP tol) {
  bsg_bed *h = handle_of(obj_bed);
  int nr = LENGTH(ind_row), nc = LENGTH(ind_col), kk = Rf_asInteger(k), niter = 0, nops = 0;
  SEXP d = PROTECT(Rf_allocVector(REALSXP, kk)), u = PROTECT(Rf_allocMatrix(REALSXP, nr, kk));
  SEXP v = PROTECT(Rf_allocMatrix(REALSXP, nc, kk));
  SEXP co = PROTECT(Rf_allocVector(REALSXP, nc)), so = PROTECT(Rf_allocVector(REALSXP, nc));
  const double *cen = (center == R_NilValue) ? NULL : REAL(center), *sca = (scale == R_NilValue) ? NULL : REAL(scale);
  chk(bsg_randomsvd(h, INTEGER(ind_row), nr, INTEGER(ind_col), nc, cen, sca, kk, Rf_asReal(tol), 1000, REAL(d), REAL(u),
                    REAL(v), REAL(co), REAL(so), &niter, &nops));
  const char *names[] = {"d", "u", "v", "niter", "nops", "center", "scale", ""};
  SEXP res = PROTECT(Rf_mkNamed(VECSXP, names));
  SET_VECTOR_ELT(res, 0, d); SET_VECTOR_ELT(res, 1, u); SET_VECTOR_ELT(res, 2, v);
  SET_VECTOR_ELT(res, 3, Rf_ScalarInteger(niter)); SET_VECTOR_ELT(res, 4, Rf_ScalarInteger(nops));
  SET_VECTOR_ELT(res, 5, co); SET_VECTOR_ELT(res, 6, so);
  Rf_setAttrib(res, R_ClassSymbol, Rf_mkString("big_SVD"));
  UNPROTECT(6);
  return res;
}

/* registration: same table shape as src/RcppExports.cpp:597-640 (only the hot-path rows shown; the other
 * entries of the reference stay as generated) */
static const R_CallMethodDef CallEntries[] = {
    {"_bigsnpr_bedXPtr", (DL_FUNC)&_bigsnpr_bedXPtr, 3},
    {"_bigsnpr_bed_colstats", (DL_FUNC)&_bigsnpr_bed_colstats, 4},
    {"_bigsnpr_bed_col_counts_cpp", (DL_FUNC)&_bigsnpr_bed_col_counts_cpp, 4},
    {"_bigsnpr_bed_row_counts_cpp", (DL_FUNC)&_bigsnpr_bed_row_counts_cpp, 4},
    {"_bigsnpr_read_bed", (DL_FUNC)&_bigsnpr_read_bed, 3},
    {"_bigsnpr_read_bed_scaled", (DL_FUNC)&_bigsnpr_read_bed_scaled, 5},
    {"_bigsnpr_bed_pMatVec4", (DL_FUNC)&_bigsnpr_bed_pMatVec4, 7},
    {"_bigsnpr_bed_cpMatVec4", (DL_FUNC)&_bigsnpr_bed_cpMatVec4, 7},
    {"_bigsnpr_snp_colstats", (DL_FUNC)&_bigsnpr_snp_colstats, 4},
    {"_bigsnpr_corMat", (DL_FUNC)&_bigsnpr_corMat, 8},
    {"_bigsnpr_ld_scores", (DL_FUNC)&_bigsnpr_ld_scores, 6},
    {"_bigsnpr_bed_clumping_chr", (DL_FUNC)&_bigsnpr_bed_clumping_chr, 12},
    {"_bigsnpr_clumping_chr", (DL_FUNC)&_bigsnpr_clumping_chr, 12},
    {"_bigsnpr_readbina2", (DL_FUNC)&_bigsnpr_readbina2, 5},
    {"_bigsnpr_writebina", (DL_FUNC)&_bigsnpr_writebina, 5},
    {"_bigsnpr_prod_and_rowSumsSq", (DL_FUNC)&_bigsnpr_prod_and_rowSumsSq, 6},
    {"_bigsnpr_multLinReg", (DL_FUNC)&_bigsnpr_multLinReg, 5},
    {"_bigsnpr_bed_tcrossprod_gpu", (DL_FUNC)&_bigsnpr_bed_tcrossprod_gpu, 5},
    {"_bigsnpr_bed_randomSVD_gpu", (DL_FUNC)&_bigsnpr_bed_randomSVD_gpu, 7},
    {NULL, NULL, 0}};

void R_init_bigsnpr_hotpath(DllInfo *dll) {
  /* in the package this table is merged into R_init_bigsnpr (src/RcppExports.cpp:637-640) */
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}

/* Minimal declarations of the public R C API used by r_shim/bigsnpr_shim.c -- ONLY for a syntax / type check of the
 * shim against include/bsgpu.h in an image without R (tests/test_abi.py).  Not R, not linkable, not shipped. */
#ifndef STUB_RINTERNALS_H
#define STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE } Rboolean;
#define NILSXP 0
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define EXTPTRSXP 22
#define RAWSXP 24
extern SEXP R_NilValue, R_GlobalEnv, R_UnboundValue, R_NamesSymbol, R_ClassSymbol;
extern double R_NaReal;
extern int R_NaInt;
#define NA_INTEGER R_NaInt
#define NA_LOGICAL R_NaInt
#define NA_REAL R_NaReal
int R_IsNaN(double);
int R_isnancpp(double);
#define ISNAN(x) (R_IsNaN(x) || (x) != (x))
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
SEXP Rf_allocVector(unsigned int, R_xlen_t);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP Rf_coerceVector(SEXP, unsigned int);
SEXP Rf_install(const char *);
SEXP Rf_mkChar(const char *);
SEXP Rf_mkString(const char *);
SEXP Rf_mkNamed(unsigned int, const char **);
SEXP Rf_lang1(SEXP);
SEXP Rf_lang2(SEXP, SEXP);
SEXP Rf_lang3(SEXP, SEXP, SEXP);
SEXP Rf_eval(SEXP, SEXP);
SEXP Rf_findVarInFrame(SEXP, SEXP);
void Rf_defineVar(SEXP, SEXP, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_GetOption1(SEXP);
SEXP Rf_ScalarInteger(int);
SEXP Rf_ScalarReal(double);
SEXP Rf_ScalarLogical(int);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
int Rf_inherits(SEXP, const char *);
int Rf_isNull(SEXP);
int Rf_isEnvironment(SEXP);
void Rf_error(const char *, ...) __attribute__((noreturn));
void Rf_warning(const char *, ...);
int LENGTH(SEXP);
R_xlen_t XLENGTH(SEXP);
int TYPEOF(SEXP);
int *INTEGER(SEXP);
int *LOGICAL(SEXP);
double *REAL(SEXP);
unsigned char *RAW(SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
const char *CHAR(SEXP);
void *R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
SEXP R_MakeExternalPtr(void *, SEXP, SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
char *R_alloc(size_t, int);
#endif

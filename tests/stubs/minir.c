/* minir.c -- a few hundred lines of "R" for tests: just enough of R's C API (the declarations in Rinternals.h of this
 * directory) to LINK r_shim/bigsnpr_shim.c into a shared object and to CALL its .Call entry points from the test-suite
 * with real vectors, environments and external pointers.  TEST INFRASTRUCTURE ONLY: it is not R, it never ships, and the
 * product never sees it.  Semantics implemented: typed vectors with attributes, environments as name -> value lists,
 * evaluation of the single call form the shim uses (`$`(env, "name")), Rf_error as a longjmp back to minir_call().
 * No garbage collection (tests are short): PROTECT / UNPROTECT are counters. */
#include <math.h>
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "R_ext/Rdynload.h"
#include "Rinternals.h"

#define CHARSXP 9
#define SYMSXP 1
#define ENVSXP 4
#define LANGSXP 6

struct attr {
  SEXP name, value;
  struct attr *next;
};
struct binding {
  char *name;
  SEXP value;
  struct binding *next;
};
struct SEXPREC {
  int type;
  R_xlen_t len;
  void *data;           /* vector payload / char* / external pointer address / SEXP[3] for calls */
  struct attr *attrs;
  struct binding *vars; /* environments */
  int nrow, ncol;
};

static struct SEXPREC nil_rec = {NILSXP, 0, NULL, NULL, NULL, 0, 0}, glob_rec = {ENVSXP, 0, NULL, NULL, NULL, 0, 0},
                      unbound_rec = {SYMSXP, 0, NULL, NULL, NULL, 0, 0};
SEXP R_NilValue = &nil_rec, R_GlobalEnv = &glob_rec, R_UnboundValue = &unbound_rec, R_NamesSymbol, R_ClassSymbol;
double R_NaReal;
int R_NaInt = (int)0x80000000;

static jmp_buf g_jmp;
static int g_jmp_armed = 0, g_protect = 0, g_warnings = 0;
static char g_error[1024], g_warning[1024];
static struct binding *g_options = NULL;
static const R_CallMethodDef *g_routines = NULL;

static SEXP mk(int type, R_xlen_t len, size_t elt) {
  SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
  s->type = type;
  s->len = len;
  s->data = (elt && len > 0) ? calloc((size_t)len, elt) : NULL;
  return s;
}

__attribute__((constructor)) static void minir_init(void) {
  union { unsigned long long u; double d; } na = {0x7FF00000000007A2ull}; /* R's NA_real_ payload 1954 */
  R_NaReal = na.d;
  R_NamesSymbol = Rf_install("names");
  R_ClassSymbol = Rf_install("class");
}

int R_IsNaN(double x) { return isnan(x); }
SEXP Rf_protect(SEXP s) { g_protect++; return s; }
void Rf_unprotect(int n) { g_protect -= n; }

SEXP Rf_allocVector(unsigned int type, R_xlen_t n) {
  size_t elt = type == REALSXP ? 8 : (type == INTSXP || type == LGLSXP) ? 4 : type == RAWSXP ? 1 : sizeof(SEXP);
  SEXP s = mk((int)type, n, elt);
  if (type == VECSXP || type == STRSXP)
    for (R_xlen_t i = 0; i < n; i++) ((SEXP *)s->data)[i] = R_NilValue;
  return s;
}
SEXP Rf_allocMatrix(unsigned int type, int nr, int nc) {
  SEXP s = Rf_allocVector(type, (R_xlen_t)nr * nc);
  s->nrow = nr;
  s->ncol = nc;
  return s;
}
SEXP Rf_mkChar(const char *c) {
  SEXP s = mk(CHARSXP, (R_xlen_t)strlen(c), 0);
  s->data = strdup(c);
  return s;
}
SEXP Rf_install(const char *c) {
  SEXP s = mk(SYMSXP, (R_xlen_t)strlen(c), 0);
  s->data = strdup(c);
  return s;
}
SEXP Rf_mkString(const char *c) {
  SEXP s = Rf_allocVector(STRSXP, 1);
  ((SEXP *)s->data)[0] = Rf_mkChar(c);
  return s;
}
const char *CHAR(SEXP s) { return (const char *)s->data; }
SEXP STRING_ELT(SEXP s, R_xlen_t i) { return ((SEXP *)s->data)[i]; }
void SET_STRING_ELT(SEXP s, R_xlen_t i, SEXP v) { ((SEXP *)s->data)[i] = v; }
SEXP VECTOR_ELT(SEXP s, R_xlen_t i) { return ((SEXP *)s->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP s, R_xlen_t i, SEXP v) { return ((SEXP *)s->data)[i] = v; }
int LENGTH(SEXP s) { return (int)s->len; }
R_xlen_t XLENGTH(SEXP s) { return s->len; }
int TYPEOF(SEXP s) { return s->type; }
int *INTEGER(SEXP s) { return (int *)s->data; }
int *LOGICAL(SEXP s) { return (int *)s->data; }
double *REAL(SEXP s) { return (double *)s->data; }
unsigned char *RAW(SEXP s) { return (unsigned char *)s->data; }
int Rf_nrows(SEXP s) { return s->nrow ? s->nrow : (int)s->len; }
int Rf_ncols(SEXP s) { return s->nrow ? s->ncol : 1; }
int Rf_isNull(SEXP s) { return s == R_NilValue; }
int Rf_isEnvironment(SEXP s) { return s->type == ENVSXP; }
SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); INTEGER(s)[0] = v; return s; }
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = v; return s; }
SEXP Rf_ScalarLogical(int v) { SEXP s = Rf_allocVector(LGLSXP, 1); LOGICAL(s)[0] = v; return s; }

int Rf_asInteger(SEXP s) {
  if (s->len < 1) return R_NaInt;
  if (s->type == INTSXP || s->type == LGLSXP) return INTEGER(s)[0];
  if (s->type == REALSXP) return isnan(REAL(s)[0]) ? R_NaInt : (int)REAL(s)[0];
  return R_NaInt;
}
int Rf_asLogical(SEXP s) { int v = Rf_asInteger(s); return v == R_NaInt ? R_NaInt : v != 0; }
double Rf_asReal(SEXP s) {
  if (s->len < 1) return R_NaReal;
  if (s->type == REALSXP) return REAL(s)[0];
  if (s->type == INTSXP || s->type == LGLSXP) return INTEGER(s)[0] == R_NaInt ? R_NaReal : (double)INTEGER(s)[0];
  return R_NaReal;
}
SEXP Rf_coerceVector(SEXP s, unsigned int type) {
  if ((unsigned)s->type == type) return s;
  SEXP r = Rf_allocVector(type, s->len);
  for (R_xlen_t i = 0; i < s->len; i++) {
    if (type == REALSXP && s->type == INTSXP) REAL(r)[i] = INTEGER(s)[i] == R_NaInt ? R_NaReal : INTEGER(s)[i];
    else if (type == INTSXP && s->type == REALSXP) INTEGER(r)[i] = isnan(REAL(s)[i]) ? R_NaInt : (int)REAL(s)[i];
    else Rf_error("minir: unsupported coercion %d -> %u", s->type, type);
  }
  return r;
}

SEXP Rf_setAttrib(SEXP s, SEXP name, SEXP val) {
  for (struct attr *a = s->attrs; a; a = a->next)
    if (!strcmp(CHAR(a->name), CHAR(name))) { a->value = val; return val; }
  struct attr *a = (struct attr *)calloc(1, sizeof *a);
  a->name = name; a->value = val; a->next = s->attrs; s->attrs = a;
  return val;
}
SEXP Rf_getAttrib(SEXP s, SEXP name) {
  for (struct attr *a = s->attrs; a; a = a->next)
    if (!strcmp(CHAR(a->name), CHAR(name))) return a->value;
  return R_NilValue;
}
int Rf_inherits(SEXP s, const char *cls) {
  SEXP k = Rf_getAttrib(s, R_ClassSymbol);
  if (k == R_NilValue) return 0;
  for (R_xlen_t i = 0; i < k->len; i++)
    if (!strcmp(CHAR(STRING_ELT(k, i)), cls)) return 1;
  return 0;
}
SEXP Rf_mkNamed(unsigned int type, const char **names) {
  int n = 0;
  while (names[n][0]) n++;
  SEXP s = Rf_allocVector(type, n), nm = Rf_allocVector(STRSXP, n);
  for (int i = 0; i < n; i++) SET_STRING_ELT(nm, i, Rf_mkChar(names[i]));
  Rf_setAttrib(s, R_NamesSymbol, nm);
  return s;
}

/* environments */
SEXP Rf_findVarInFrame(SEXP env, SEXP sym) {
  for (struct binding *b = env->vars; b; b = b->next)
    if (!strcmp(b->name, CHAR(sym))) return b->value;
  return R_UnboundValue;
}
void Rf_defineVar(SEXP sym, SEXP val, SEXP env) {
  for (struct binding *b = env->vars; b; b = b->next)
    if (!strcmp(b->name, CHAR(sym))) { b->value = val; return; }
  struct binding *b = (struct binding *)calloc(1, sizeof *b);
  b->name = strdup(CHAR(sym)); b->value = val; b->next = env->vars; env->vars = b;
}
SEXP Rf_lang1(SEXP f) { SEXP s = mk(LANGSXP, 1, sizeof(SEXP)); ((SEXP *)s->data)[0] = f; return s; }
SEXP Rf_lang2(SEXP f, SEXP a) { SEXP s = mk(LANGSXP, 2, sizeof(SEXP)); ((SEXP *)s->data)[0] = f; ((SEXP *)s->data)[1] = a; return s; }
SEXP Rf_lang3(SEXP f, SEXP a, SEXP b) {
  SEXP s = mk(LANGSXP, 3, sizeof(SEXP));
  ((SEXP *)s->data)[0] = f; ((SEXP *)s->data)[1] = a; ((SEXP *)s->data)[2] = b;
  return s;
}
/* the one call form the shim evaluates: `$`(env, "name") */
SEXP Rf_eval(SEXP call, SEXP rho) {
  (void)rho;
  if (call->type != LANGSXP) return call;
  SEXP *el = (SEXP *)call->data;
  if (call->len == 3 && !strcmp(CHAR(el[0]), "$") && el[1]->type == ENVSXP) {
    SEXP v = Rf_findVarInFrame(el[1], Rf_install(CHAR(STRING_ELT(el[2], 0))));
    if (v == R_UnboundValue) Rf_error("minir: object has no field '%s'", CHAR(STRING_ELT(el[2], 0)));
    return v;
  }
  Rf_error("minir: cannot evaluate this call");
}
SEXP Rf_GetOption1(SEXP sym) {
  for (struct binding *b = g_options; b; b = b->next)
    if (!strcmp(b->name, CHAR(sym))) return b->value;
  return R_NilValue;
}

void *R_ExternalPtrAddr(SEXP s) { return s->data; }
void R_ClearExternalPtr(SEXP s) { s->data = NULL; }
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot) { (void)tag; (void)prot; SEXP s = mk(EXTPTRSXP, 0, 0); s->data = p; return s; }
static R_CFinalizer_t g_fin[256];
static SEXP g_fin_obj[256];
static int g_nfin = 0;
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t f, Rboolean onexit) { (void)onexit; if (g_nfin < 256) { g_fin[g_nfin] = f; g_fin_obj[g_nfin++] = s; } }
char *R_alloc(size_t n, int size) { return (char *)calloc(n ? n : 1, (size_t)size); }

void Rf_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof g_error, fmt, ap);
  va_end(ap);
  if (g_jmp_armed) longjmp(g_jmp, 1);
  fprintf(stderr, "minir: uncaught error: %s\n", g_error);
  abort();
}
void Rf_warning(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_warning, sizeof g_warning, fmt, ap);
  va_end(ap);
  g_warnings++;
}
int R_registerRoutines(DllInfo *dll, const void *c, const R_CallMethodDef *call, const void *f, const void *e) {
  (void)dll; (void)c; (void)f; (void)e;
  g_routines = call;
  return 1;
}
Rboolean R_useDynamicSymbols(DllInfo *dll, Rboolean v) { (void)dll; return v; }

/* ---- test-side helpers (called through ctypes) ---------------------------------------------------------------- */
SEXP minir_int_vec(const int *v, int n) { SEXP s = Rf_allocVector(INTSXP, n); if (n) memcpy(s->data, v, (size_t)n * 4); return s; }
SEXP minir_real_vec(const double *v, int n) { SEXP s = Rf_allocVector(REALSXP, n); if (n) memcpy(s->data, v, (size_t)n * 8); return s; }
SEXP minir_real_mat(const double *v, int nr, int nc) { SEXP s = Rf_allocMatrix(REALSXP, nr, nc); if (nr > 0 && nc > 0) memcpy(s->data, v, (size_t)nr * nc * 8); return s; }
SEXP minir_lgl(int v) { return Rf_ScalarLogical(v); }
SEXP minir_str(const char *c) { return Rf_mkString(c); }
SEXP minir_env_new(void) { return mk(ENVSXP, 0, 0); }
void minir_env_set(SEXP env, const char *name, SEXP v) { Rf_defineVar(Rf_install(name), v, env); }
SEXP minir_env_get(SEXP env, const char *name) { return Rf_findVarInFrame(env, Rf_install(name)); }
SEXP minir_nil(void) { return R_NilValue; }
void minir_set_option(const char *name, SEXP v) {
  struct binding *b = (struct binding *)calloc(1, sizeof *b);
  b->name = strdup(name); b->value = v; b->next = g_options; g_options = b;
}
int minir_type(SEXP s) { return s->type; }
long minir_len(SEXP s) { return (long)s->len; }
void *minir_data(SEXP s) { return s->data; }
SEXP minir_list_get(SEXP s, int i) { return VECTOR_ELT(s, i); }
SEXP minir_list_by_name(SEXP s, const char *name) {
  SEXP nm = Rf_getAttrib(s, R_NamesSymbol);
  for (R_xlen_t i = 0; nm != R_NilValue && i < nm->len; i++)
    if (!strcmp(CHAR(STRING_ELT(nm, i)), name)) return VECTOR_ELT(s, i);
  return R_NilValue;
}
int minir_nrow(SEXP s) { return s->nrow; }
int minir_ncol(SEXP s) { return s->ncol; }
const char *minir_last_error(void) { return g_error; }
const char *minir_last_warning(void) { return g_warning; }
int minir_warning_count(void) { return g_warnings; }
int minir_protect_depth(void) { return g_protect; }
void minir_run_finalizers(void) {
  for (int i = 0; i < g_nfin; i++) g_fin[i](g_fin_obj[i]);
  g_nfin = 0;
}
int minir_routine_count(void) {
  int n = 0;
  while (g_routines && g_routines[n].name) n++;
  return n;
}
const char *minir_routine_name(int i) { return g_routines[i].name; }
int minir_routine_nargs(int i) { return g_routines[i].numArgs; }

typedef SEXP (*f1)(SEXP);
/* .Call(name, args...): looks the routine up in the registered table (like R does with .registration = TRUE), checks the
 * arity, calls it; an Rf_error inside comes back as NULL with the message in minir_last_error() */
SEXP minir_dot_call(const char *name, int nargs, SEXP *a) {
  const R_CallMethodDef *volatile r = g_routines;
  for (; r && r->name; r++)
    if (!strcmp(r->name, name)) break;
  g_error[0] = 0;
  if (!r || !r->name) { snprintf(g_error, sizeof g_error, "minir: no registered routine '%s'", name); return NULL; }
  if (r->numArgs != nargs) { snprintf(g_error, sizeof g_error, "minir: %s takes %d arguments, got %d", name, r->numArgs, nargs); return NULL; }
  g_jmp_armed = 1;
  if (setjmp(g_jmp)) { g_jmp_armed = 0; return NULL; }
  SEXP volatile res = NULL;
  void *volatile f = (void *)r->fun;
  switch (nargs) {
    case 3: res = ((SEXP(*)(SEXP, SEXP, SEXP))f)(a[0], a[1], a[2]); break;
    case 4: res = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP))f)(a[0], a[1], a[2], a[3]); break;
    case 5: res = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP))f)(a[0], a[1], a[2], a[3], a[4]); break;
    case 6: res = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(a[0], a[1], a[2], a[3], a[4], a[5]); break;
    case 7: res = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6]); break;
    case 8: res = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); break;
    case 12:
      res = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6],
                                                                                                a[7], a[8], a[9], a[10], a[11]);
      break;
    default: snprintf(g_error, sizeof g_error, "minir: arity %d not wired", nargs); res = NULL;
  }
  g_jmp_armed = 0;
  return res;
}

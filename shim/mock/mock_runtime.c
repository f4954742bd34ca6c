/* mock_runtime.c -- a few hundred bytes of "R" behind shim/mock/Rinternals.h plus the harness tests/test_shim.py drives
 * through ctypes: build argument objects, call a registered .Call routine by name (Rf_error is caught with setjmp and
 * reported), read results.  Objects are never freed (short-lived test process). */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "Rinternals.h"
#include "R_ext/Rdynload.h"

struct SEXPREC {
  int type;
  R_xlen_t len;
  int nrow, ncol;        /* dim attribute (0 = none) */
  void* data;
  SEXP names;
};
static struct SEXPREC nil_rec = {NILSXP, 0, 0, 0, NULL, NULL}, names_rec = {NILSXP, 0, 0, 0, NULL, NULL};
SEXP R_NilValue = &nil_rec, R_NamesSymbol = &names_rec;

static jmp_buf g_jmp;
static int g_jmp_set = 0;
static char g_err[1024] = "";
static const R_CallMethodDef* g_table = NULL;

SEXP Rf_allocVector(int type, R_xlen_t n) {
  SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
  size_t el = type == REALSXP ? sizeof(double) : (type == INTSXP || type == LGLSXP) ? sizeof(int) : sizeof(void*);
  s->type = type; s->len = n; s->data = calloc((size_t)(n > 0 ? n : 1), el); s->names = R_NilValue;
  return s;
}
SEXP Rf_allocMatrix(int type, int nrow, int ncol) {
  SEXP s = Rf_allocVector(type, (R_xlen_t)nrow * ncol);
  s->nrow = nrow; s->ncol = ncol;
  return s;
}
double* REAL(SEXP x) { if (x->type != REALSXP) Rf_error("REAL() on a non-double object"); return (double*)x->data; }
int* INTEGER(SEXP x) { if (x->type != INTSXP && x->type != LGLSXP) Rf_error("INTEGER() on a non-integer object"); return (int*)x->data; }
int* LOGICAL(SEXP x) { return INTEGER(x); }
R_xlen_t XLENGTH(SEXP x) { return x->len; }
int Rf_nrows(SEXP x) { return x->nrow ? x->nrow : (int)x->len; }
int Rf_ncols(SEXP x) { return x->nrow ? x->ncol : 1; }
int Rf_isNull(SEXP x) { return x == R_NilValue || x->type == NILSXP; }
SEXP STRING_ELT(SEXP x, R_xlen_t i) { return ((SEXP*)x->data)[i]; }
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP*)x->data)[i] = v; }
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) { return ((SEXP*)x->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP*)x->data)[i] = v; return v; }
const char* CHAR(SEXP x) { return (const char*)x->data; }
SEXP Rf_mkChar(const char* s) {
  SEXP c = (SEXP)calloc(1, sizeof(struct SEXPREC));
  c->type = CHARSXP; c->len = (R_xlen_t)strlen(s); c->data = strdup(s); c->names = R_NilValue;
  return c;
}
SEXP Rf_mkString(const char* s) { SEXP v = Rf_allocVector(STRSXP, 1); SET_STRING_ELT(v, 0, Rf_mkChar(s)); return v; }
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = v; return s; }
SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); INTEGER(s)[0] = v; return s; }
SEXP Rf_ScalarLogical(int v) { SEXP s = Rf_allocVector(LGLSXP, 1); INTEGER(s)[0] = v != 0; return s; }
int Rf_asInteger(SEXP x) { return x->type == REALSXP ? (int)REAL(x)[0] : INTEGER(x)[0]; }
int Rf_asLogical(SEXP x) { return x->type == REALSXP ? REAL(x)[0] != 0 : INTEGER(x)[0] != 0; }
double Rf_asReal(SEXP x) { return x->type == REALSXP ? REAL(x)[0] : (double)INTEGER(x)[0]; }
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP val) { if (name == R_NamesSymbol) x->names = val; return val; }
SEXP Rf_getAttrib(SEXP x, SEXP name) { return name == R_NamesSymbol ? x->names : R_NilValue; }
SEXP Rf_protect(SEXP x) { return x; }
void Rf_unprotect(int n) { (void)n; }
char* R_alloc(size_t n, int size) { return (char*)calloc(n ? n : 1, (size_t)size); }
void Rf_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  if (g_jmp_set) longjmp(g_jmp, 1);
  fprintf(stderr, "Rf_error outside a .Call: %s\n", g_err);
  abort();
}
int R_registerRoutines(DllInfo* info, const void* c, const R_CallMethodDef* call, const void* f, const void* e) {
  (void)info; (void)c; (void)f; (void)e;
  g_table = call;
  return 1;
}
int R_useDynamicSymbols(DllInfo* info, int value) { (void)info; (void)value; return 0; }

/* ---- harness (ctypes) ----------------------------------------------------------------------------------------- */
void R_init_MagmaClustR(DllInfo* dll);
void R_unload_MagmaClustR(DllInfo* dll);
int mock_init(void) { R_init_MagmaClustR(NULL); return g_table != NULL; }
void mock_unload(void) { R_unload_MagmaClustR(NULL); }
int mock_n_routines(void) { int n = 0; while (g_table && g_table[n].name) ++n; return n; }
const char* mock_routine_name(int i) { return g_table[i].name; }
int mock_routine_nargs(int i) { return g_table[i].numArgs; }
const char* mock_last_error(void) { return g_err; }
SEXP mock_nil(void) { return R_NilValue; }
SEXP mock_real(const double* v, R_xlen_t n, int nrow, int ncol) {
  SEXP s = nrow > 0 ? Rf_allocMatrix(REALSXP, nrow, ncol) : Rf_allocVector(REALSXP, n);
  memcpy(s->data, v, sizeof(double) * (size_t)n);
  return s;
}
SEXP mock_int(const int* v, R_xlen_t n) { SEXP s = Rf_allocVector(INTSXP, n); memcpy(s->data, v, sizeof(int) * (size_t)n); return s; }
SEXP mock_lgl(int v) { return Rf_ScalarLogical(v); }
SEXP mock_str(const char* s) { return Rf_mkString(s); }
/* .Call(name, args...): returns the result or NULL after an Rf_error (message in mock_last_error) */
SEXP mock_call(const char* name, int nargs, SEXP* a) {
  for (int i = 0; g_table && g_table[i].name; ++i) {
    if (strcmp(g_table[i].name, name) != 0) continue;
    if (g_table[i].numArgs != nargs) { snprintf(g_err, sizeof(g_err), "%s takes %d arguments", name, g_table[i].numArgs); return NULL; }
    g_err[0] = 0;
    g_jmp_set = 1;
    if (setjmp(g_jmp)) { g_jmp_set = 0; return NULL; }
    SEXP r = NULL;
    void (*f)(void) = (void (*)(void))g_table[i].fun;   /* generic function pointer: cast through void(*)(void) */
    typedef SEXP (*F1)(SEXP); typedef SEXP (*F2)(SEXP, SEXP); typedef SEXP (*F3)(SEXP, SEXP, SEXP);
    typedef SEXP (*F6)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP); typedef SEXP (*F7)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
    typedef SEXP (*F9)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
    switch (nargs) {
      case 1: r = ((F1)f)(a[0]); break;
      case 2: r = ((F2)f)(a[0], a[1]); break;
      case 3: r = ((F3)f)(a[0], a[1], a[2]); break;
      case 6: r = ((F6)f)(a[0], a[1], a[2], a[3], a[4], a[5]); break;
      case 7: r = ((F7)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6]); break;
      case 9: r = ((F9)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]); break;
      default: snprintf(g_err, sizeof(g_err), "mock_call: arity %d not wired", nargs); g_jmp_set = 0; return NULL;
    }
    g_jmp_set = 0;
    return r;
  }
  snprintf(g_err, sizeof(g_err), "no routine %s", name);
  return NULL;
}
int mock_type(SEXP x) { return x->type; }
R_xlen_t mock_length(SEXP x) { return x->len; }
int mock_nrow(SEXP x) { return x->nrow; }
int mock_ncol(SEXP x) { return x->ncol; }
const double* mock_real_ptr(SEXP x) { return (const double*)x->data; }
SEXP mock_list_get(SEXP x, const char* name) {
  if (x->type != VECSXP || Rf_isNull(x->names)) return NULL;
  for (R_xlen_t i = 0; i < x->len; ++i)
    if (strcmp(CHAR(STRING_ELT(x->names, i)), name) == 0) return VECTOR_ELT(x, i);
  return NULL;
}

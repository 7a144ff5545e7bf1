#define _GNU_SOURCE
/* tests/r_stub/fake_r.c — TEST INFRASTRUCTURE: implementation of the stand-in R C API declared in Rinternals.h, plus the
 * psy_rstub_* helpers tests/test_r_shim.py uses to build arguments, call registered routines and read results. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "Rinternals.h"
#include "R_ext/Rdynload.h"

struct psy_rstub_sexp {
  int type, len;
  void* data;           /* double[] / int[] / SEXP[] / char[] / external pointer */
  SEXP names, dim, prot;
  R_CFinalizer_t fin;
};

static struct psy_rstub_sexp nil_obj = {NILSXP, 0, NULL, NULL, NULL, NULL, NULL}, names_sym = {NILSXP, 0, NULL, NULL, NULL, NULL, NULL},
                             dim_sym = {NILSXP, 0, NULL, NULL, NULL, NULL, NULL};
SEXP R_NilValue = &nil_obj, R_NamesSymbol = &names_sym, R_DimSymbol = &dim_sym;

static jmp_buf* cur_jmp = NULL;
static char err_msg[4096], warn_msg[4096];
static int n_warnings = 0;

static SEXP new_sexp(int type, int len, size_t elem) {
  SEXP s = (SEXP)calloc(1, sizeof *s);
  s->type = type; s->len = len;
  s->data = calloc((size_t)(len > 0 ? len : 1), elem);
  s->names = s->dim = s->prot = R_NilValue;
  return s;
}

int TYPEOF(SEXP x) { return x->type; }
int LENGTH(SEXP x) { return x->len; }
double* REAL(SEXP x) { if (x->type != REALSXP) Rf_error("REAL() can only be applied to a 'numeric', not a type %d", x->type); return (double*)x->data; }
int* INTEGER(SEXP x) { if (x->type != INTSXP && x->type != LGLSXP) Rf_error("INTEGER() can only be applied to a 'integer', not a type %d", x->type); return (int*)x->data; }
int* LOGICAL(SEXP x) { if (x->type != LGLSXP) Rf_error("LOGICAL() can only be applied to a 'logical', not a type %d", x->type); return (int*)x->data; }
SEXP STRING_ELT(SEXP x, ptrdiff_t i) { if (x->type != STRSXP || i < 0 || i >= x->len) Rf_error("STRING_ELT: bad access"); return ((SEXP*)x->data)[i]; }
void SET_STRING_ELT(SEXP x, ptrdiff_t i, SEXP v) { if (x->type != STRSXP || i < 0 || i >= x->len) Rf_error("SET_STRING_ELT: bad access"); ((SEXP*)x->data)[i] = v; }
SEXP VECTOR_ELT(SEXP x, ptrdiff_t i) { if (x->type != VECSXP || i < 0 || i >= x->len) Rf_error("VECTOR_ELT: bad access"); return ((SEXP*)x->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP x, ptrdiff_t i, SEXP v) { if (x->type != VECSXP || i < 0 || i >= x->len) Rf_error("SET_VECTOR_ELT: bad access"); ((SEXP*)x->data)[i] = v; return v; }
const char* CHAR(SEXP x) { if (x->type != CHARSXP) Rf_error("CHAR() can only be applied to a 'CHARSXP'"); return (const char*)x->data; }

SEXP Rf_allocVector(unsigned type, ptrdiff_t n) {
  switch (type) {
    case REALSXP: return new_sexp(REALSXP, (int)n, sizeof(double));
    case INTSXP: case LGLSXP: return new_sexp((int)type, (int)n, sizeof(int));
    case STRSXP: { SEXP s = new_sexp(STRSXP, (int)n, sizeof(SEXP)); for (int i = 0; i < n; i++) ((SEXP*)s->data)[i] = Rf_mkChar(""); return s; }
    case VECSXP: { SEXP s = new_sexp(VECSXP, (int)n, sizeof(SEXP)); for (int i = 0; i < n; i++) ((SEXP*)s->data)[i] = R_NilValue; return s; }
    default: Rf_error("allocVector: type %u not supported by the stand-in", type);
  }
}
SEXP Rf_allocMatrix(unsigned type, int nrow, int ncol) {
  SEXP s = Rf_allocVector(type, (ptrdiff_t)nrow * ncol);
  s->dim = Rf_allocVector(INTSXP, 2);
  INTEGER(s->dim)[0] = nrow; INTEGER(s->dim)[1] = ncol;
  return s;
}
SEXP Rf_mkChar(const char* str) {
  SEXP s = new_sexp(CHARSXP, (int)strlen(str), 1);
  free(s->data);
  s->data = strdup(str);
  return s;
}
SEXP Rf_mkString(const char* str) { SEXP s = Rf_allocVector(STRSXP, 1); SET_STRING_ELT(s, 0, Rf_mkChar(str)); return s; }
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = v; return s; }
SEXP Rf_duplicate(SEXP x) {
  if (x == R_NilValue) return x;
  SEXP s;
  switch (x->type) {
    case REALSXP: s = new_sexp(REALSXP, x->len, sizeof(double)); memcpy(s->data, x->data, sizeof(double) * (size_t)x->len); break;
    case INTSXP: case LGLSXP: s = new_sexp(x->type, x->len, sizeof(int)); memcpy(s->data, x->data, sizeof(int) * (size_t)x->len); break;
    case STRSXP: s = new_sexp(STRSXP, x->len, sizeof(SEXP)); memcpy(s->data, x->data, sizeof(SEXP) * (size_t)x->len); break;  /* CHARSXPs are immutable */
    case VECSXP: s = new_sexp(VECSXP, x->len, sizeof(SEXP)); for (int i = 0; i < x->len; i++) ((SEXP*)s->data)[i] = Rf_duplicate(((SEXP*)x->data)[i]); break;
    case CHARSXP: return x;
    default: return x;  /* external pointers are reference objects */
  }
  s->names = Rf_duplicate(x->names);
  s->dim = Rf_duplicate(x->dim);
  return s;
}
SEXP Rf_getAttrib(SEXP x, SEXP name) { return name == R_NamesSymbol ? x->names : (name == R_DimSymbol ? x->dim : R_NilValue); }
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP v) { if (name == R_NamesSymbol) x->names = v; else if (name == R_DimSymbol) x->dim = v; return v; }
int Rf_nrows(SEXP x) { return x->dim != R_NilValue ? INTEGER(x->dim)[0] : x->len; }   /* like R: a plain vector is one column */
int Rf_ncols(SEXP x) { return x->dim != R_NilValue ? INTEGER(x->dim)[1] : 1; }
int Rf_asInteger(SEXP x) { return x->len < 1 ? 0 : (x->type == REALSXP ? (int)((double*)x->data)[0] : ((int*)x->data)[0]); }
int Rf_asLogical(SEXP x) { return x->len < 1 ? 0 : (x->type == REALSXP ? ((double*)x->data)[0] != 0.0 : ((int*)x->data)[0] != 0); }
double Rf_asReal(SEXP x) { return x->len < 1 ? 0.0 : (x->type == REALSXP ? ((double*)x->data)[0] : (double)((int*)x->data)[0]); }
void Rf_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(err_msg, sizeof err_msg, fmt, ap); va_end(ap);
  if (cur_jmp) longjmp(*cur_jmp, 1);
  fprintf(stderr, "R stand-in: error outside psy_rstub_call: %s\n", err_msg);
  abort();
}
void Rf_warning(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(warn_msg, sizeof warn_msg, fmt, ap); va_end(ap); n_warnings++; }
char* R_alloc(size_t n, int size) { return (char*)calloc(n > 0 ? n : 1, (size_t)size); }

SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) { (void)tag; SEXP s = new_sexp(EXTPTRSXP, 0, 1); free(s->data); s->data = p; s->prot = prot; return s; }
SEXP R_ExternalPtrProtected(SEXP s) { return s->prot; }
void R_SetExternalPtrProtected(SEXP s, SEXP p) { s->prot = p; }
void* R_ExternalPtrAddr(SEXP s) { return s->type == EXTPTRSXP ? s->data : NULL; }
void R_ClearExternalPtr(SEXP s) { s->data = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit) { (void)onexit; s->fin = fun; }

/* ---- routine registration ------------------------------------------------------------------------------------------ */
static const R_CallMethodDef* registered = NULL;
static int dynamic_symbols = 1;
int R_registerRoutines(DllInfo* info, const void* c, const R_CallMethodDef* call, const void* f, const void* ext) {
  (void)info; (void)c; (void)f; (void)ext; registered = call; return 1;
}
Rboolean R_useDynamicSymbols(DllInfo* info, Rboolean value) { (void)info; int old = dynamic_symbols; dynamic_symbols = value; return old; }

/* ---- helpers for the Python harness -------------------------------------------------------------------------------- */
int psy_rstub_n_registered(void) { int k = 0; if (registered) while (registered[k].name) k++; return k; }
const char* psy_rstub_registered_name(int k) { return registered[k].name; }
int psy_rstub_registered_arity(int k) { return registered[k].numArgs; }
int psy_rstub_dynamic_symbols(void) { return dynamic_symbols; }
const char* psy_rstub_error(void) { return err_msg; }
const char* psy_rstub_warning(void) { return warn_msg; }
int psy_rstub_warnings(void) { return n_warnings; }
void psy_rstub_run_finalizer(SEXP s) { if (s->fin) s->fin(s); }

/* .Call(name, ...): looks the routine up in the registered table, checks the arity like R does, and runs it with Rf_error
 * caught.  Returns NULL on error (message in psy_rstub_error()). */
typedef SEXP (*fn12)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP psy_rstub_call(const char* name, int nargs, SEXP* a) {
  err_msg[0] = 0;
  const R_CallMethodDef* volatile d = NULL;
  for (int k = 0; registered && registered[k].name; k++)
    if (strcmp(registered[k].name, name) == 0) d = &registered[k];
  if (!d) { snprintf(err_msg, sizeof err_msg, "C symbol name \"%s\" not in load table", name); return NULL; }
  if (d->numArgs != nargs) { snprintf(err_msg, sizeof err_msg, "Incorrect number of arguments (%d), expecting %d for '%s'", nargs, d->numArgs, name); return NULL; }
  if (nargs > 12) { snprintf(err_msg, sizeof err_msg, "stand-in supports at most 12 arguments"); return NULL; }
  SEXP v[12];
  for (int i = 0; i < 12; i++) v[i] = i < nargs ? a[i] : R_NilValue;
  jmp_buf jb;
  jmp_buf* prev = cur_jmp;
  cur_jmp = &jb;
  SEXP volatile out = NULL;
  if (setjmp(jb) == 0) {
    /* calling a function with more (pointer) arguments than it declares is fine under the SysV x86-64 / AAPCS64 ABIs */
    out = ((fn12)d->fun)(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11]);
  }
  cur_jmp = prev;
  return out;
}

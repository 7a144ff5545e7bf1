/* tests/r_stub/Rinternals.h — TEST INFRASTRUCTURE ONLY.
 *
 * A small FUNCTIONAL stand-in for the part of R's C API that r/psydiff_b200_shim.c uses, so that the .Call layer can be
 * compiled and executed by the test suite on a machine without R (tests/test_r_shim.py).  Same names, argument order and
 * observable behaviour as the documented API ("Writing R Extensions", section 5); memory is simply leaked (no GC),
 * PROTECT/UNPROTECT are no-ops, Rf_error unwinds by longjmp to psy_rstub_call.  Not R, and not shipped.
 */
#ifndef PSY_R_STUB_RINTERNALS_H
#define PSY_R_STUB_RINTERNALS_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct psy_rstub_sexp* SEXP;
typedef int Rboolean;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif

#define NILSXP 0
#define CHARSXP 9
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define EXTPTRSXP 22

extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;

int TYPEOF(SEXP x);
int LENGTH(SEXP x);
double* REAL(SEXP x);
int* INTEGER(SEXP x);
int* LOGICAL(SEXP x);
SEXP STRING_ELT(SEXP x, ptrdiff_t i);
void SET_STRING_ELT(SEXP x, ptrdiff_t i, SEXP v);
SEXP VECTOR_ELT(SEXP x, ptrdiff_t i);
SEXP SET_VECTOR_ELT(SEXP x, ptrdiff_t i, SEXP v);
const char* CHAR(SEXP x);

SEXP Rf_allocVector(unsigned type, ptrdiff_t n);
SEXP Rf_allocMatrix(unsigned type, int nrow, int ncol);
SEXP Rf_mkChar(const char* s);
SEXP Rf_mkString(const char* s);
SEXP Rf_ScalarReal(double v);
SEXP Rf_duplicate(SEXP x);
SEXP Rf_getAttrib(SEXP x, SEXP name);
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP v);
int Rf_nrows(SEXP x);
int Rf_ncols(SEXP x);
int Rf_asInteger(SEXP x);
int Rf_asLogical(SEXP x);
double Rf_asReal(SEXP x);
void Rf_error(const char* fmt, ...) __attribute__((noreturn));
void Rf_warning(const char* fmt, ...);
char* R_alloc(size_t n, int size);

#define PROTECT(x) (x)
#define UNPROTECT(n) ((void)(n))

typedef void (*R_CFinalizer_t)(SEXP);
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot);
void* R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);
SEXP R_ExternalPtrProtected(SEXP s);
void R_SetExternalPtrProtected(SEXP s, SEXP p);

#ifdef __cplusplus
}
#endif
#endif

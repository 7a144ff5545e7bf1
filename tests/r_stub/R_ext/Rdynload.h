/* tests/r_stub/R_ext/Rdynload.h — TEST INFRASTRUCTURE: stand-in for R's routine registration. */
#ifndef PSY_R_STUB_RDYNLOAD_H
#define PSY_R_STUB_RDYNLOAD_H
#include "../Rinternals.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef void* (*DL_FUNC)(void);
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct psy_rstub_dll DllInfo;
int R_registerRoutines(DllInfo* info, const void* c, const R_CallMethodDef* call, const void* f, const void* ext);
Rboolean R_useDynamicSymbols(DllInfo* info, Rboolean value);
#ifdef __cplusplus
}
#endif
#endif

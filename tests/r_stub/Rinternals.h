/* Minimal stand-in for R's C API: ONLY the declarations src/g3cuda_R.c uses, so that the shim can be syntax- and
 * type-checked in an image without R (tests/test_r_glue.py). Test infrastructure; never shipped or linked. */
#ifndef G3_R_STUB_RINTERNALS_H
#define G3_R_STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#define TRUE 1
#define FALSE 0
#define REALSXP 14
#define VECSXP 19
extern SEXP R_NilValue;
typedef void (*R_CFinalizer_t)(SEXP);
typedef void* (*DL_FUNC)(void);
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct _DllInfo DllInfo;
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
R_xlen_t XLENGTH(SEXP);
int* INTEGER(SEXP);
double* REAL(SEXP);
SEXP Rf_allocVector(unsigned int, R_xlen_t);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
Rboolean Rf_isNull(SEXP);
Rboolean Rf_isReal(SEXP);
int Rf_asLogical(SEXP);
void Rf_error(const char*, ...);
int R_registerRoutines(DllInfo*, const void*, const R_CallMethodDef*, const void*, const void*);
Rboolean R_useDynamicSymbols(DllInfo*, Rboolean);
#endif

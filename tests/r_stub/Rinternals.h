/* Minimal stand-in for R's <Rinternals.h>: ONLY the declarations r_overlay/src/zz_shim.c uses, with the prototypes of R 4.x
 * (R is not installed in this image).  tests/test_r_overlay.py compiles the shim against it so that a typo, a wrong arity or a
 * missing symbol in the shim fails a test instead of surfacing at an R maintainer's desk.  Test infrastructure only. */
#ifndef ZZ_R_STUB_RINTERNALS_H
#define ZZ_R_STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef unsigned int SEXPTYPE;
typedef ptrdiff_t R_xlen_t;
typedef unsigned char Rbyte;
typedef enum { FALSE = 0, TRUE } Rboolean;
#define INTSXP 13
#define REALSXP 14
#define VECSXP 19
#define RAWSXP 24
extern SEXP R_NilValue, R_DimSymbol, R_NamesSymbol;
double *REAL(SEXP x);
int *INTEGER(SEXP x);
int *LOGICAL(SEXP x);
Rbyte *RAW(SEXP x);
int LENGTH(SEXP x);
const char *CHAR(SEXP x);
SEXP STRING_ELT(SEXP x, R_xlen_t i);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n);
SEXP Rf_allocMatrix(SEXPTYPE type, int nrow, int ncol);
SEXP Rf_mkNamed(SEXPTYPE type, const char **names);
SEXP Rf_getAttrib(SEXP x, SEXP name);
SEXP Rf_ScalarReal(double v);
SEXP Rf_ScalarLogical(int v);
double Rf_asReal(SEXP x);
int Rf_asInteger(SEXP x);
int Rf_asLogical(SEXP x);
Rboolean Rf_isNull(SEXP x);
SEXP Rf_protect(SEXP x);
void Rf_unprotect(int n);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
void Rf_error(const char *fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
char *R_alloc(size_t n, int size);
typedef void (*R_CFinalizer_t)(SEXP);
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
void *R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);
#endif

/* Minimal stand-in for R's Rinternals.h: ONLY the declarations b200als_shim.c uses, with R's real signatures
 * (R-4.x, src/include/Rinternals.h).  It exists so the shim can be syntax- and type-checked in an image without R
 * (r-package/tests/check_shim.sh); it is never linked or shipped. */
#ifndef STUB_RINTERNALS_H
#define STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE } Rboolean;
typedef unsigned int SEXPTYPE;
#define INTSXP 13
#define REALSXP 14
extern SEXP R_NilValue, R_DimSymbol;
int* INTEGER(SEXP);
double* REAL(SEXP);
R_xlen_t XLENGTH(SEXP);
SEXP R_do_slot(SEXP obj, SEXP name);
SEXP Rf_install(const char*);
SEXP Rf_getAttrib(SEXP, SEXP);
Rboolean Rf_isNull(SEXP);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_ScalarReal(double);
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_lengthgets(SEXP, R_xlen_t);
void Rf_error(const char*, ...) __attribute__((noreturn));
char* R_alloc(size_t, int);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
typedef void (*R_CFinalizer_t)(SEXP);
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot);
void* R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);
#endif

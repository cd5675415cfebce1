#ifndef RINTERNALS_STUB_H
#define RINTERNALS_STUB_H
typedef struct SEXPREC *SEXP;
typedef enum { FALSE = 0, TRUE } Rboolean;
extern SEXP R_NilValue;
double *REAL(SEXP x);
int *INTEGER(SEXP x);
int asInteger(SEXP x);
double asReal(SEXP x);
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
void *R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
#endif

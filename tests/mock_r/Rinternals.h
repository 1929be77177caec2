/* see R.h in this directory */
#ifndef MOCK_RINTERNALS_H
#define MOCK_RINTERNALS_H
#include <stddef.h>
#include <stdint.h>

typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE = 1 } Rboolean;
#define NILSXP 0
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define VECSXP 19

typedef struct mock_sexp {
    int type;
    R_xlen_t length;
    int nrow, ncol;            /* 0 / 0 for a plain vector */
    void* data;                /* double[], int[] or struct mock_sexp*[] */
} *SEXP;

extern SEXP R_NilValue;
extern int mock_protect_balance;

SEXP Rf_allocVector(int type, R_xlen_t n);
SEXP Rf_allocMatrix(int type, int nrow, int ncol);
SEXP Rf_ScalarInteger(int v);
SEXP Rf_ScalarLogical(int v);
SEXP Rf_ScalarReal(double v);
int Rf_nrows(SEXP x);
int Rf_ncols(SEXP x);
int Rf_asLogical(SEXP x);
int Rf_asInteger(SEXP x);
double Rf_asReal(SEXP x);
int Rf_isNull(SEXP x);
R_xlen_t Rf_xlength(SEXP x);
#define LENGTH(x) ((int)Rf_xlength(x))
double* REAL(SEXP x);
int* INTEGER(SEXP x);
int* LOGICAL(SEXP x);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP mock_protect(SEXP x);
void mock_unprotect(int n);
#define PROTECT(x) mock_protect(x)
#define UNPROTECT(n) mock_unprotect(n)
char* R_alloc(size_t n, int size);
void Rf_error(const char* fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
#endif

/* Mock R runtime (see R.h in this directory). */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "Rinternals.h"
#include "R_ext/Rdynload.h"

static struct mock_sexp nil_value = {NILSXP, 0, 0, 0, NULL};
SEXP R_NilValue = &nil_value;
int mock_protect_balance = 0;
jmp_buf mock_error_jmp;
int mock_error_armed = 0;
char mock_error_message[1024];

SEXP Rf_allocVector(int type, R_xlen_t n) {
    SEXP s = (SEXP)calloc(1, sizeof(*s));
    size_t el = type == REALSXP ? sizeof(double) : type == VECSXP ? sizeof(SEXP) : sizeof(int);
    s->type = type;
    s->length = n;
    s->data = calloc(n > 0 ? (size_t)n : 1, el);
    if (type == VECSXP) for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->data)[i] = R_NilValue;
    return s;
}
SEXP Rf_allocMatrix(int type, int nrow, int ncol) {
    SEXP s = Rf_allocVector(type, (R_xlen_t)nrow * ncol);
    s->nrow = nrow;
    s->ncol = ncol;
    return s;
}
SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); ((int*)s->data)[0] = v; return s; }
SEXP Rf_ScalarLogical(int v) { SEXP s = Rf_allocVector(LGLSXP, 1); ((int*)s->data)[0] = v; return s; }
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); ((double*)s->data)[0] = v; return s; }
int Rf_nrows(SEXP x) { return x->nrow ? x->nrow : (int)x->length; }
int Rf_ncols(SEXP x) { return x->nrow ? x->ncol : 1; }
int Rf_asLogical(SEXP x) { return x->type == REALSXP ? ((double*)x->data)[0] != 0.0 : ((int*)x->data)[0] != 0; }
int Rf_asInteger(SEXP x) { return x->type == REALSXP ? (int)((double*)x->data)[0] : ((int*)x->data)[0]; }
double Rf_asReal(SEXP x) { return x->type == REALSXP ? ((double*)x->data)[0] : (double)((int*)x->data)[0]; }
int Rf_isNull(SEXP x) { return x == R_NilValue || x->type == NILSXP; }
R_xlen_t Rf_xlength(SEXP x) { return x->length; }
static void need(SEXP x, int type, const char* what) {
    if (x->type != type) { fprintf(stderr, "mock R: %s() on an object of type %d\n", what, x->type); abort(); }
}
double* REAL(SEXP x) { need(x, REALSXP, "REAL"); return (double*)x->data; }
int* INTEGER(SEXP x) { need(x, INTSXP, "INTEGER"); return (int*)x->data; }
int* LOGICAL(SEXP x) { need(x, LGLSXP, "LOGICAL"); return (int*)x->data; }
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) { need(x, VECSXP, "SET_VECTOR_ELT"); ((SEXP*)x->data)[i] = v; return v; }
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) { need(x, VECSXP, "VECTOR_ELT"); return ((SEXP*)x->data)[i]; }
SEXP mock_protect(SEXP x) { ++mock_protect_balance; return x; }
void mock_unprotect(int n) { mock_protect_balance -= n; }
char* R_alloc(size_t n, int size) { return (char*)calloc(n > 0 ? n : 1, (size_t)size); }
void Rf_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(mock_error_message, sizeof(mock_error_message), fmt, ap);
    va_end(ap);
    if (mock_error_armed) longjmp(mock_error_jmp, 1);
    fprintf(stderr, "Error: %s\n", mock_error_message);
    exit(3);
}
int R_registerRoutines(DllInfo* info, const void* c, const R_CallMethodDef* call_methods, const void* f, const void* e) {
    (void)c; (void)f; (void)e;
    info->call_methods = call_methods;
    return 1;
}
Rboolean R_useDynamicSymbols(DllInfo* info, Rboolean value) { info->dynamic_symbols = value; return TRUE; }

/* Driver of r-shim/src/shim.c under the mock R runtime (tests/test_shim.py).
 *   shim_driver list                 registered .Call routines (name nargs), CPU only
 *   shim_driver core in.bin out.bin  _RUVIIIPRPS_ruvIIICore on the problem in in.bin, results to out.bin (needs a GPU)
 *   shim_driver prims out.bin        the five matrix primitives of RcppExports.R:4-26 on a fixed small case (needs a GPU)
 * Routines are looked up BY NAME in the table R_init_RUVIIIPRPS registers, as R's .Call does. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "Rinternals.h"
#include "R_ext/Rdynload.h"

void R_init_RUVIIIPRPS(DllInfo* dll);
void R_unload_RUVIIIPRPS(DllInfo* dll);

static DllInfo dll;
static DL_FUNC lookup(const char* name, int nargs) {
    for (const R_CallMethodDef* m = dll.call_methods; m && m->name; ++m)
        if (strcmp(m->name, name) == 0) {
            if (m->numArgs != nargs) { fprintf(stderr, "%s registered with %d args, called with %d\n", name, m->numArgs, nargs); exit(2); }
            return m->fun;
        }
    fprintf(stderr, "%s is not registered\n", name);
    exit(2);
}
static void rd(void* p, size_t sz, size_t n, FILE* f) { if (fread(p, sz, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }
static void wr(const void* p, size_t sz, size_t n, FILE* f) { if (fwrite(p, sz, n, f) != n) { fprintf(stderr, "short write\n"); exit(2); } }
static void write_real(SEXP x, FILE* f) {
    int64_t len = Rf_isNull(x) ? 0 : (int64_t)Rf_xlength(x);
    wr(&len, 8, 1, f);
    if (len) wr(REAL(x), 8, (size_t)len, f);
}

int main(int argc, char** argv) {
    R_init_RUVIIIPRPS(&dll);
    if (argc >= 2 && strcmp(argv[1], "list") == 0) {
        for (const R_CallMethodDef* m = dll.call_methods; m && m->name; ++m) printf("%s %d\n", m->name, m->numArgs);
        printf("dynamic_symbols %d\n", dll.dynamic_symbols);
        return 0;
    }
    if (argc >= 4 && strcmp(argv[1], "core") == 0) {
        FILE* f = fopen(argv[2], "rb");
        if (!f) { perror(argv[2]); return 2; }
        int32_t hdr[8];   /* n, m1, n_ps, nk, mode, applyLog, returnInfo, n_gid */
        rd(hdr, 4, 8, f);
        double pseudo;
        rd(&pseudo, 8, 1, f);
        const int n = hdr[0], m1 = hdr[1], n_ps = hdr[2], nk = hdr[3], n_gid = hdr[7];
        SEXP assay = Rf_allocMatrix(REALSXP, n, m1), tRep = Rf_allocMatrix(REALSXP, n, n_ps);
        SEXP gid = Rf_allocVector(INTSXP, n_gid), ctl = Rf_allocVector(LGLSXP, n), k = Rf_allocVector(INTSXP, nk);
        rd(REAL(assay), 8, (size_t)n * m1, f);
        rd(REAL(tRep), 8, (size_t)n * n_ps, f);
        rd(INTEGER(gid), 4, (size_t)n_gid, f);
        rd(LOGICAL(ctl), 4, (size_t)n, f);
        rd(INTEGER(k), 4, (size_t)nk, f);
        fclose(f);
        typedef SEXP (*fn9)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
        fn9 core = (fn9)lookup("_RUVIIIPRPS_ruvIIICore", 9);
        SEXP out = core(assay, tRep, gid, ctl, k, Rf_ScalarInteger(hdr[4]), Rf_ScalarLogical(hdr[5]), Rf_ScalarReal(pseudo),
                        Rf_ScalarLogical(hdr[6]));
        if (mock_protect_balance != 0) { fprintf(stderr, "PROTECT imbalance %d\n", mock_protect_balance); return 4; }
        FILE* g = fopen(argv[3], "wb");
        if (!g) { perror(argv[3]); return 2; }
        SEXP newY = VECTOR_ELT(out, 0);
        for (int i = 0; i < nk; ++i) write_real(VECTOR_ELT(newY, i), g);
        write_real(VECTOR_ELT(out, 1), g);
        write_real(VECTOR_ELT(out, 2), g);
        int32_t r = INTEGER(VECTOR_ELT(out, 3))[0];
        wr(&r, 4, 1, g);
        fclose(g);
        R_unload_RUVIIIPRPS(&dll);
        return 0;
    }
    if (argc >= 3 && strcmp(argv[1], "prims") == 0) {
        /* A 3 x 4, B 4 x 2, S 3 x 3 (invertible), replicate matrix for fastResidop: rows {0, 1} and {2} */
        SEXP A = Rf_allocMatrix(REALSXP, 3, 4), B = Rf_allocMatrix(REALSXP, 4, 2), S = Rf_allocMatrix(REALSXP, 3, 3),
             M = Rf_allocMatrix(REALSXP, 3, 2);
        for (int i = 0; i < 12; ++i) REAL(A)[i] = (i * 7 % 5) - 1.5 + 0.25 * i;
        for (int i = 0; i < 8; ++i) REAL(B)[i] = (i * 3 % 4) + 0.5 - 0.125 * i;
        const double s[9] = {4, 1, 2, 1, 3, 0, 2, 0, 5};
        memcpy(REAL(S), s, sizeof(s));
        const double m[6] = {1, 1, 0, 0, 0, 1};
        memcpy(REAL(M), m, sizeof(m));
        typedef SEXP (*fn2)(SEXP, SEXP);
        typedef SEXP (*fn1)(SEXP);
        FILE* g = fopen(argv[2], "wb");
        if (!g) { perror(argv[2]); return 2; }
        write_real(A, g); write_real(B, g); write_real(S, g);
        write_real(((fn2)lookup("_RUVIIIPRPS_matrixMult", 2))(A, B), g);
        write_real(((fn2)lookup("_RUVIIIPRPS_matSubtraction", 2))(A, A), g);
        write_real(((fn1)lookup("_RUVIIIPRPS_matTranspose", 1))(A), g);
        write_real(((fn1)lookup("_RUVIIIPRPS_matInverse", 1))(S), g);
        write_real(((fn2)lookup("_RUVIIIPRPS_fastResidop", 2))(A, M), g);
        fclose(g);
        if (mock_protect_balance != 0) { fprintf(stderr, "PROTECT imbalance %d\n", mock_protect_balance); return 4; }
        R_unload_RUVIIIPRPS(&dll);
        return 0;
    }
    fprintf(stderr, "usage: shim_driver list | core in.bin out.bin | prims out.bin\n");
    return 2;
}

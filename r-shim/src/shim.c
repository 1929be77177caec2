/*
 * .Call shim between the RUVIIIPRPS R package and libruviii_b200.so (include/ruviii.h).
 *
 * NOT COMPILED AGAINST R IN THE BUILD CONTAINER (no R there: no Rinternals.h, no libR.so).  It IS compiled and run against
 * a small mock of the R C API (tests/mock_r/: SEXPs on malloc) by tests/test_shim.py -- on the CPU to catch syntax / ABI
 * drift, on the GPU box end to end against the Python binding -- so the argument unpacking, the layouts and the result
 * packing below are exercised; what remains unverified is the real R runtime (PROTECT discipline, R_alloc lifetimes).
 * The file is deliberately free of logic: argument unpacking, one C-ABI call, result packing.
 *
 * Naming follows the package's own (dangling) Rcpp boundary, R/RcppExports.R:4-26:
 *     .Call('_RUVIIIPRPS_<fn>', PACKAGE = 'RUVIIIPRPS', ...)
 * Build:  R CMD SHLIB shim.c -I<repo>/include -L<repo>/ruviii_b200/lib -lruviii_b200  (-> RUVIIIPRPS.so)
 * NAMESPACE gains:  useDynLib(RUVIIIPRPS, .registration = TRUE)
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdlib.h>
#include <string.h>
#include "ruviii.h"

static ruviii_handle* g_handle = NULL;           /* one engine per R session, created on first use */

static ruviii_handle* engine(void) {
    if (!g_handle) {
        int ndev = ruviii_device_count();
        const char* env = getenv("RUVIII_NUM_GPUS");
        if (env && atoi(env) > 0 && atoi(env) < ndev) ndev = atoi(env);
        if (ndev < 1) Rf_error("RUVIIIPRPS: no sm_100 GPU visible; the CUDA engine has no CPU fallback");
        if (ruviii_create(&g_handle, NULL, ndev) != RUVIII_OK) Rf_error("RUVIIIPRPS: %s", ruviii_last_error(NULL));
    }
    return g_handle;
}

/* _RUVIIIPRPS_ruvIIICore(assay, tReplicate, groupId, ctl, k, mode, applyLog, pseudoCount, returnInfo)
 *   assay       numeric n x m1 (genes x samples): assay(se.obj, assay.name)           -- ruvIII.R:88-92
 *   tReplicate  numeric n x n_ps: t(replicate.data)                                   -- ruvIII.R:93
 *   groupId     integer, as.integer(factor(row.names(Y))) (STACKED: m1 + n_ps; FAST: n_ps) -- ruvIII.R:101
 *   ctl         logical n (after tological())                                         -- ruvIII.R:102
 *   k           integer vector (ruvIIIMultipleK's k)
 * returns list(newY = list of n x m1 matrices (one per k; exactly the new assay t(newY[1:m1, ])),
 *              fullalpha = n x r (R side applies t()), W = numeric, blocks of k x m_w (R side applies t()), r)
 * ONE engine call (ruviii_normalise_blocks): every newY block is written by the engine straight into its own R matrix --
 * from R's pageable heap through the engine's pinned bounce buffers, upload / update / download overlapped. */
SEXP _RUVIIIPRPS_ruvIIICore(SEXP assay, SEXP tRep, SEXP groupId, SEXP ctl, SEXP k, SEXP mode, SEXP applyLog,
                            SEXP pseudoCount, SEXP returnInfo) {
    const int n = Rf_nrows(assay), m1 = Rf_ncols(assay), n_ps = Rf_ncols(tRep), nk = LENGTH(k);
    const int info_wanted = Rf_asLogical(returnInfo);
    ruviii_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.struct_size = sizeof(prm);
    prm.mode = Rf_asInteger(mode);
    prm.apply_log = Rf_asLogical(applyLog);
    prm.pseudo_count = Rf_asReal(pseudoCount);
    prm.n_alpha_rows = (info_wanted && prm.mode == RUVIII_MODE_STACKED) ? -1 : 0;
    unsigned char* ctl8 = (unsigned char*)R_alloc(n, 1);
    int n_ctl = 0;
    for (int i = 0; i < n; ++i) { ctl8[i] = LOGICAL(ctl)[i] == TRUE; n_ctl += ctl8[i]; }
    const int n_gid = LENGTH(groupId);
    int* gid = (int*)R_alloc(n_gid > 0 ? n_gid : 1, sizeof(int));
    int p = 0;                                    /* ncol(M): factor codes are dense 1..p */
    for (int i = 0; i < n_gid; ++i) { gid[i] = INTEGER(groupId)[i] - 1; if (INTEGER(groupId)[i] > p) p = INTEGER(groupId)[i]; }
    /* output sizes follow from the plan's own formulas (ruvIII.R:128-129,140,142): r = min(m - ncol(M), sum(ctl)) */
    const int m_for_M = prm.mode == RUVIII_MODE_FAST ? n_ps : m1 + n_ps;
    const int m_w = prm.mode == RUVIII_MODE_FAST ? m1 : m1 + n_ps;
    const int no_rep = p >= m_for_M;
    int r = m_for_M - p;
    if (prm.mode == RUVIII_MODE_STACKED && n_ctl < r) r = n_ctl;
    if (r < 0) r = 0;
    int kmax = 0;
    for (int i = 0; i < nk; ++i) if (INTEGER(k)[i] > kmax) kmax = INTEGER(k)[i];
    const int kmax_eff = kmax < r ? kmax : r;
    const int have_info = info_wanted && !no_rep && kmax_eff > 0;
    R_xlen_t w_len = 0;
    for (int i = 0; i < nk; ++i) w_len += (R_xlen_t)m_w * (INTEGER(k)[i] < r ? INTEGER(k)[i] : r);
    const int fa_rows = prm.n_alpha_rows < 0 ? r : kmax_eff;

    SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
    SEXP newY = PROTECT(Rf_allocVector(VECSXP, nk));
    double** blocks = (double**)R_alloc(nk > 0 ? nk : 1, sizeof(double*));
    for (int i = 0; i < nk; ++i) {
        SEXP mat = PROTECT(Rf_allocMatrix(REALSXP, n, m1));
        SET_VECTOR_ELT(newY, i, mat);
        UNPROTECT(1);                              /* held by newY from here on */
        blocks[i] = REAL(mat);
    }
    SEXP fa = PROTECT(Rf_allocMatrix(REALSXP, n, have_info ? fa_rows : 0));
    SEXP W = PROTECT(Rf_allocVector(REALSXP, have_info ? w_len : 0));
    ruviii_handle* h = engine();
    ruviii_info info;
    if (ruviii_normalise_blocks(h, &prm, REAL(assay), n, m1, n, REAL(tRep), n, n_ps, gid, ctl8, INTEGER(k), nk, blocks, NULL,
                                have_info ? REAL(fa) : NULL, have_info ? REAL(W) : NULL, &info) != RUVIII_OK) {
        UNPROTECT(4);
        Rf_error("RUVIIIPRPS: %s", ruviii_last_error(h));
    }
    if (info.r != r || info.no_replicates != no_rep) {
        UNPROTECT(4);
        Rf_error("RUVIIIPRPS: internal error, the engine planned r = %d where the shim expected %d", info.r, r);
    }
    SET_VECTOR_ELT(out, 0, newY);
    SET_VECTOR_ELT(out, 1, fa);
    SET_VECTOR_ELT(out, 2, W);
    SET_VECTOR_ELT(out, 3, Rf_ScalarInteger(info.r));
    UNPROTECT(4);
    return out;
}

/* RcppExports.R:4-6  fastResidop(A, B): B is a replicate (0/1) matrix; its column index per row is the group */
SEXP _RUVIIIPRPS_fastResidop(SEXP A, SEXP B) {
    const int rows = Rf_nrows(A), n = Rf_ncols(A), p = Rf_ncols(B);
    int* gid = (int*)R_alloc(rows, sizeof(int));
    for (int i = 0; i < rows; ++i) {
        gid[i] = 0;
        for (int j = 0; j < p; ++j) if (REAL(B)[i + (R_xlen_t)j * rows] != 0.0) { gid[i] = j; break; }
    }
    /* the engine works on rows with the genes contiguous: hand it t(A) and transpose back */
    SEXP tA = PROTECT(Rf_allocMatrix(REALSXP, n, rows));
    for (int i = 0; i < rows; ++i) for (int j = 0; j < n; ++j) REAL(tA)[j + (R_xlen_t)i * n] = REAL(A)[i + (R_xlen_t)j * rows];
    SEXP tO = PROTECT(Rf_allocMatrix(REALSXP, n, rows));
    if (ruviii_fast_residop(engine(), REAL(tA), rows, n, gid, REAL(tO)) != RUVIII_OK) {
        UNPROTECT(2);
        Rf_error("RUVIIIPRPS: %s", ruviii_last_error(g_handle));
    }
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, rows, n));
    for (int i = 0; i < rows; ++i) for (int j = 0; j < n; ++j) REAL(out)[i + (R_xlen_t)j * rows] = REAL(tO)[j + (R_xlen_t)i * n];
    UNPROTECT(3);
    return out;
}
SEXP _RUVIIIPRPS_fastResidop2(SEXP A, SEXP B) { return _RUVIIIPRPS_fastResidop(A, B); }

/* RcppExports.R:12-26  the four generic primitives (never called by the package; provided for completeness) */
SEXP _RUVIIIPRPS_matrixMult(SEXP A, SEXP B) {
    const int m = Rf_nrows(A), kk = Rf_ncols(A), n = Rf_ncols(B);
    if (Rf_nrows(B) != kk) Rf_error("non-conformable arguments");
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, m, n));
    if (ruviii_mat_mult(engine(), REAL(A), m, kk, REAL(B), n, REAL(out)) != RUVIII_OK) {
        UNPROTECT(1);
        Rf_error("RUVIIIPRPS: %s", ruviii_last_error(g_handle));
    }
    UNPROTECT(1);
    return out;
}
SEXP _RUVIIIPRPS_matSubtraction(SEXP A, SEXP B) {
    const int m = Rf_nrows(A), n = Rf_ncols(A);
    if (Rf_nrows(B) != m || Rf_ncols(B) != n) Rf_error("non-conformable arrays");
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, m, n));
    if (ruviii_mat_subtract(engine(), REAL(A), REAL(B), (int64_t)m * n, REAL(out)) != RUVIII_OK) {
        UNPROTECT(1);
        Rf_error("RUVIIIPRPS: %s", ruviii_last_error(g_handle));
    }
    UNPROTECT(1);
    return out;
}
SEXP _RUVIIIPRPS_matTranspose(SEXP A) {
    const int m = Rf_nrows(A), n = Rf_ncols(A);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, m));
    if (ruviii_mat_transpose(engine(), REAL(A), m, n, REAL(out)) != RUVIII_OK) {
        UNPROTECT(1);
        Rf_error("RUVIIIPRPS: %s", ruviii_last_error(g_handle));
    }
    UNPROTECT(1);
    return out;
}
SEXP _RUVIIIPRPS_matInverse(SEXP A) {
    const int n = Rf_nrows(A);
    if (Rf_ncols(A) != n) Rf_error("'a' must be a square matrix");
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, n));
    if (ruviii_mat_inverse(engine(), REAL(A), n, REAL(out)) != RUVIII_OK) {
        UNPROTECT(1);
        Rf_error("RUVIIIPRPS: %s", ruviii_last_error(g_handle));
    }
    UNPROTECT(1);
    return out;
}

/* ruvIII.R:116 / fastruvIII.R:133-134  RUV1(Y, eta, ctl): tEta = t(design.matrix(eta, include.intercept)) is q x n in R, whose
 * column-major bytes are n x q row-major -- the engine wants q x n row-major, i.e. the bytes of the n x q design matrix
 * itself, so the R side passes design.matrix(eta) (n x q) untransposed.  NULL clears. */
SEXP _RUVIIIPRPS_setEta(SEXP design) {
    ruviii_handle* h = engine();
    int rc = Rf_isNull(design) ? ruviii_set_eta(h, NULL, 0, 0)
                               : ruviii_set_eta(h, REAL(design), Rf_ncols(design), Rf_nrows(design));
    if (rc != RUVIII_OK) Rf_error("RUVIIIPRPS: %s", ruviii_last_error(h));
    return R_NilValue;
}

/* prpsForCategoricalUV.R:123-139 / prpsForContinuousUV.R:190-201: every pseudo-sample is rowMeans(expre.data[, index]).
 * setStart (length n_ps + 1, 0-based CSR offsets) and setRows (0-based column indices of the assay) come from the
 * table() / dplyr bookkeeping that stays in R.  Returns genes x pseudo-samples, as metadata$PRPS stores it. */
SEXP _RUVIIIPRPS_prpsMeans(SEXP assay, SEXP setStart, SEXP setRows, SEXP applyLog, SEXP pseudoCount) {
    const int n = Rf_nrows(assay), m1 = Rf_ncols(assay), n_ps = LENGTH(setStart) - 1;
    ruviii_handle* h = engine();
    SEXP ps = PROTECT(Rf_allocMatrix(REALSXP, n, n_ps));
    if (ruviii_prps(h, REAL(assay), n, m1, n, Rf_asLogical(applyLog), Rf_asReal(pseudoCount), INTEGER(setStart),
                    INTEGER(setRows), n_ps, REAL(ps), n) != RUVIII_OK) {
        UNPROTECT(1);
        Rf_error("RUVIIIPRPS: %s", ruviii_last_error(h));
    }
    UNPROTECT(1);
    return ps;
}

/* computePCA.R:128-134  runSVD(x = transpose(temp.data), k = nb.pcs, center, scale): list(u = m1 x k, d, v = n x k).
 * The engine returns u as m1 x k row-major = the bytes of a k x m1 R matrix (the R side applies t()); v comes back as
 * k x n row-major = exactly R's n x k column-major v. */
SEXP _RUVIIIPRPS_fastPCA(SEXP assay, SEXP k, SEXP applyLog, SEXP pseudoCount, SEXP center, SEXP scale) {
    const int n = Rf_nrows(assay), m1 = Rf_ncols(assay), kk = Rf_asInteger(k);
    ruviii_handle* h = engine();
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SEXP tu = PROTECT(Rf_allocMatrix(REALSXP, kk, m1)), d = PROTECT(Rf_allocVector(REALSXP, kk)),
         v = PROTECT(Rf_allocMatrix(REALSXP, n, kk));
    if (ruviii_pca(h, REAL(assay), n, m1, n, -1, Rf_asLogical(applyLog), Rf_asReal(pseudoCount), Rf_asLogical(center),
                   Rf_asLogical(scale), kk, REAL(tu), REAL(d), REAL(v), NULL) != RUVIII_OK) {
        UNPROTECT(4);
        Rf_error("RUVIIIPRPS: %s", ruviii_last_error(h));
    }
    SET_VECTOR_ELT(out, 0, tu);
    SET_VECTOR_ELT(out, 1, d);
    SET_VECTOR_ELT(out, 2, v);
    UNPROTECT(4);
    return out;
}

/* supervisedFindNCG.R:179-191,214-229: per-gene one-way ANOVA F (matrixTests::row_oneway_equalvar(...)$statistic) and / or
 * Pearson correlation (Rfast::correls(..., type = "pearson")) of every gene with one variable.
 *   groups     integer m1 (0-based level per sample, < 0 = sample left out) or NULL
 *   covariate  numeric m1 or NULL
 * returns list(F = numeric n or NULL, r = numeric n or NULL) */
SEXP _RUVIIIPRPS_geneStats(SEXP assay, SEXP groups, SEXP covariate, SEXP applyLog, SEXP pseudoCount) {
    const int n = Rf_nrows(assay), m1 = Rf_ncols(assay);
    const int has_g = !Rf_isNull(groups), has_c = !Rf_isNull(covariate);
    if ((has_g && LENGTH(groups) != m1) || (has_c && LENGTH(covariate) != m1)) Rf_error("one entry per sample is needed");
    ruviii_handle* h = engine();
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SEXP F = PROTECT(has_g ? Rf_allocVector(REALSXP, n) : R_NilValue), r = PROTECT(has_c ? Rf_allocVector(REALSXP, n) : R_NilValue);
    if (ruviii_gene_stats(h, REAL(assay), n, m1, n, -1, Rf_asLogical(applyLog), Rf_asReal(pseudoCount), has_g ? INTEGER(groups) : NULL,
                          has_c ? REAL(covariate) : NULL, has_g ? REAL(F) : NULL, has_c ? REAL(r) : NULL, NULL) != RUVIII_OK) {
        UNPROTECT(3);
        Rf_error("RUVIIIPRPS: %s", ruviii_last_error(h));
    }
    SET_VECTOR_ELT(out, 0, F);
    SET_VECTOR_ELT(out, 1, r);
    UNPROTECT(3);
    return out;
}

/* expr.data.reg <- t(lm(t(expr.data) ~ ...)$residuals), supervisedFindNCG.R:111-133.
 *   assay   numeric genes x samples (n x m1, column-major = m1 x n row-major)
 *   design  numeric m1 x p: model.matrix(~ v1 + v2 + ..., colData) built in R
 * returns the residuals as a genes x samples matrix (they also stay on the device for geneStats-like calls with slot = -2) */
SEXP _RUVIIIPRPS_regressOut(SEXP assay, SEXP design, SEXP applyLog, SEXP pseudoCount) {
    const int n = Rf_nrows(assay), m1 = Rf_ncols(assay), p = Rf_ncols(design);
    if (Rf_nrows(design) != m1) Rf_error("the model matrix needs one row per sample");
    ruviii_handle* h = engine();
    double* d = (double*)R_alloc((size_t)m1 * p, sizeof(double));      /* column-major -> the row-major m1 x p the ABI takes */
    for (int j = 0; j < p; ++j)
        for (int i = 0; i < m1; ++i) d[(size_t)i * p + j] = REAL(design)[(size_t)j * m1 + i];
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, m1));
    if (ruviii_regress_out(h, REAL(assay), n, m1, n, -1, Rf_asLogical(applyLog), Rf_asReal(pseudoCount), d, p, REAL(out), n, NULL, NULL) !=
        RUVIII_OK) {
        UNPROTECT(1);
        Rf_error("RUVIIIPRPS: %s", ruviii_last_error(h));
    }
    UNPROTECT(1);
    return out;
}

static const R_CallMethodDef call_methods[] = {
    {"_RUVIIIPRPS_ruvIIICore", (DL_FUNC)&_RUVIIIPRPS_ruvIIICore, 9},
    {"_RUVIIIPRPS_fastResidop", (DL_FUNC)&_RUVIIIPRPS_fastResidop, 2},
    {"_RUVIIIPRPS_fastResidop2", (DL_FUNC)&_RUVIIIPRPS_fastResidop2, 2},
    {"_RUVIIIPRPS_matrixMult", (DL_FUNC)&_RUVIIIPRPS_matrixMult, 2},
    {"_RUVIIIPRPS_matSubtraction", (DL_FUNC)&_RUVIIIPRPS_matSubtraction, 2},
    {"_RUVIIIPRPS_matTranspose", (DL_FUNC)&_RUVIIIPRPS_matTranspose, 1},
    {"_RUVIIIPRPS_matInverse", (DL_FUNC)&_RUVIIIPRPS_matInverse, 1},
    {"_RUVIIIPRPS_setEta", (DL_FUNC)&_RUVIIIPRPS_setEta, 1},
    {"_RUVIIIPRPS_prpsMeans", (DL_FUNC)&_RUVIIIPRPS_prpsMeans, 5},
    {"_RUVIIIPRPS_fastPCA", (DL_FUNC)&_RUVIIIPRPS_fastPCA, 6},
    {"_RUVIIIPRPS_geneStats", (DL_FUNC)&_RUVIIIPRPS_geneStats, 5},
    {"_RUVIIIPRPS_regressOut", (DL_FUNC)&_RUVIIIPRPS_regressOut, 4},
    {NULL, NULL, 0}};

void R_init_RUVIIIPRPS(DllInfo* dll) {
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
void R_unload_RUVIIIPRPS(DllInfo* dll) {
    (void)dll;
    if (g_handle) { ruviii_destroy(g_handle); g_handle = NULL; }
}

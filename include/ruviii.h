/*
 * ruviii.h -- C ABI of the B200-native RUV-III normalisation engine (libruviii_b200.so).
 *
 * This is the drop-in boundary for the hot path of RMolania/RUVIIIPRPS: the arithmetic of
 *   ruvIII()          R/ruvIII.R:116-149
 *   fastruvIII()      R/fastruvIII.R:133-180
 *   ruvIIIMultipleK() R/ruvIIIMultipleK.R:75-94  (one call computes every k; the reference loops ruvIII per k)
 * as driven by normalise() (R/normalise.R:143-161).  The reference declares, but never implemented, a
 * native boundary in R/RcppExports.R:4-26 (`.Call('_RUVIIIPRPS_<fn>', PACKAGE='RUVIIIPRPS', ...)`); the
 * entry points below are what an `_RUVIIIPRPS_*` .Call shim binds (see INTEGRATION.md and r-shim/).
 *
 * Plain C: pointers and sizes only.  No CPU fallback exists: every compute entry point needs a
 * CUDA device of compute capability 10.0 (sm_100a) and returns RUVIII_ERR_CUDA otherwise.
 *
 * Layout conventions (all FP64):
 *   - An R matrix is column-major.  The assay is genes x samples (n x m1) column-major, which is the same
 *     bytes as "samples x genes row-major with the genes contiguous".  Every matrix argument below is
 *     described in that row-major reading, so `Y` IS the assay buffer and `newY_out` IS the buffer of the
 *     new assay `t(newY[1:m1, ])` (ruvIII.R:158): no transposition on either side.
 *   - `PS` is `replicate.data` (pseudo-samples x genes, normalise.R:143) in the same row-major reading,
 *     i.e. R's `t(replicate.data)` buffer -- the genes x pseudo-samples matrix normalise.R:143 starts from.
 *   - `fullalpha_out` is rows x n row-major (R: t(fullalpha)); `W_out` blocks are m_w x k row-major (R: t(W)).
 *   - Inputs are read-only, never written, never retained after the call returns.
 */
#ifndef RUVIII_B200_H
#define RUVIII_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RUVIII_ABI_VERSION 1

/* status codes */
#define RUVIII_OK               0
#define RUVIII_ERR_INVALID      1   /* bad argument / dimension mismatch (len(ctl) != n, group id < 0, k < 0, ...) */
#define RUVIII_ERR_CUDA         2   /* CUDA runtime failure, or no sm_100 device present                        */
#define RUVIII_ERR_NCCL         3
#define RUVIII_ERR_SOLVER       4   /* cuSOLVER failure / eigensolver did not converge                          */
#define RUVIII_ERR_SINGULAR     5   /* ac %*% t(ac) is not positive definite: R's solve() (ruvIII.R:144) errors  */
#define RUVIII_ERR_UNSUPPORTED  6
#define RUVIII_ERR_STATE        7   /* resident API called out of order                                          */

/* which R entry point's semantics to follow (they are NOT interchangeable) */
#define RUVIII_MODE_STACKED 0  /* ruvIII.R: Y = rbind(samples, PRPS); mu over all m rows; W has m rows            */
#define RUVIII_MODE_FAST    1  /* fastruvIII.R: Y = samples; M from PRPS rows only; mu over m1 rows; W has m1 rows */

typedef struct ruviii_handle ruviii_handle;

typedef struct ruviii_params {
    int32_t struct_size;     /* sizeof(ruviii_params), for forward compatibility                              */
    int32_t mode;            /* RUVIII_MODE_*                                                                  */
    int32_t apply_log;       /* ruvIII.R:88-92: log2(assay + pseudo_count) on the assay rows (not on PS)       */
    int32_t average_rep;     /* ruvIII.R:148-149 (apply.average.rep); 0 = off (default); see ruviii_download_averaged */
    double  pseudo_count;    /* ruvIII.R:42                                                                    */
    int32_t n_alpha_rows;    /* rows of fullalpha to produce: 0 = max(ks) (all the update needs);              *
                              * -1 = r = min(m - ncol(M), sum(ctl)) as ruvIII.R:140 keeps (return.info = TRUE)  */
    int32_t gram_side;       /* 0 = auto (smaller side), 1 = sample side Y0 Y0', 2 = gene side Y0' Y0          */
    int32_t want_ps_rows;    /* also produce newY for the PRPS rows (STACKED, return.info & !save.se.obj)       */
    int32_t keep_resident;   /* ruviii_plan only: the matrices uploaded under the PREVIOUS plan of this handle stay (same m1, n,  *
                              * n_ps, apply_log / pseudo_count): re-plan with another ctl / k / grouping without another upload  *
                              * -- normalise() learns ctl from statistics of the assay it has already uploaded                   */
    int32_t reserved[6];
} ruviii_params;

/* what one run did; times are device times (CUDA events on the compute stream of shard 0), in ms */
typedef struct ruviii_info {
    int32_t struct_size;
    int32_t m, m1, n_ps, n;            /* rows of the stacked matrix, samples, pseudo-samples, genes (local)   */
    int32_t p;                         /* ncol(M): number of replicate groups incl. singletons                  */
    int32_t active_rows;               /* rows belonging to non-singleton groups (nonzero rows of Y0)           */
    int32_t r;                         /* min(m - p, sum(ctl))  (ruvIII.R:140)                                  */
    int32_t n_ctl;                     /* sum(ctl) over all shards                                              */
    int32_t kmax_eff;                  /* min(max(ks), r)  (ruvIII.R:142)                                       */
    int32_t no_replicates;             /* 1 when ncol(M) >= m: newY = Y (ruvIII.R:128-129)                      */
    int32_t gram_side;                 /* 1 sample side, 2 gene side                                            */
    int32_t n_shards;                  /* GPUs used by this handle (this process)                               */
    int32_t kernel_launches;           /* kernels of this library launched by the last run                      */
    float ms_log, ms_residop, ms_gram, ms_gram_reduce, ms_allreduce_gram, ms_eigen, ms_bcast,
          ms_project, ms_ctl, ms_allreduce_ctl, ms_solve, ms_update, ms_total;
    float ms_h2d, ms_d2h;              /* host<->device copies (only in ruviii_normalise / upload / download)   */
    double eig_top, eig_kmax, eig_next;/* lambda_1, lambda_kmax, lambda_kmax+1 (gap that conditions newY)       */
    int32_t eig_solver;                /* 0 none, 1 cuSOLVER syevd, 2 cuSOLVER syevdx, 3 block subspace iteration   */
    int32_t eig_iterations;            /* subspace-iteration passes (solver 3); 0 otherwise                        */
    int32_t gram_order;                /* order of the Gram actually formed: active_rows - #non-singleton groups    */
    int32_t pipelined;                 /* ruviii_normalise overlapped upload, compute and download: 0 no, 1 zero-copy pipeline (pinned),
                                        * 2 staged path on pinned buffers, 3 staged path through pinned bounce buffers (pageable) */
    double host_t_enter, host_t_done;  /* CLOCK_MONOTONIC seconds at ruviii_run entry and when its last stream drained */
    float ms_gather;                   /* control-gene gather on its own stream (overlaps the eigensolve; not part of the sum) */
    int32_t eig_rr;                    /* solver 3, persistent form: Rayleigh-Ritz steps                                      */
    int32_t eig_filter_products;       /*   products spent inside Chebyshev filter rounds (part of eig_iterations)             */
    int32_t eig_chol2;                 /*   Cholesky-QR rounds that needed a second pass                                      */
} ruviii_info;

/* ---- lifecycle ------------------------------------------------------------------------------------- */

int  ruviii_abi_version(void);

/* Number of visible sm_100 devices (0 when there is no GPU; never fails). */
int  ruviii_device_count(void);

/* One process driving `ndev` GPUs (what an R session does): genes are sharded over the devices internally,
 * NCCL communicators are created with ncclCommInitAll.  devices == NULL means 0..ndev-1. */
int  ruviii_create(ruviii_handle** out, const int32_t* devices, int32_t ndev);

/* One process per GPU (torchrun-style launch).  Every rank passes the 128-byte id rank 0 obtained from
 * ruviii_nccl_unique_id().  Each rank then calls the entry points below with ITS gene shard
 * (n = local genes, ctl/Y/PS/newY restricted to those columns); W and eigenvalues are replicated. */
int  ruviii_nccl_unique_id(void* out128);
int  ruviii_create_rank(ruviii_handle** out, int32_t device, int32_t rank, int32_t nranks,
                        const void* nccl_unique_id128);

void ruviii_destroy(ruviii_handle* h);

/* Message of the last failure on this handle (or of the last failed create when h == NULL). */
const char* ruviii_last_error(const ruviii_handle* h);

/* ---- one-shot, host buffers in and out: what the R shim calls ------------------------------------------
 *
 * Replaces ruvIII.R:116-149 / fastruvIII.R:133-180 and the per-k loop of ruvIIIMultipleK.R:75-94.
 *   Y            m1 x n   (row stride ldY   >= n)  the assay, see layout note
 *   PS           n_ps x n (row stride ldPS  >= n)  replicate.data
 *   group_of_row length m1 + n_ps in STACKED mode (factor codes of row.names(Y), ruvIII.R:101; any ids >= 0,
 *                need not be dense), length n_ps in FAST mode (row.names(replicate.data), fastruvIII.R:117)
 *   ctl          length n, nonzero = negative-control gene (tological(), ruvIII.R:80-84,102)
 *   ks, nk       the k values (ruvIIIMultipleK's vector); k = 0 copies Y (ruvIII.R:134-136)
 *   newY_out     nk blocks of m1 x n (row stride n), block i = result for ks[i]
 *   newY_ps_out  NULL, or nk blocks of n_ps x n (STACKED with want_ps_rows)
 *   fullalpha_out NULL, or rows x n with rows = info->kmax_eff (n_alpha_rows = 0) or info->r (-1)
 *   W_out        NULL, or nk blocks back to back, block i = m_w x min(ks[i], r) row-major, m_w = m (STACKED) / m1 (FAST)
 *   info         NULL or filled on return
 */
int ruviii_normalise(ruviii_handle* h, const ruviii_params* params,
                     const double* Y, int64_t ldY, int32_t m1, int32_t n,
                     const double* PS, int64_t ldPS, int32_t n_ps,
                     const int32_t* group_of_row, const uint8_t* ctl,
                     const int32_t* ks, int32_t nk,
                     double* newY_out, double* newY_ps_out,
                     double* fullalpha_out, double* W_out, ruviii_info* info);

/* Same call with one output pointer PER k (block i = newY for ks[i], m1 x n, row stride n): what the R shim uses, so that
 * every RUV_K<k> assay is written straight into its own R-allocated matrix (no intermediate nk x m1 x n buffer). */
int ruviii_normalise_blocks(ruviii_handle* h, const ruviii_params* params,
                            const double* Y, int64_t ldY, int32_t m1, int32_t n,
                            const double* PS, int64_t ldPS, int32_t n_ps,
                            const int32_t* group_of_row, const uint8_t* ctl,
                            const int32_t* ks, int32_t nk,
                            double* const* newY_blocks, double* newY_ps_out,
                            double* fullalpha_out, double* W_out, ruviii_info* info);

/* ---- resident API: plan once, keep Y on the device, run many times ------------------------------------- */

int ruviii_plan(ruviii_handle* h, const ruviii_params* params, int32_t m1, int32_t n, int32_t n_ps,
                const int32_t* group_of_row, const uint8_t* ctl, const int32_t* ks, int32_t nk);
int ruviii_upload(ruviii_handle* h, const double* Y, int64_t ldY, const double* PS, int64_t ldPS);
int ruviii_run(ruviii_handle* h, ruviii_info* info);
int ruviii_download(ruviii_handle* h, double* newY_out, double* newY_ps_out,
                    double* fullalpha_out, double* W_out);
/* apply.average.rep (ruvIII.R:148-149): nk blocks of p x n (row stride n), row g = mean of the rows of newY in
 * replicate group g, groups in ascending id order (= factor level order when ids are as.integer(factor(names))). */
int ruviii_download_averaged(ruviii_handle* h, double* averaged_out);
/* replicate.data as it sits on the device (n_ps x n, row stride ldPS): uploaded, or built from the assay by the sets given to
 * ruviii_set_prps_sets (what normalise() stores in metadata$PRPS, normalise.R:143 / supervisedPRPS.R). */
int ruviii_download_replicate_data(ruviii_handle* h, double* PS_out, int64_t ldPS);
/* Small replicated results of the last run: eigenvalues (descending, `count` of them are written). */
int ruviii_eigenvalues(ruviii_handle* h, double* out, int32_t count);

/* Skip residop/Gram/eigensolve and start from a previously returned fullalpha (ruvIII.R:138, the reference's
 * only reuse mechanism).  rows x n row-major; must be called after ruviii_plan and before ruviii_run. */
int ruviii_set_fullalpha(ruviii_handle* h, const double* fullalpha, int32_t rows);

/* ruv::RUV1 pre-step (ruvIII.R:116; fastruvIII.R:133-134 applies it to Y and to replicate.data).  ruv is not vendored in the
 * reference tree; the statement restated here is ruv 0.9.7.1's
 *     Y - Y[, ctl] %*% t(eta[, ctl]) %*% solve(eta[, ctl] %*% t(eta[, ctl])) %*% eta       with eta q x n.
 * eta: q x n row-major, i.e. the bytes of R's n x q gene-wise design matrix, the intercept column already included by
 * the caller when include.intercept = TRUE (design.matrix()).  Sticky: it applies to every later plan / ruviii_normalise
 * on this handle (n must match; with one process per GPU pass the local gene slab of eta).  eta = NULL or q = 0 clears
 * it (the reference default eta = NULL: RUV1 is the identity).  q <= 32.  A singular eta_c eta_c' is RUVIII_ERR_SINGULAR. */
int ruviii_set_eta(ruviii_handle* h, const double* eta, int32_t q, int32_t n);

/* ---- PRPS construction (the step immediately before the path) -----------------------------------
 * prpsForCategoricalUV.R:123-139 and prpsForContinuousUV.R:190-201 build every pseudo-sample as
 *     rowMeans(expre.data[, index.sample])          with expre.data = log2(assay + pseudo.count) when apply.log.
 * The index sets (which samples share a biology x batch cell, or sit at the top / bottom of a continuous variable) are
 * host-side bookkeeping on colData and stay there; the means are this segmented kernel.  CSR: pseudo-sample j averages
 * the samples set_rows[set_start[j] .. set_start[j+1]) (0-based rows of Y = columns of the assay).
 *
 * ruviii_prps: one shot, host buffers (only the member rows are uploaded).  PS_out is n_ps x n row-major, i.e. the bytes
 * of the genes x pseudo-samples matrix the reference stores in metadata$PRPS. */
int ruviii_prps(ruviii_handle* h, const double* Y, int64_t ldY, int32_t m1, int32_t n, int32_t apply_log, double pseudo_count,
                const int32_t* set_start, const int32_t* set_rows, int32_t n_ps, double* PS_out, int64_t ldPS);
/* Fused form: register the sets (sticky, like ruviii_set_eta; n_ps = 0 clears) and pass PS = NULL to ruviii_upload /
 * ruviii_normalise with the same n_ps: replicate.data is then built on the device from the uploaded assay (after its
 * log transform, before RUV1) and never crosses PCIe.  The group ids of the PRPS rows still come in group_of_row. */
int ruviii_set_prps_sets(ruviii_handle* h, const int32_t* set_start, const int32_t* set_rows, int32_t n_ps);

/* ---- computePCA fast path (runs on every RUV_K* assay right after the path, normAssessment.R:113) ----------
 * computePCA.R:120-134: runSVD(x = t(assay), k = nb.pcs, center, scale): top-k singular triplets of the samples x genes matrix
 * after centring every gene over the samples (and dividing by its sd when scale, as base::scale does).
 *   X != NULL : host matrix, m1 x n row-major (the assay bytes), log2(X + pseudo_count) first when apply_log (computePCA.R:124-127)
 *   X == NULL : data already on the device -- slot >= 0: output block `slot` (the newY of ks[slot]) of the last ruviii_run;
 *               slot == -1: the uploaded assay after the plan's own log transform (applied in place now if no run has done it
 *               yet; apply_log / pseudo_count of this call are ignored).  m1, n, ldX are ignored.
 *   u_out  m1 x k row-major (sing.val$u), d_out k (sing.val$d), v_out NULL or k x n row-major (= the bytes of R's n x k v).
 *   Signs: each u_j has its largest-magnitude component positive; v_j follows.  k < m1 up to 32 (the device's own top-k
 *   eigensolver); 32 < k <= min(m1, n) takes the library eigensolver for the k leading pairs and projects in chunks of 32:
 *   k = min(m1, n) is the full svd(scale(t(x))) of computePCA.R:150-190 (fast.pca = FALSE).
 *   stage_ms6: NULL or 6 floats (device ms): upload+log, centre, Gram, allreduce+eigensolve, broadcast+projection, total. */
int ruviii_pca(ruviii_handle* h, const double* X, int64_t ldX, int32_t m1, int32_t n, int32_t slot, int32_t apply_log,
               double pseudo_count, int32_t center, int32_t scale, int32_t k, double* u_out, double* d_out, double* v_out,
               float* stage_ms6);

/* ---- negative-control-gene statistics (the other step immediately before the path) ---------------------
 * supervisedFindNCG.R:179-191,326-341: matrixTests::row_oneway_equalvar(x, g)$statistic -- the one-way ANOVA F statistic of
 * every gene across the levels of a categorical variable; :214-229,270-285: Rfast::correls(y, x, type = "pearson") -- the
 * Pearson correlation of every gene with a continuous variable.  Neither package is vendored in the reference; the
 * definitions are the textbook ones (stats.cu).  Ranking, rounding and the rank product (:342-357) stay on the host.
 *   X / slot         as in ruviii_pca (host matrix m1 x n row-major, or data resident on the device)
 *   group_of_sample  NULL, or m1 ids; samples with id < 0 are left out (levels with < 3 samples, :183-185); needs F_out (n)
 *   covariate        NULL, or m1 values; needs r_out (n).  With both, r is taken over the samples the ANOVA keeps.
 *   ms_out           NULL or the device time of the pass */
int ruviii_gene_stats(ruviii_handle* h, const double* X, int64_t ldX, int32_t m1, int32_t n, int32_t slot, int32_t apply_log,
                      double pseudo_count, const int32_t* group_of_sample, const double* covariate, double* F_out, double* r_out,
                      float* ms_out);

/* Spearman's rho of every gene with a continuous variable: supervisedFindNCG.R:214-229,270-285 with corr.method = "spearman"
 * (Rfast::correls(type = "spearman"): Pearson's r of the ranks, ties averaged).  X / slot as in ruviii_pca; a monotone transform
 * such as log2(x + pseudo.count) does not change ranks, so there is no apply_log.  m1 <= 16384 (one gene's samples are sorted in
 * shared memory); more is RUVIII_ERR_UNSUPPORTED. */
int ruviii_gene_spearman(ruviii_handle* h, const double* X, int64_t ldX, int32_t m1, int32_t n, int32_t slot, const double* covariate,
                         double* rho_out, float* ms_out);

/* regress.out.uv.variables / regress.out.bio.variables: supervisedFindNCG.R:111-133,
 *     y <- lm(t(expr.data) ~ se.obj$v1 + se.obj$v2 + ...);  expr.data.reg <- t(y$residuals)
 * X / slot / apply_log as in ruviii_pca.  design: the m1 x p model matrix of the formula, row-major, HOST memory (intercept
 * column, treatment dummies of every factor, the continuous variables as they are): residuals depend on its column space only.
 * The device builds an orthonormal basis with lm's own rank rule (a column whose norm falls below 1e-7 of its original norm
 * once the earlier columns are removed is aliased and dropped; *rank_out = columns kept) and removes it from every gene.
 * The residuals (m1 x n) stay on the device: ruviii_gene_stats / ruviii_gene_spearman read them with X = NULL, slot = -2
 * until the next ruviii_regress_out; residuals_out (NULL, or m1 x n row-major with pitch ld_out) also receives them.
 * p <= 256 and p < m1. */
int ruviii_regress_out(ruviii_handle* h, const double* X, int64_t ldX, int32_t m1, int32_t n, int32_t slot, int32_t apply_log,
                       double pseudo_count, const double* design, int32_t p, double* residuals_out, int64_t ld_out, int32_t* rank_out,
                       float* ms_out);

/* ---- the six primitives RcppExports.R:4-26 declares, backed by the same device kernels ------------------
 * Column-major (R) matrices, host pointers.  fastResidop(A, B) = A - B (B'B)^-1 B'A where B is a 0/1
 * replicate matrix (fastruvIII.R:88-95): B is passed as its factor codes. */
int ruviii_fast_residop(ruviii_handle* h, const double* A_rowmajor_rows_x_n, int32_t rows, int32_t n,
                        const int32_t* group_of_row, double* out);

/* The other four primitives of RcppExports.R:12-26 -- matrixMult(A, B), matSubtraction(A, B), matTranspose(A), matInverse(A).
 * The reference declares them (`.Call('_RUVIIIPRPS_matrixMult', ...)` etc.) but ships no implementation and never calls them;
 * they are provided so that the declared boundary is complete.  Column-major (R) host matrices, first GPU of the handle;
 * none is on the hot path: the product is a plain cuBLAS Dgemm, the inverse cuSOLVER getrf + getrs against the identity
 * (LAPACK's dgesv route, what R's solve(A) runs; an exactly singular matrix is RUVIII_ERR_SINGULAR, R stops too). */
int ruviii_mat_mult(ruviii_handle* h, const double* A, int32_t m, int32_t k, const double* B, int32_t n, double* out /* m x n */);
int ruviii_mat_subtract(ruviii_handle* h, const double* A, const double* B, int64_t count, double* out);
int ruviii_mat_transpose(ruviii_handle* h, const double* A, int32_t m, int32_t n, double* out /* n x m */);
int ruviii_mat_inverse(ruviii_handle* h, const double* A, int32_t n, double* out /* n x n */);

/* Page-locked host memory.  ruviii_normalise overlaps upload, compute and download for pageable and page-locked buffers
 * alike (pageable ones go through pinned bounce buffers filled by a pool of host threads, RUVIII_HOST_THREADS); with
 * page-locked Y / newY_out (these helpers, cudaHostAlloc or cudaHostRegister) the DMA engines use the caller's memory
 * directly.  Returns NULL when no CUDA device is present. */
void* ruviii_host_alloc(size_t bytes);
void  ruviii_host_free(void* p);

/* Device-side barrier over all shards of all processes (a one-element allreduce followed by a stream drain): aligns
 * the ranks to within microseconds, which a host-side (gloo/MPI) barrier does not.  No-op for a single shard. */
int ruviii_barrier(ruviii_handle* h);

/* ---- micro-benchmarks used by bench.py to state practical ceilings beside the roofline ----------------- */
/* which: 0 = device copy GB/s (bytes read+written), 1 = cuBLAS Dgemm TFLOP/s at size^3, 2 = DMMA issue peak TFLOP/s,
 *        3 = ms to gather 11,000 x `size` scattered doubles out of pinned host memory (zero-copy over PCIe),
 *        4 = GB/s of copy variant `size` (0 linear; 1..7 the update kernel's tiling with different rows in flight,
 *            rows per CTA and cache hints) over an 11,000 x 20,000 matrix,
 *        5 = median us of an in-place FP64 allreduce of `size` KiB over all shards (every rank must call),
 *        6 = aggregate GB/s of `size` MiB host->device plus `size` MiB device->host copied simultaneously (PCIe duplex) */
int ruviii_probe(ruviii_handle* h, int32_t which, int32_t size, double* value_out);

#ifdef __cplusplus
}
#endif
#endif /* RUVIII_B200_H */

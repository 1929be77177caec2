// Launchers of the hand-written sm_100a kernels of the RUV-III hot path.  Each one cites the R
// statement it replaces; DESIGN.md gives the roofline and algorithmic bytes of each.
#pragma once
#include "common.cuh"

namespace ruviii {

// Rows of the stacked matrix Y = rbind(t(expr.data), replicate.data) (ruvIII.R:93) are never stacked
// physically: rows [0, m1) live in the assay buffer, rows [m1, m) in the replicate.data buffer.
struct RowSrc {
    const double* Y;
    int64_t ldY;
    const double* PS;
    int64_t ldPS;
    int m1;
    __host__ __device__ const double* row(int i) const {
        return i < m1 ? Y + (int64_t)i * ldY : PS + (int64_t)(i - m1) * ldPS;
    }
};

// K8  ruvIII.R:89  log2(assay + pseudo.count), in place on the device copy of the assay.
void launch_log2_inplace(double* Y, int64_t ld, int rows, int n, double pseudo_count, cudaStream_t s);

// K1  ruvIII.R:139 residop(Y, M) / fastruvIII.R:88-95.  M is one-hot, so Y0 = Y - group mean; rows of
// singleton groups are exactly zero and are not produced.  grp_start/grp_rows: CSR of the non-singleton
// groups; compact row (grp position) t of Y0 is the residual of stacked row grp_rows[t].
void launch_residop(RowSrc src, const int* grp_start, const int* grp_rows, int n_groups, int max_group,
                    int n, double* Y0, int64_t ldY0, cudaStream_t s);
// K1, compact form used by the pipeline: the rows of one replicate group sum to zero, so residop's output for a
// group of s members has rank s - 1.  Z = H Y0 with H the orthonormal Helmert contrasts of every group
//     z_j = sqrt(j / (j + 1)) (mean(y_1..y_j) - y_{j+1}),  j = 1..s-1
// has sum(s_g - 1) = m - ncol(M) rows, the same nonzero singular values as Y0 and Y0 = H'Z, hence
// t(U) %*% Y0 = t(U_z) %*% Z for the left singular vectors U_z of Z: Gram, eigensolve and projection all run on
// the smaller Z (800 instead of 1,200 rows at C3: 2.25x fewer Gram flops).  Row (grp_start[g] - g + j - 1) of Z
// is contrast j of group g.
void launch_helmert(RowSrc src, const int* grp_start, const int* grp_rows, int n_groups, int max_group, int n, double* Z,
                    int64_t ldZ, cudaStream_t s);

// K9  ruvIII.R:148-149  apply.average.rep: out[g, ] = mean of the rows of newY in replicate group g (all p groups).
void launch_group_mean(RowSrc src, const int* all_start, const int* all_rows, int n_groups, int n, double* out, int64_t ld,
                       cudaStream_t s);

// Transpose (gene-side Gram: t(Z) with the rows of Z contiguous) and alpha = diag(sigma) V' (appendix A-5).
void launch_transpose(const double* in, int64_t ld_in, int rows, int cols, double* out, int64_t ld_out, cudaStream_t s);
void launch_alpha_from_v(const double* V_rowmajor, const double* evals, int n, int kk, int kp, int64_t ld, double* alpha,
                         cudaStream_t s);

// K2  ruvIII.R:140  Y0 %*% t(Y0): symmetric FP64 Gram on DMMA, split over the contraction (genes).
struct GramPlan {
    int R = 0, K = 0;            // order of the Gram, contraction length (padded to kRowAlign)
    int bm = 128, ctas_per_sm = 1;   // output tile edge (128 or 80) and resident CTAs per SM of that configuration
    int tiles_1d = 0, n_tiles = 0, splits = 1, kchunk = 0;
    size_t workspace_bytes = 0;  // splits x n_tiles x bm x bm doubles
    int impl = 1;                // 0 = plain DFMA reference kernel, 1 = cp.async + DMMA (tile by padded work),
                                 // 2 = TMA + DMMA (128 tiles), 4 = cp.async + DMMA with 128 tiles forced
};
GramPlan plan_gram(int R, int K, int n_sms, int impl);
// G (R x R, ld = R) = sum over this shard's genes; both triangles written, bitwise symmetric.
// tile_map: device array of n_tiles int2 (ti, tj) with ti <= tj, filled by fill_gram_tile_map.
void fill_gram_tile_map(const GramPlan& gp, int2* host_map);
int launch_gram(const GramPlan& gp, const double* Y0, int64_t ldY0, const int2* tile_map, double* workspace,
                double* G, const void* tma_desc, cudaStream_t s);   // returns kernels launched

// K3 post-processing: cuSOLVER returns ascending eigenpairs; keep the top `n_keep` in descending order with
// a deterministic sign (largest-|component| positive): Vd column-major R x n_keep (optional, may be NULL),
// U row-major R x kp holding the first kk components (zero padded) for the projection kernel.
void launch_pick_top(const double* evec_colmajor, int64_t ldv, const double* eval_asc, int n_found, int R,
                     int n_keep, int kk, int kp, double* Vd_colmajor, double* U_rowmajor, double* eval_desc,
                     cudaStream_t s);

// K3  ruvIII.R:140  svd(Y0 %*% t(Y0))$u[, 1:k]: block subspace iteration + Rayleigh-Ritz (eigen_topk.cu).
struct TopkWork {
    double *Q, *X, *Z, *Zp, *Sp, *Hp, *Rp, *M, *V, *theta, *resid, *D;
    int* state;                  // device: [0] converged, [1] iterations at the last check, [2] Cholesky breakdown
};
struct TopkResult {
    bool converged = false;
    bool pending = false;        // persistent solver launched without waiting: the verdict is in h_state once the stream drained
    int iterations = 0, launches = 0;
};
constexpr int kTopkBlock = 32;   // block width B of the subspace iteration
// Largest k the iteration is tried for: convergence of pair j goes with lambda_{B+1} / lambda_j, so what matters is a gap
// below the block, not B >= 2k; beyond kTopkMaxK (or without convergence in kTopkPatience passes when k > B / 2) cuSOLVER
// takes over.
constexpr int kTopkMaxK = 28;
constexpr int kTopkPatience = 45;
size_t topk_work_doubles(int R, int B);
void topk_bind(TopkWork& w, double* base, int* state, int R, int B);
// U: row-major R x kp (first kk columns, sign fixed); evals: n_keep descending Ritz values (only the first kk converged)
TopkResult topk_eigen(const double* G, int R, int kk, int kp, int n_keep, int B, TopkWork& w, double* U, double* evals,
                      double tol, int max_iter, int n_sms, int* h_state, cudaStream_t s);

// K3, persistent form (eigen_persist.cu): the same iteration as ONE cooperative kernel for Gram orders that fit a
// co-resident grid (R <= 32 x #SMs), convergence decided on the device, Chebyshev-filtered continuation for slowly decaying
// spectra.  topk_persist_rows: rows of G per CTA (8/16/32) or 0 when the order is out of range (or RUVIII_EIG_PERSIST=0).
// persist_ws: topk_persist_doubles(n_sms) doubles; w.state must hold 16 ints (zeroed once).  flag_out (optional): receives
// 0.0 / 1.0 (converged / not) so that the next allreduce tells every rank whether any rank has to fall back.
int topk_persist_rows(int R, int n_sms);
size_t topk_persist_doubles(int n_sms);
TopkResult topk_eigen_persistent(const double* G, int R, int kk, int kp, int n_keep, TopkWork& w, double* persist_ws, double* U,
                                 double* evals, double* flag_out, double tol, int max_iter, int n_sms, int* h_state, bool sync,
                                 cudaStream_t s);

// K4  ruvIII.R:140  fullalpha = t(U) %*% Y  (rows 1..kk; = t(U) %*% Y0 because U is orthogonal to col(M)).
// part: rsplit x kp x ld partial sums over row chunks of Y0 (summed by launch_alpha_finalize).
int  project_rsplit(int R, int n, int n_sms);
// k_used (optional): columns of U beyond it are zero padding and are skipped (their rows of `part` are written as zeros).
void launch_project(const double* Y0, int64_t ldY0, int R, int n, const double* U_rowmajor, int kp,
                    int rsplit, double* part, int64_t ld, cudaStream_t s, int k_used = 0);
// alpha[j][g] = sum_split part; acT[c][j] = alpha[j][ctl_idx[c]]   (ruvIII.R:142-143)
void launch_alpha_finalize(const double* part, int rsplit, int kp, int n, int64_t ld, double* alpha,
                           const int* ctl_idx, int n_ctl, double* acT, cudaStream_t s);

// K5  ruvIII.R:144  A = Y[, ctl] %*% t(ac)  (rows_w x kp, row-major) and Gm = ac %*% t(ac) (kp x kp).
// The centring Y.stand = Y - 1 mu' (ruvIII.R:118-120) is applied later as A - 1 colMeans(A) (exact algebra).
// Yc (optional): the control columns gathered beforehand by launch_ctl_gather (rows_w x ldc), read contiguously.
void launch_ctl_gather(RowSrc src, int rows_w, const int* ctl_idx, int n_ctl, double* Yc, int64_t ldc, int log_rows,
                       double pseudo_count, cudaStream_t s);
void launch_ctl_products(RowSrc src, int rows_w, const int* ctl_idx, int n_ctl, const double* acT, int kp,
                         double* A, double* Gm, const double* Yc, int64_t ldc, cudaStream_t s);   // Gm == NULL: A only
// Gm = ac %*% t(ac) alone (one latency-bound CTA): the engine runs it beside the products on the auxiliary stream
void launch_ac_gram(const double* acT, int n_ctl, int kp, double* Gm, cudaStream_t s);

// K5c ruvIII.R:144  solve(ac %*% t(ac)) via Cholesky Gm[:kk,:kk] = L L'.  Produces Linv (kp x kp, lower), the
// column means of A over the first mu_rows rows, and a status word (nonzero = not positive definite).
void launch_chol_solve(const double* A, int rows_w, int mu_rows, const double* Gm, int kk, int kp,
                       double* Linv, double* colmean, int* status, cudaStream_t s);
// At[i][q] = sum_{j<=q} (A[i][j] - colmean[j]) Linv[q][j]    (A~ = A_c L^-T)
void launch_atilde(const double* A, int rows_w, const double* colmean, const double* Linv, int kk, int kp,
                   double* At, cudaStream_t s);
// at[q][g] = sum_{j<=q} Linv[q][j] alpha[j][g]               (a~ = L^-1 alpha)
void launch_alpha_tilde(const double* alpha, const double* Linv, int kk, int kp, int n, int64_t ld,
                        double* at, cudaStream_t s);
// The three steps above in ONE cooperative launch (solve.cu): every CTA sums a slice of the rows, one grid barrier, every CTA
// factors redundantly and writes its rows of A~ / its genes of a~.  ws: solve_fused_ws_doubles(n_sms) doubles, zeroed once.
size_t solve_fused_ws_doubles(int n_sms);
void launch_solve_fused(const double* A, int rows_w, int mu_rows, const double* Gm, int kk, int kp, const double* alpha, int n,
                        int64_t ld, double* ws, double* Linv, double* colmean, int* status, double* At, double* at, int n_sms,
                        cudaStream_t s);
// W_k = A~[:, :k] Linv[:k, :k]  (rows_w x k, row-major, dense)  -- the W of ruvIII.R:144 for one k
void launch_w_from_atilde(const double* At, int rows_w, const double* Linv, int k, int kp, double* W,
                          cudaStream_t s);

// K6  ruvIII.R:145  newY = Y - W %*% alpha, for every requested k in one read of Y:
// newY_k = Y - sum_{q<k} A~[:, q] (x) a~[q, :]   (nested Cholesky form).
// kmask bit (k) set (k = 0..kk) => block out_index(k) of `out` receives newY_k; blocks are out_stride apart.
void launch_update(const double* Yrows, int64_t ldY, int rows, int n, const double* At_rows, int kp,
                   const double* at, int64_t ldat, int kk, uint64_t kmask, const int* out_slot_of_k,
                   double* out, int64_t ldout, int64_t out_stride, cudaStream_t s);

// K6 impl 3 (update.cu): a~ in shared memory, ~64 registers, 8 CTAs per SM, software-pipelined row loads.
void launch_update_smem(const double* Yrows, int64_t ldY, int rows, int n, const double* At_rows, int kp,
                        const double* at, int64_t ldat, int kk, uint64_t kmask, const int* out_slot_of_k,
                        double* out, int64_t ldout, int64_t out_stride, cudaStream_t s);
// K6 impl 2 (update_tma.cu): same update, Y staged through a TMA + mbarrier shared-memory ring.
void update_tma_prepare(const double* Y, int64_t ldY, int rows, void* desc_out_host128);
void launch_update_tma(const void* tmap_host, int rows, int n, const double* At_rows, int kp, const double* at,
                       int64_t ldat, int kk, uint64_t kmask, const int* out_slot_of_k, double* out, int64_t ldout,
                       int64_t out_stride, cudaStream_t s);

// K10 computePCA.R:120-134  centre (and scale) the columns of a samples x genes matrix; v = t(alpha) / d  (pca.cu)
int colstat_blocks(int rows, int n, int n_sms);
int launch_center_columns(const double* X, int64_t ld, int rows, int n, int center, int scale, int blocks, double* part,
                          double* mean, double* inv_sd, double* Xc, int64_t ldc, cudaStream_t s);   // returns kernels launched
void launch_v_from_alpha(double* alpha, const double* evals, int kk, int64_t ld, double* d_out, cudaStream_t s);

// K11 supervisedFindNCG.R:179-191,214-229  per-gene one-way ANOVA F and Pearson correlation in one streaming pass (stats.cu)
int gene_stats_piece_rows(int n, int n_rows, int n_sms);
void launch_gene_stats(const double* X, int64_t ld, int n, const int* piece_start, const int* piece_rows, const int* piece_group,
                       int n_pieces, int n_groups, int n_rows, const double* cov, double cov_ss, int log_first, double pseudo_count,
                       double* part, int64_t ldp, double* F, double* r, cudaStream_t s);

// K11b supervisedFindNCG.R:214-229 with corr.method = "spearman": per-gene Spearman rho (spearman.cu); m1 <= spearman_max_rows()
int spearman_max_rows();
void launch_gene_spearman(const double* X, int64_t ld, int m1, int n, const double* crank, double cmean, double css, double* rho,
                          cudaStream_t s);

// K12 supervisedFindNCG.R:111-133  orthonormal basis of an m1 x p model matrix (regress.cu): W = p x m1 scratch,
// Q = ceil(p / 32) chunks, chunk c a row-major m1 x pad_k(columns) block at Q + c * m1 * 32; aliased columns are zero.
void launch_design_qr(const double* D, int64_t ldd, int m1, int p, double tol, double* W, double* Q, int* rank_out, cudaStream_t s);

// micro-benchmarks (ceilings quoted beside the roofline)
float probe_copy_gbs(size_t bytes, cudaStream_t s);
float probe_dmma_tflops(int n_sms, cudaStream_t s);
float probe_host_gather_ms(int rows, int n, int n_idx, cudaStream_t s);
void probe_copy_variants(double* results8, cudaStream_t s);

}  // namespace ruviii

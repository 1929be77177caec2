// K11: per-gene one-way ANOVA F statistic and Pearson correlation against a sample covariate -- the arithmetic of
// negative-control-gene selection.
//
// Reference: supervisedFindNCG.R:179-191,326-341 `matrixTests::row_oneway_equalvar(x, g)$statistic` (classical one-way
// ANOVA, equal variances) and :214-229,270-285 `Rfast::correls(y, x, type = "pearson")[, "correlation"]`; neither package
// is vendored, so both are restated from their documented definitions:
//     F = [sum_g n_g (mean_g - mean)^2 / (G - 1)] / [sum_g sum_{i in g} (x_i - mean_g)^2 / (N - G)]
//     r = sum (x_i - mean_x)(y_i - mean_y) / sqrt(sum (x_i - mean_x)^2 sum (y_i - mean_y)^2)
//
// One streaming pass over the samples x genes matrix, HBM-bound (8 B per cell).  A thread owns two adjacent genes; the
// samples arrive as CSR "pieces": every ANOVA group is cut into pieces of at most a few hundred rows so that grid.y
// (one piece per CTA row) fills the GPU whatever the group sizes are (5 biologies or 73 batches), and a piece never
// straddles two groups.  Each (piece, gene) emits three additive sums
//     S1 = sum x',   S2 = sum x'^2,   C = sum x' c          (x' = x - x0, c = centred covariate)
// and the finalize kernel walks the pieces in order, closing a group (A += S_g^2 / n_g) whenever the group id changes --
// fixed summation order, no atomics.  The shift x0 (the gene's value in the first listed sample) removes the
// cancellation a raw one-pass sum of squares would suffer on log counts (mean ~8, sd ~1).
#include "kernels.cuh"

#include <algorithm>

namespace ruviii {

namespace {

constexpr int kStatThreads = 128;

__global__ void __launch_bounds__(kStatThreads)
gene_stats_kernel(const double* __restrict__ X, int64_t ld, int n_pairs, const int* __restrict__ piece_start,
                  const int* __restrict__ piece_rows, const double* __restrict__ cov, int log_first, double pseudo_count,
                  double* __restrict__ part, int64_t ldp) {
    const int pair = blockIdx.x * kStatThreads + threadIdx.x;
    if (pair >= n_pairs) return;
    const int64_t col = 2 * (int64_t)pair;
    auto load = [&](int row) {
        double2 v = ld_stream(X + (int64_t)row * ld + col);
        if (log_first) { v.x = log2(v.x + pseudo_count); v.y = log2(v.y + pseudo_count); }
        return v;
    };
    const double2 x0 = load(piece_rows[0]);
    double2 S1 = make_double2(0.0, 0.0), S2 = S1, C = S1;
    const int s = piece_start[blockIdx.y], e = piece_start[blockIdx.y + 1];
    int t = s;
    for (; t + 4 <= e; t += 4) {                              // four independent 16-byte loads in flight per thread
        int r[4];
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) r[u] = piece_rows[t + u];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = load(r[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double c = cov ? cov[r[u]] : 0.0;
            const double dx = v[u].x - x0.x, dy = v[u].y - x0.y;
            S1.x += dx; S1.y += dy;
            S2.x = fma(dx, dx, S2.x); S2.y = fma(dy, dy, S2.y);
            C.x = fma(dx, c, C.x); C.y = fma(dy, c, C.y);
        }
    }
    for (; t < e; ++t) {
        const int r = piece_rows[t];
        const double2 v = load(r);
        const double c = cov ? cov[r] : 0.0;
        const double dx = v.x - x0.x, dy = v.y - x0.y;
        S1.x += dx; S1.y += dy;
        S2.x = fma(dx, dx, S2.x); S2.y = fma(dy, dy, S2.y);
        C.x = fma(dx, c, C.x); C.y = fma(dy, c, C.y);
    }
    double* p = part + (int64_t)blockIdx.y * 3 * ldp + col;
    *reinterpret_cast<double2*>(p) = S1;
    *reinterpret_cast<double2*>(p + ldp) = S2;
    *reinterpret_cast<double2*>(p + 2 * ldp) = C;
}

__global__ void __launch_bounds__(256)
gene_stats_finalize_kernel(const double* __restrict__ part, int n_pieces, const int* __restrict__ piece_start,
                           const int* __restrict__ piece_group, int64_t ldp, int n, int n_rows, int n_groups, double cov_ss,
                           double* __restrict__ F, double* __restrict__ r) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    double A = 0.0, S1 = 0.0, S2 = 0.0, C = 0.0, sg = 0.0;
    int ng = 0;
    for (int b = 0; b < n_pieces; ++b) {
        const double* p = part + (int64_t)b * 3 * ldp + g;
        if (b > 0 && piece_group[b] != piece_group[b - 1]) {      // the previous group is complete
            A = fma(sg / (double)ng, sg, A);
            sg = 0.0;
            ng = 0;
        }
        sg += p[0];
        ng += piece_start[b + 1] - piece_start[b];
        S1 += p[0]; S2 += p[ldp]; C += p[2 * ldp];
    }
    A = fma(sg / (double)ng, sg, A);
    const double N = (double)n_rows, mean_sq = S1 * S1 / N;
    if (F) {
        const double ssb = A - mean_sq, ssw = S2 - A;
        F[g] = (ssb / (double)(n_groups - 1)) / (ssw / (N - (double)n_groups));
    }
    if (r) r[g] = C / sqrt((S2 - mean_sq) * cov_ss);             // sum c = 0, so sum x' c = sum (x - mean) c
}

}  // namespace

// rows per piece such that grid.x * n_pieces is about two waves of resident CTAs
int gene_stats_piece_rows(int n, int n_rows, int n_sms) {
    const int gx = ceil_div((n + 1) / 2, kStatThreads);
    const int target = std::max(1, ceil_div(n_sms * 16 * 2, gx));
    return std::max(32, ceil_div(n_rows, target));
}

// piece_start / piece_rows: CSR of the pieces (device); piece_group: group id of every piece (device, pieces of one group
// adjacent); cov: device, one centred covariate value per sample row of X (NULL: no correlation); part: n_pieces x 3 x ldp.
void launch_gene_stats(const double* X, int64_t ld, int n, const int* piece_start, const int* piece_rows, const int* piece_group,
                       int n_pieces, int n_groups, int n_rows, const double* cov, double cov_ss, int log_first, double pseudo_count,
                       double* part, int64_t ldp, double* F, double* r, cudaStream_t s) {
    const int n_pairs = (n + 1) / 2;
    dim3 grid(ceil_div(n_pairs, kStatThreads), n_pieces);
    gene_stats_kernel<<<grid, kStatThreads, 0, s>>>(X, ld, n_pairs, piece_start, piece_rows, cov, log_first, pseudo_count, part, ldp);
    RUVIII_LAUNCH_CHECK();
    gene_stats_finalize_kernel<<<ceil_div(n, 256), 256, 0, s>>>(part, n_pieces, piece_start, piece_group, ldp, n, n_rows, n_groups,
                                                              cov_ss, F, r);
    RUVIII_LAUNCH_CHECK();
}

}  // namespace ruviii

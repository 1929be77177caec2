// K12  supervisedFindNCG.R:111-133  lm(t(expr.data) ~ covariates)$residuals: an orthonormal basis of the model matrix.
//
// The residuals of a least-squares fit depend on the column space of the design only: with Q an orthonormal basis of
// span(D), residuals = Y - Q (Q' Y).  Q' Y is K4 (project_kernel) and Y - Q (Q' Y) is K6 (the update kernel), so the one new
// piece is the factorisation of the m1 x p design (p = a few tens of columns: intercept, factor dummies, covariates).
// One CTA: classical Gram-Schmidt with re-orthogonalisation ("twice is enough"), column by column.  A column whose norm
// falls to tol x its original norm is aliased and dropped (zeroed) -- the rule of R's dqrdc2 (lm's QR, tol = 1e-7), so
// a rank-deficient design (nested factors) gives the same fitted space as lm.  Every reduction has a fixed order.
//
// Output: chunks of at most 32 columns, chunk c a row-major m1 x pad_k(columns in chunk) block at Q + c * m1 * 32 -- the
// layout K4 (U, row-major R x kp) and K6 (A~, rows x kp) read.
#include "kernels.cuh"

namespace ruviii {

namespace {

constexpr int QT = 1024;

__device__ double block_sum(double v, double* red) {          // fixed-order sum over the CTA; every thread gets the total
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < QT / 32; ++w) t += red[w];
    return t;
}

__global__ void __launch_bounds__(QT)
design_qr_kernel(const double* __restrict__ D, int64_t ldd, int m1, int p, double tol, double* __restrict__ W /* p x m1, column j at W + j*m1 */,
                 double* __restrict__ Q, int* __restrict__ rank_out) {
    __shared__ double red[QT / 32];
    __shared__ double c[256];
    __shared__ int kept[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int rank = 0;
    for (int j = 0; j < p; ++j) {
        double* wj = W + (int64_t)j * m1;
        double s0 = 0.0;
        for (int i = tid; i < m1; i += QT) { const double v = D[(int64_t)i * ldd + j]; wj[i] = v; s0 += v * v; }
        const double norm0 = sqrt(block_sum(s0, red));
        for (int pass = 0; pass < 2; ++pass) {
            for (int i = warp; i < j; i += QT / 32) {          // c_i = <q_i, w_j>, one warp per earlier column
                double s = 0.0;
                if (kept[i]) {
                    const double* wi = W + (int64_t)i * m1;
                    for (int r = lane; r < m1; r += 32) s = fma(wi[r], wj[r], s);
                    s = warp_sum(s);
                }
                if (lane == 0) c[i] = s;
            }
            __syncthreads();
            for (int r = tid; r < m1; r += QT) {
                double v = wj[r];
                for (int i = 0; i < j; ++i)
                    if (kept[i]) v = fma(-c[i], W[(int64_t)i * m1 + r], v);
                wj[r] = v;
            }
            __syncthreads();
        }
        double s1 = 0.0;
        for (int i = tid; i < m1; i += QT) s1 += wj[i] * wj[i];
        const double norm1 = sqrt(block_sum(s1, red));
        const bool keep = norm0 > 0.0 && norm1 > tol * norm0;
        const double sc = keep ? 1.0 / norm1 : 0.0;
        for (int i = tid; i < m1; i += QT) wj[i] *= sc;
        if (tid == 0) kept[j] = keep ? 1 : 0;
        rank += keep ? 1 : 0;
        __syncthreads();
    }
    // row-major chunks for K4 / K6
    const int n_chunks = (p + 31) / 32;
    for (int ch = 0; ch < n_chunks; ++ch) {
        const int kc = min(32, p - 32 * ch);
        const int kp = kc <= 4 ? 4 : kc <= 8 ? 8 : kc <= 16 ? 16 : 32;
        double* q = Q + (int64_t)ch * m1 * 32;
        for (int64_t e = tid; e < (int64_t)m1 * kp; e += QT) {
            const int r = (int)(e / kp), jj = (int)(e % kp);
            q[e] = jj < kc ? W[(int64_t)(32 * ch + jj) * m1 + r] : 0.0;
        }
    }
    if (tid == 0) *rank_out = rank;
}

}  // namespace

// D: m1 x p row-major (pitch ldd) on the device; W: p x m1 doubles of scratch; Q: ceil(p / 32) x m1 x 32 doubles.
void launch_design_qr(const double* D, int64_t ldd, int m1, int p, double tol, double* W, double* Q, int* rank_out, cudaStream_t s) {
    design_qr_kernel<<<1, QT, 0, s>>>(D, ldd, m1, p, tol, W, Q, rank_out);
    RUVIII_LAUNCH_CHECK();
}

}  // namespace ruviii

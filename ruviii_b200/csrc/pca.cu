// K10: column centring / scaling for computePCA's fast path.
//
// Reference: computePCA.R:120-134 `runSVD(x = transpose(temp.data), k = nb.pcs, center = center, scale = scale)`,
// i.e. the top nb.pcs singular triplets of the samples x genes matrix after subtracting every gene's mean over the
// samples (and dividing by its standard deviation when scale = TRUE, as base::scale does: sqrt(sum((x - mean)^2) /
// (m - 1))).  The decomposition itself reuses K2 (Gram on DMMA), K3 (top-k eigensolver) and K4 (projection); what is
// new is this HBM-bound pair of passes over the matrix: column sums, then centre (+ sums of squares) in one
// read-modify-write.  Row blocks give the grid its parallelism; partial sums are combined in a fixed order.
#include "kernels.cuh"

#include <algorithm>

namespace ruviii {

namespace {

constexpr int kStatThreads = 128;

// part[by][g] = sum over the rows of block by of X[r][g]   (thread = two adjacent genes)
__global__ void __launch_bounds__(kStatThreads)
colsum_kernel(const double* __restrict__ X, int64_t ld, int rows, int n_pairs, int rows_per_block, double* __restrict__ part) {
    const int pair = blockIdx.x * kStatThreads + threadIdx.x;
    if (pair >= n_pairs) return;
    const int64_t col = 2 * (int64_t)pair;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(r0 + rows_per_block, rows);
    double2 s = make_double2(0.0, 0.0);
    int r = r0;
    for (; r + 4 <= r1; r += 4) {                       // four independent 16-byte loads in flight per thread
        const double2 a = ld_stream(X + (int64_t)r * ld + col), b = ld_stream(X + (int64_t)(r + 1) * ld + col);
        const double2 c = ld_stream(X + (int64_t)(r + 2) * ld + col), d = ld_stream(X + (int64_t)(r + 3) * ld + col);
        s.x += (a.x + b.x) + (c.x + d.x);
        s.y += (a.y + b.y) + (c.y + d.y);
    }
    for (; r < r1; ++r) {
        const double2 a = ld_stream(X + (int64_t)r * ld + col);
        s.x += a.x;
        s.y += a.y;
    }
    *reinterpret_cast<double2*>(part + (int64_t)blockIdx.y * ld + col) = s;
}

// out[g] = f(sum_by part[by][g]):  mode 0: mean = sum / rows;  mode 1: 1 / sd = rsqrt(sum / (rows - 1))  (0 -> 1, as
// dividing by a zero sd would give NaN columns; base::scale does produce NaN there -- see DESIGN.md)
__global__ void __launch_bounds__(256)
colstat_finalize_kernel(const double* __restrict__ part, int nblocks, int64_t ld, int rows, int mode, double* __restrict__ out) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ld) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += part[(int64_t)b * ld + g];
    if (mode == 0) out[g] = s / (double)rows;
    else out[g] = s > 0.0 ? 1.0 / sqrt(s / (double)(rows - 1)) : 1.0;
}

// Xc = (log2(X + pc))? - mean, optionally accumulating the per-block sums of squares of the centred values
__global__ void __launch_bounds__(kStatThreads)
center_kernel(const double* __restrict__ X, int64_t ld, int rows, int n_pairs, int rows_per_block,
              const double* __restrict__ mean, double* __restrict__ Xc, int64_t ldc, double* __restrict__ sq_part) {
    const int pair = blockIdx.x * kStatThreads + threadIdx.x;
    if (pair >= n_pairs) return;
    const int64_t col = 2 * (int64_t)pair;
    const double2 mu = *reinterpret_cast<const double2*>(mean + col);
    const int r0 = blockIdx.y * rows_per_block, r1 = min(r0 + rows_per_block, rows);
    double2 q = make_double2(0.0, 0.0);
    int r = r0;
    for (; r + 4 <= r1; r += 4) {
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ld_stream(X + (int64_t)(r + u) * ld + col);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            v[u].x -= mu.x;
            v[u].y -= mu.y;
            q.x = fma(v[u].x, v[u].x, q.x);
            q.y = fma(v[u].y, v[u].y, q.y);
            *reinterpret_cast<double2*>(Xc + (int64_t)(r + u) * ldc + col) = v[u];     // re-read by the Gram: keep in L2
        }
    }
    for (; r < r1; ++r) {
        double2 v = ld_stream(X + (int64_t)r * ld + col);
        v.x -= mu.x;
        v.y -= mu.y;
        q.x = fma(v.x, v.x, q.x);
        q.y = fma(v.y, v.y, q.y);
        *reinterpret_cast<double2*>(Xc + (int64_t)r * ldc + col) = v;
    }
    if (sq_part) *reinterpret_cast<double2*>(sq_part + (int64_t)blockIdx.y * ldc + col) = q;
}

__global__ void __launch_bounds__(kStatThreads)
colscale_kernel(double* __restrict__ Xc, int64_t ld, int rows, int n_pairs, int rows_per_block, const double* __restrict__ inv_sd) {
    const int pair = blockIdx.x * kStatThreads + threadIdx.x;
    if (pair >= n_pairs) return;
    const int64_t col = 2 * (int64_t)pair;
    const double2 w = *reinterpret_cast<const double2*>(inv_sd + col);
    const int r0 = blockIdx.y * rows_per_block, r1 = min(r0 + rows_per_block, rows);
    for (int r = r0; r < r1; ++r) {
        double2* p = reinterpret_cast<double2*>(Xc + (int64_t)r * ld + col);
        double2 v = *p;
        v.x *= w.x;
        v.y *= w.y;
        *p = v;
    }
}

// v[j][g] = alpha[j][g] / d[j]   (right singular vectors from the projection t(U) Xc = diag(d) t(V)); d[j] = sqrt(lambda_j)
__global__ void __launch_bounds__(256)
v_from_alpha_kernel(double* __restrict__ alpha, const double* __restrict__ evals, int kk, int64_t ld, double* __restrict__ d_out) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (j >= kk) return;
    const double d = sqrt(fmax(evals[j], 0.0));
    if (g == 0 && d_out) d_out[j] = d;
    if (g < ld) alpha[(int64_t)j * ld + g] = d > 0.0 ? alpha[(int64_t)j * ld + g] / d : 0.0;
}

}  // namespace

int colstat_blocks(int rows, int n, int n_sms) {
    const int gx = ceil_div((n + 1) / 2, kStatThreads);
    return std::max(1, std::min(ceil_div(rows, 32), ceil_div(n_sms * 16 * 2, gx)));
}

// mean (ld doubles), part (blocks x ld doubles) are workspaces; Xc may alias X (in-place centring)
int launch_center_columns(const double* X, int64_t ld, int rows, int n, int center, int scale, int blocks, double* part,
                          double* mean, double* inv_sd, double* Xc, int64_t ldc, cudaStream_t s) {
    const int n_pairs = (n + 1) / 2;
    const int rpb = ceil_div(rows, blocks);
    const int nb = ceil_div(rows, rpb);
    dim3 grid(ceil_div(n_pairs, kStatThreads), nb);
    int launches = 0;
    if (center) {
        colsum_kernel<<<grid, kStatThreads, 0, s>>>(X, ld, rows, n_pairs, rpb, part);
        colstat_finalize_kernel<<<ceil_div(ld, 256), 256, 0, s>>>(part, nb, ld, rows, 0, mean);
        launches += 2;
    } else {
        RUVIII_CUDA_CHECK(cudaMemsetAsync(mean, 0, (size_t)ld * 8, s));
    }
    center_kernel<<<grid, kStatThreads, 0, s>>>(X, ld, rows, n_pairs, rpb, mean, Xc, ldc, scale ? part : nullptr);
    launches += 1;
    if (scale) {
        colstat_finalize_kernel<<<ceil_div(ldc, 256), 256, 0, s>>>(part, nb, ldc, rows, 1, inv_sd);
        colscale_kernel<<<grid, kStatThreads, 0, s>>>(Xc, ldc, rows, n_pairs, rpb, inv_sd);
        launches += 2;
    }
    RUVIII_LAUNCH_CHECK();
    return launches;
}

void launch_v_from_alpha(double* alpha, const double* evals, int kk, int64_t ld, double* d_out, cudaStream_t s) {
    dim3 grid(ceil_div(ld, 256), kk);
    v_from_alpha_kernel<<<grid, 256, 0, s>>>(alpha, evals, kk, ld, d_out);
    RUVIII_LAUNCH_CHECK();
}

}  // namespace ruviii

// K5c + Ã + ã in ONE cooperative launch (ruvIII.R:118-120,144): column means of A = Y[, ctl] ac' over the rows that define mu,
// Cholesky of ac ac' = L L' with R's singularity test, L^-1, A~ = (A - 1 colMeans(A)) L^-T and a~ = L^-1 alpha.
//
// The three-kernel form (chol_solve + atilde + alpha_tilde, project.cu) costs ~60 us at C3 -- a single CTA sums 12,200 x kp
// doubles, then two more dependent launches -- which is 7 % of the 8-GPU pass.  Here every CTA sums a slice of the rows, one
// grid barrier later every CTA folds the partials in the same fixed order, factors the k x k matrix redundantly (one warp) and
// then produces its own rows of A~ and its own genes of a~.
#include "kernels.cuh"

namespace ruviii {

namespace {

constexpr int ST = 256;

__device__ __forceinline__ void solve_grid_barrier(unsigned* bar, unsigned P) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        unsigned v;
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while (v < P);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}

__global__ void __launch_bounds__(ST)
solve_fused_kernel(const double* __restrict__ A, int rows_w, int mu_rows, const double* __restrict__ Gm, int kk, int kp,
                   const double* __restrict__ alpha, int n, int64_t ld, double* __restrict__ part, unsigned* __restrict__ bar,
                   double* __restrict__ Linv_out, double* __restrict__ colmean_out, int* __restrict__ status,
                   double* __restrict__ At, double* __restrict__ at) {
    __shared__ double L[kMaxK][kMaxK + 1];
    __shared__ double Li[kMaxK][kMaxK + 1];
    __shared__ double red[ST];
    __shared__ double cm[kMaxK];
    __shared__ int bad;
    const int tid = threadIdx.x, cta = blockIdx.x, P = gridDim.x;
    const int rows_per = (rows_w + P - 1) / P;
    const int i0 = cta * rows_per, i1 = min(rows_w, i0 + rows_per);
    const int j = tid % kp, lr = tid / kp, stride = ST / kp;
    // ---- partial column sums of this CTA's rows (those that define mu; mu_rows = 0: no centring, RUV1) ----
    {
        double s = 0.0;
        if (lr < stride)
            for (int i = i0 + lr; i < min(i1, mu_rows); i += stride) s += A[(int64_t)i * kp + j];
        red[tid] = lr < stride ? s : 0.0;
        __syncthreads();
        if (tid < kp) {
            double t = 0.0;
            for (int q = 0; q < stride; ++q) t += red[q * kp + tid];
            __stcg(part + (int64_t)cta * kMaxK + tid, t);
        }
    }
    solve_grid_barrier(bar, (unsigned)P);
    // ---- every CTA: fold the partials (same order everywhere), factor, invert ----
    {
        double s = 0.0;
        if (lr < stride)
            for (int c = lr; c < P; c += stride) s += __ldcg(part + (int64_t)c * kMaxK + j);
        __syncthreads();
        red[tid] = lr < stride ? s : 0.0;
        __syncthreads();
        if (tid < kp) {
            double t = 0.0;
            for (int q = 0; q < stride; ++q) t += red[q * kp + tid];
            cm[tid] = mu_rows > 0 ? t / (double)mu_rows : 0.0;
        }
    }
    for (int e = tid; e < kMaxK * kMaxK; e += ST) {
        const int r = e / kMaxK, c = e % kMaxK;
        L[r][c] = (r < kk && c < kk) ? Gm[r * kp + c] : (r == c ? 1.0 : 0.0);
        Li[r][c] = 0.0;
    }
    if (tid == 0) bad = 0;
    __syncthreads();
    if (tid < 32) {
        const int i = tid;
        for (int c = 0; c < kk; ++c) {                       // right-looking Cholesky, one warp, row i per lane
            if (i == c) {
                // R's solve() (dgesv + rcond check) stops on a computationally singular matrix; an exactly singular Gm leaves
                // a pivot of a few ulps of its diagonal entry with either sign, so "not positive" alone would miss it
                const double d = L[c][c];
                if (!(d > 16.0 * 2.220446049250313e-16 * Gm[c * kp + c])) { bad = c + 1; L[c][c] = 1.0; } else L[c][c] = sqrt(d);
            }
            __syncwarp();
            if (i > c && i < kk) L[i][c] /= L[c][c];
            __syncwarp();
            if (i > c && i < kk)
                for (int q = c + 1; q <= i; ++q) L[i][q] -= L[i][c] * L[q][c];
            __syncwarp();
        }
        if (i < kk) {                                        // inverse of the factor: forward substitution, one column per lane
            for (int r = i; r < kk; ++r) {
                double s = (r == i) ? 1.0 : 0.0;
                for (int q = i; q < r; ++q) s -= L[r][q] * Li[q][i];
                Li[r][i] = s / L[r][r];
            }
        }
    }
    __syncthreads();
    if (cta == 0) {
        for (int e = tid; e < kp * kp; e += ST) {
            const int r = e / kp, c = e - r * kp;
            Linv_out[e] = (r < kk && c <= r) ? Li[r][c] : 0.0;
        }
        if (tid < kp) colmean_out[tid] = cm[tid];
        if (tid == 0) *status = bad;
    }
    // ---- A~[i][q] = sum_{j <= q} (A[i][j] - colmean[j]) L^-1[q][j] for this CTA's rows ----
    for (int e = tid; e < (i1 - i0) * kp; e += ST) {
        const int i = i0 + e / kp, q = e % kp;
        double s = 0.0;
        if (q < kk)
            for (int c = 0; c <= q; ++c) s = fma(A[(int64_t)i * kp + c] - cm[c], Li[q][c], s);
        At[(int64_t)i * kp + q] = s;
    }
    // ---- a~[q][g] = sum_{j <= q} L^-1[q][j] alpha[j][g] for this CTA's genes (the padding columns too: they are zero) ----
    const int64_t g_per = (ld + P - 1) / P;
    const int64_t g0 = (int64_t)cta * g_per, g1 = min(ld, g0 + g_per);
    const int64_t width = g1 > g0 ? g1 - g0 : 0;
    for (int64_t e = tid; e < width * kp; e += ST) {
        const int q = (int)(e / width);
        const int64_t g = g0 + (e - (int64_t)q * width);
        double s = 0.0;
        if (q < kk)
            for (int c = 0; c <= q; ++c) s = fma(Li[q][c], alpha[(int64_t)c * ld + g], s);
        at[(int64_t)q * ld + g] = s;
    }
    // the last CTA out re-arms the barrier for the next launch
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned old = atomicAdd(bar + 1, 1u);
        if (old == (unsigned)P - 1) { bar[0] = 0; bar[1] = 0; __threadfence(); }
    }
}

}  // namespace

size_t solve_fused_ws_doubles(int n_sms) { return (size_t)n_sms * kMaxK + 8; }

// part: solve_fused_ws_doubles(n_sms) doubles, the first 8 of which (zeroed once) hold the barrier counters.
void launch_solve_fused(const double* A, int rows_w, int mu_rows, const double* Gm, int kk, int kp, const double* alpha, int n,
                        int64_t ld, double* ws, double* Linv, double* colmean, int* status, double* At, double* at, int n_sms,
                        cudaStream_t s) {
    const int P = std::max(1, std::min(n_sms, std::max(ceil_div(rows_w, 64), ceil_div(ld, 2048))));
    unsigned* bar = reinterpret_cast<unsigned*>(ws);
    double* part = ws + 8;
    void* params[] = {(void*)&A, (void*)&rows_w, (void*)&mu_rows, (void*)&Gm, (void*)&kk, (void*)&kp, (void*)&alpha, (void*)&n, (void*)&ld,
                      (void*)&part, (void*)&bar, (void*)&Linv, (void*)&colmean, (void*)&status, (void*)&At, (void*)&at};
    RUVIII_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)solve_fused_kernel, dim3(P), dim3(ST), params, 0, s));
}

}  // namespace ruviii

// K11b: per-gene Spearman correlation with a sample covariate (supervisedFindNCG.R:214-229,270-285 with corr.method =
// "spearman": Rfast::correls(y, x, type = "spearman") -- Pearson's r of the ranks, ties averaged; Rfast is not vendored in the
// reference, the definition is the textbook one).
//
// One CTA per gene.  The gene's m1 values (one column of the samples x genes matrix; the 4 genes of a 32-byte sector are
// handled by CTAs that run at the same time, so the strided reads mostly hit L2) go to shared memory with their sample
// index, a bitonic network sorts the pairs, neighbouring equal values are grouped with two scans (runs can be thousands
// long: zero counts), every position gets the average rank (first + last) / 2 + 1 of its run, and the correlation with
// the covariate's ranks (computed once on the host, O(m1 log m1)) is accumulated in sorted order.  Monotone transforms
// (log2(x + pseudo.count)) do not change ranks, so none is applied.
#include "kernels.cuh"

#include <cfloat>
#include <climits>

namespace ruviii {

namespace {

constexpr int SPT = 1024;

template <class T>
__device__ __forceinline__ T block_scan_max_inclusive(T v, T* warp_tmp) {        // inclusive max-scan over the CTA's threads
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = max(v, u);
    }
    if (lane == 31) warp_tmp[warp] = v;
    __syncthreads();
    if (warp == 0) {
        T w = lane < SPT / 32 ? warp_tmp[lane] : T(INT_MIN);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const T u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w = max(w, u);
        }
        warp_tmp[lane] = w;
    }
    __syncthreads();
    if (warp > 0) v = max(v, warp_tmp[warp - 1]);
    __syncthreads();
    return v;
}

// N = padded power of two >= m1 (<= 16384); shared: double key[N]; int idx[N]; unsigned head[N / 32]
__global__ void __launch_bounds__(SPT)
spearman_kernel(const double* __restrict__ X, int64_t ld, int m1, int N, const double* __restrict__ crank, double cmean, double css,
                double* __restrict__ rho) {
    extern __shared__ __align__(16) unsigned char sp_smem[];
    double* key = reinterpret_cast<double*>(sp_smem);
    int* idx = reinterpret_cast<int*>(key + N);
    unsigned* head = reinterpret_cast<unsigned*>(idx + N);
    __shared__ int warp_tmp[32];
    __shared__ double red[2][SPT / 32];
    const int g = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < N; i += SPT) {
        key[i] = i < m1 ? X[(int64_t)i * ld + g] : DBL_MAX;
        idx[i] = i;
    }
    __syncthreads();
    // ---- bitonic sort of (key, idx), ascending ----
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < N / 2; t += SPT) {
                const int lo = 2 * t - (t & (j - 1));          // index with bit j clear
                const int hi = lo + j;
                const bool up = (lo & k) == 0;
                const double a = key[lo], b = key[hi];
                if ((a > b) == up) {
                    key[lo] = b; key[hi] = a;
                    const int ia = idx[lo]; idx[lo] = idx[hi]; idx[hi] = ia;
                }
            }
            __syncthreads();
        }
    }
    // ---- runs of equal values: head bit of every position ----
    for (int w = tid; w < N / 32; w += SPT) {
        unsigned bits = 0;
        for (int b = 0; b < 32; ++b) {
            const int p = w * 32 + b;
            if (p == 0 || key[p] != key[p - 1]) bits |= 1u << b;
        }
        head[w] = bits;
    }
    __syncthreads();
    // the key array is no longer needed: it becomes first[N] | last[N] (ints)
    int* first = reinterpret_cast<int*>(key);
    int* last = first + N;
    const int per = N / SPT;                                   // consecutive positions per thread (N >= SPT)
    {
        // first[p] = largest head position <= p  (inclusive max-scan)
        int loc[16];
        int run = INT_MIN;
        for (int u = 0; u < per; ++u) {
            const int p = tid * per + u;
            if (head[p >> 5] >> (p & 31) & 1u) run = p;
            loc[u] = run;
        }
        const int incl = block_scan_max_inclusive<int>(run, warp_tmp);
        // exclusive prefix over the earlier threads = inclusive value of thread tid - 1
        __shared__ int incl_all[SPT];
        incl_all[tid] = incl;
        __syncthreads();
        const int prev = tid > 0 ? incl_all[tid - 1] : INT_MIN;
        for (int u = 0; u < per; ++u) first[tid * per + u] = max(loc[u], prev);
        __syncthreads();
        // last[p] = (smallest head position > p) - 1: the same scan from the right on negated positions
        int locr[16];
        int runr = INT_MIN;                                     // holds -(next head position) as a max
        for (int u = per - 1; u >= 0; --u) {
            const int p = (SPT - 1 - tid) * per + u;           // thread tid walks block SPT-1-tid from its right end
            locr[u] = runr;                                     // next head strictly to the right of p inside this block
            if (head[p >> 5] >> (p & 31) & 1u) runr = -p;
        }
        const int inclr = block_scan_max_inclusive<int>(runr, warp_tmp);
        incl_all[tid] = inclr;
        __syncthreads();
        const int prevr = tid > 0 ? incl_all[tid - 1] : INT_MIN;
        for (int u = 0; u < per; ++u) {
            const int p = (SPT - 1 - tid) * per + u;
            const int nh = max(locr[u], prevr);                 // -(next head) or INT_MIN
            last[p] = (nh == INT_MIN ? N : -nh) - 1;
        }
        __syncthreads();
    }
    // ---- Pearson correlation of the average ranks with the covariate's ranks ----
    const double rmean = 0.5 * ((double)m1 + 1.0);
    double sxy = 0.0, sxx = 0.0;
    for (int p = tid; p < m1; p += SPT) {
        const int l = min(last[p], m1 - 1);                     // the padding keys sort behind every real value
        const double r = 0.5 * (double)(first[p] + l) + 1.0 - rmean;
        sxy = fma(r, crank[idx[p]] - cmean, sxy);
        sxx = fma(r, r, sxx);
    }
    sxy = warp_sum(sxy);
    sxx = warp_sum(sxx);
    if ((tid & 31) == 0) { red[0][tid >> 5] = sxy; red[1][tid >> 5] = sxx; }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < SPT / 32; ++w) { a += red[0][w]; b += red[1][w]; }
        rho[g] = a / sqrt(b * css);
    }
}

}  // namespace

int spearman_max_rows() { return 16384; }

// crank: device, the covariate's average ranks (1-based) per sample; cmean / css: their mean and centred sum of squares
void launch_gene_spearman(const double* X, int64_t ld, int m1, int n, const double* crank, double cmean, double css, double* rho,
                          cudaStream_t s) {
    int N = SPT;
    while (N < m1) N <<= 1;
    const size_t smem = (size_t)N * (sizeof(double) + sizeof(int)) + (size_t)N / 32 * sizeof(unsigned);
    RUVIII_CUDA_CHECK(cudaFuncSetAttribute(spearman_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    spearman_kernel<<<n, SPT, smem, s>>>(X, ld, m1, N, crank, cmean, css, rho);
    RUVIII_LAUNCH_CHECK();
}

}  // namespace ruviii

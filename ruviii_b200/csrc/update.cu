// K6: streaming rank-k update  newY = Y - W %*% alpha  (ruvIII.R:145, fastruvIII.R:176), all requested k in
// one read of Y (replaces the per-k loop of ruvIIIMultipleK.R:75-94).
//
// With Gm = ac ac' = L L', A~ = (Y_c - 1 mu_c') ac' L^-T and a~ = L^-1 alpha the correction for k is the
// sum of the first k outer products A~[:, q] (x) a~[q, :], so
//     newY_k = newY_{k-1} - A~[:, k] (x) a~[k, :]
// and every requested k is a snapshot of one running accumulator.
//
// HBM-bound: 8 B read + 8 B written per cell and per requested k.  One thread owns two adjacent genes
// (16-byte accesses; a warp covers 512 contiguous bytes of a row) and keeps its 2 x kk entries of a~ in
// registers; the A~ rows of the CTA's row block are broadcast from shared memory.  Four rows are in flight
// per thread.  Y and newY are touched exactly once, so they bypass L1 allocation.
#include "kernels.cuh"

#include <cstdlib>
#include <cstring>
#include <vector>

namespace ruviii {

namespace {

constexpr int kUpdThreads = 128;
constexpr int kUpdRows = 64;       // rows per CTA
constexpr int kUpdUnroll = 4;      // rows in flight per thread

struct KSlots { signed char slot[kMaxK + 1]; };

__device__ __forceinline__ double2 lds_v2(const double* p) {
    double2 v;
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}

template <int KP>
__global__ void __launch_bounds__(kUpdThreads, KP <= 8 ? 6 : KP <= 16 ? 4 : 2)
update_kernel(const double* __restrict__ Y, int64_t ldY, int rows, int n_pairs, const double* __restrict__ At,
              const double* __restrict__ at, int64_t ldat, int kk, uint64_t kmask, KSlots slots,
              double* __restrict__ out, int64_t ldout, int64_t out_stride, int rows_per_cta) {
    __shared__ __align__(16) double sAt[kUpdRows][KP];
    const int pair = blockIdx.x * kUpdThreads + threadIdx.x;
    const bool active = pair < n_pairs;
    const int64_t col = 2 * (int64_t)pair;
    double2 a[KP];
#pragma unroll
    for (int q = 0; q < KP; ++q)
        a[q] = (active && q < kk) ? *reinterpret_cast<const double2*>(at + (int64_t)q * ldat + col)
                                  : make_double2(0.0, 0.0);
    const bool want0 = kmask & 1ull;
    const int cta_row0 = blockIdx.y * rows_per_cta;
    const int cta_row1 = min(cta_row0 + rows_per_cta, rows);
    // a CTA keeps its gene pair columns (a~ in registers) and walks down its rows in blocks of kUpdRows
    for (int row0 = cta_row0; row0 < cta_row1; row0 += kUpdRows) {
        const int nrows = min(kUpdRows, cta_row1 - row0);
        __syncthreads();
        for (int e = threadIdx.x; e < kUpdRows * KP; e += kUpdThreads) {
            const int r = e / KP;
            sAt[r][e - r * KP] = r < nrows ? At[(int64_t)(row0 + r) * KP + (e - r * KP)] : 0.0;
        }
        __syncthreads();
        if (!active) continue;
        for (int rb = 0; rb < nrows; rb += kUpdUnroll) {
            double2 y[kUpdUnroll];
#pragma unroll
            for (int u = 0; u < kUpdUnroll; ++u)
                y[u] = (rb + u < nrows) ? ld_stream(Y + (int64_t)(row0 + rb + u) * ldY + col) : make_double2(0.0, 0.0);
            if (want0) {
                double* o = out + (int64_t)slots.slot[0] * out_stride + col;
#pragma unroll
                for (int u = 0; u < kUpdUnroll; ++u)
                    if (rb + u < nrows) st_stream(o + (int64_t)(row0 + rb + u) * ldout, y[u]);
            }
#pragma unroll
            for (int q = 0; q < KP; q += 2) {
                if (q < kk) {
                    // volatile shared loads pin the A~ reads next to their use: without them the compiler hoists all
                    // KP x 4 broadcasts to the top of the row step and the kernel needs > 200 registers
                    double2 w[kUpdUnroll];
#pragma unroll
                    for (int u = 0; u < kUpdUnroll; ++u) w[u] = lds_v2(&sAt[rb + u][q]);   // rb + u < kUpdRows always
#pragma unroll
                    for (int u = 0; u < kUpdUnroll; ++u) {
                        y[u].x = fma(-w[u].x, a[q].x, y[u].x);
                        y[u].y = fma(-w[u].x, a[q].y, y[u].y);
                    }
                    if ((kmask >> (q + 1)) & 1ull) {
                        double* o = out + (int64_t)slots.slot[q + 1] * out_stride + col;
#pragma unroll
                        for (int u = 0; u < kUpdUnroll; ++u)
                            if (rb + u < nrows) st_stream(o + (int64_t)(row0 + rb + u) * ldout, y[u]);
                    }
                    if (q + 1 < kk) {
#pragma unroll
                        for (int u = 0; u < kUpdUnroll; ++u) {
                            y[u].x = fma(-w[u].y, a[q + 1].x, y[u].x);
                            y[u].y = fma(-w[u].y, a[q + 1].y, y[u].y);
                        }
                        if ((kmask >> (q + 2)) & 1ull) {
                            double* o = out + (int64_t)slots.slot[q + 2] * out_stride + col;
#pragma unroll
                            for (int u = 0; u < kUpdUnroll; ++u)
                                if (rb + u < nrows) st_stream(o + (int64_t)(row0 + rb + u) * ldout, y[u]);
                        }
                    }
                }
            }
        }
    }
}

// K6 impl 3: a~ lives in shared memory instead of registers.  The register version above needs 2 x kp doubles per
// thread for a~ (128 registers at kp = 16 -> 4 CTAs = 16 warps per SM), and with so few warps the load, FP64 and
// store phases of a warp barely overlap (ncu: 57 % DRAM, long-scoreboard stalls) although a plain copy with the same
// tiling reaches the copy peak.  Here a~ (kk x 128 double2) and the A~ rows sit in shared memory, a thread needs ~64
// registers (8 CTAs = 32 warps per SM), the next four rows are loaded before the current four are touched, and one
// kernel serves every k <= 32 (runtime loop over q, no register arrays).
template <int PF>   // PF > 0: rows PF*4 .. PF*4+3 ahead are pulled into L2 with prefetch.global.L2 (costs no registers)
__global__ void __launch_bounds__(kUpdThreads, 8)
update_smem_kernel(const double* __restrict__ Y, int64_t ldY, int rows, int n_pairs, const double* __restrict__ At, int kp,
                   const double* __restrict__ at, int64_t ldat, int kk, uint64_t kmask, KSlots slots,
                   double* __restrict__ out, int64_t ldout, int64_t out_stride) {
    extern __shared__ __align__(16) unsigned char upd_smem[];
    double2* sA = reinterpret_cast<double2*>(upd_smem);                       // [kk][128]
    double* sAt = reinterpret_cast<double*>(upd_smem + (size_t)kk * kUpdThreads * 16);   // [kUpdRows][kk]
    const int tid = threadIdx.x;
    const int pair = blockIdx.x * kUpdThreads + tid;
    const bool active = pair < n_pairs;
    const int64_t col = 2 * (int64_t)pair;
    const int row0 = blockIdx.y * kUpdRows;
    const int nrows = min(kUpdRows, rows - row0);
    for (int q = 0; q < kk; ++q)
        sA[q * kUpdThreads + tid] = active ? *reinterpret_cast<const double2*>(at + (int64_t)q * ldat + col) : make_double2(0.0, 0.0);
    for (int e = tid; e < kUpdRows * kk; e += kUpdThreads) {
        const int r = e / kk, q = e - r * kk;
        sAt[e] = r < nrows ? At[(int64_t)(row0 + r) * kp + q] : 0.0;
    }
    __syncthreads();
    if (!active) return;
    const bool want0 = kmask & 1ull;
    const double* yp = Y + (int64_t)row0 * ldY + col;
    double2 nxt[kUpdUnroll];
#pragma unroll
    for (int u = 0; u < kUpdUnroll; ++u) nxt[u] = u < nrows ? ld_stream(yp + (int64_t)u * ldY) : make_double2(0.0, 0.0);
    for (int rb = 0; rb < nrows; rb += kUpdUnroll) {
        double2 y[kUpdUnroll];
#pragma unroll
        for (int u = 0; u < kUpdUnroll; ++u) y[u] = nxt[u];
#pragma unroll
        for (int u = 0; u < kUpdUnroll; ++u)                                  // next four rows are in flight during the math
            nxt[u] = (rb + kUpdUnroll + u < nrows) ? ld_stream(yp + (int64_t)(rb + kUpdUnroll + u) * ldY) : make_double2(0.0, 0.0);
        if (PF > 0) {
#pragma unroll
            for (int u = 0; u < kUpdUnroll; ++u)
                if (rb + (PF + 1) * kUpdUnroll + u < nrows)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(yp + (int64_t)(rb + (PF + 1) * kUpdUnroll + u) * ldY));
        }
        if (want0) {
            double* o = out + (int64_t)slots.slot[0] * out_stride + col;
#pragma unroll
            for (int u = 0; u < kUpdUnroll; ++u)
                if (rb + u < nrows) st_stream(o + (int64_t)(row0 + rb + u) * ldout, y[u]);
        }
        const double* w = sAt + rb * kk;
#pragma unroll 2
        for (int q = 0; q < kk; ++q) {
            const double2 a = sA[q * kUpdThreads + tid];
#pragma unroll
            for (int u = 0; u < kUpdUnroll; ++u) {
                const double wq = w[u * kk + q];                               // rows beyond nrows are zero padded
                y[u].x = fma(-wq, a.x, y[u].x);
                y[u].y = fma(-wq, a.y, y[u].y);
            }
            if ((kmask >> (q + 1)) & 1ull) {
                double* o = out + (int64_t)slots.slot[q + 1] * out_stride + col;
#pragma unroll
                for (int u = 0; u < kUpdUnroll; ++u)
                    if (rb + u < nrows) st_stream(o + (int64_t)(row0 + rb + u) * ldout, y[u]);
            }
        }
    }
}

// ---- micro-benchmarks -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) copy_kernel(const double2* __restrict__ a, double2* __restrict__ b, size_t n2) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
        b[i] = a[i];
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

}  // namespace

void launch_update_smem(const double* Yrows, int64_t ldY, int rows, int n, const double* At_rows, int kp,
                        const double* at, int64_t ldat, int kk, uint64_t kmask, const int* out_slot_of_k, double* out,
                        int64_t ldout, int64_t out_stride, cudaStream_t s) {
    if (rows <= 0 || n <= 0) return;
    KSlots slots;
    for (int i = 0; i <= kMaxK; ++i) slots.slot[i] = (signed char)out_slot_of_k[i];
    const int n_pairs = (n + 1) / 2;
    const size_t smem = (size_t)std::max(kk, 1) * (kUpdThreads * 16 + kUpdRows * 8);
    static const int pf = [] { const char* v = std::getenv("RUVIII_UPD_PREFETCH"); return v && *v ? std::atoi(v) : 2; }();
    dim3 grid(ceil_div(n_pairs, kUpdThreads), ceil_div(rows, kUpdRows));
#define RUVIII_UPDS(PF)                                                                                                   \
    do {                                                                                                                  \
        RUVIII_CUDA_CHECK(cudaFuncSetAttribute(update_smem_kernel<PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        update_smem_kernel<PF><<<grid, kUpdThreads, smem, s>>>(Yrows, ldY, rows, n_pairs, At_rows, kp, at, ldat, kk, kmask, \
                                                               slots, out, ldout, out_stride);                            \
    } while (0)
    if (pf <= 0) RUVIII_UPDS(0);
    else if (pf == 1) RUVIII_UPDS(1);
    else if (pf == 2) RUVIII_UPDS(2);
    else RUVIII_UPDS(4);
#undef RUVIII_UPDS
    RUVIII_LAUNCH_CHECK();
}

void launch_update(const double* Yrows, int64_t ldY, int rows, int n, const double* At_rows, int kp,
                   const double* at, int64_t ldat, int kk, uint64_t kmask, const int* out_slot_of_k, double* out,
                   int64_t ldout, int64_t out_stride, cudaStream_t s) {
    if (rows <= 0 || n <= 0) return;
    KSlots slots;
    for (int i = 0; i <= kMaxK; ++i) slots.slot[i] = (signed char)out_slot_of_k[i];
    const int n_pairs = (n + 1) / 2;
    // rows per CTA: 64.  Long column strips (one wave of persistent CTAs) were measured slower (r01e: 0.774 vs 0.730 ms
    // at C3; a pure copy with 1,600-row strips drops from 6.4 to 4.8 TB/s); RUVIII_UPD_ROWS overrides for experiments.
    const int col_tiles = ceil_div(n_pairs, kUpdThreads);
    static const int env_rows = [] { const char* v = std::getenv("RUVIII_UPD_ROWS"); return v && *v ? std::atoi(v) : 0; }();
    int rows_per_cta = kUpdRows;
    if (env_rows > 0) rows_per_cta = (int)round_up(env_rows, kUpdRows);
    dim3 grid(col_tiles, ceil_div(rows, rows_per_cta));
#define RUVIII_UPD(KP)                                                                                         \
    update_kernel<KP><<<grid, kUpdThreads, 0, s>>>(Yrows, ldY, rows, n_pairs, At_rows, at, ldat, kk, kmask,    \
                                                   slots, out, ldout, out_stride, rows_per_cta)
    switch (kp) {
        case 4: RUVIII_UPD(4); break;
        case 8: RUVIII_UPD(8); break;
        case 16: RUVIII_UPD(16); break;
        default: RUVIII_UPD(32); break;
    }
#undef RUVIII_UPD
    RUVIII_LAUNCH_CHECK();
}

// gather of m1 x n_ctl scattered doubles straight out of pinned HOST memory (zero-copy over PCIe): how fast could
// the control-gene columns be fetched ahead of the bulk upload?  Returns milliseconds.
__global__ void __launch_bounds__(256)
host_gather_kernel(const double* __restrict__ Yh, int64_t ld, int rows, const int* __restrict__ idx, int n_idx,
                   double* __restrict__ out) {
    const int64_t total = (int64_t)rows * n_idx;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / n_idx;
        const int c = (int)(e - r * n_idx);
        out[e] = Yh[r * ld + idx[c]];
    }
}

float probe_host_gather_ms(int rows, int n, int n_idx, cudaStream_t s) {
    double* Yh = nullptr;
    RUVIII_CUDA_CHECK(cudaHostAlloc((void**)&Yh, (size_t)rows * n * 8, cudaHostAllocDefault));
    std::memset(Yh, 0, (size_t)rows * n * 8);
    std::vector<int> idx(n_idx);
    for (int i = 0; i < n_idx; ++i) idx[i] = (int)((int64_t)i * n / n_idx) + (i * 7) % 5;
    int* didx = nullptr;
    double* out = nullptr;
    RUVIII_CUDA_CHECK(cudaMalloc(&didx, n_idx * sizeof(int)));
    RUVIII_CUDA_CHECK(cudaMalloc(&out, (size_t)rows * n_idx * 8));
    RUVIII_CUDA_CHECK(cudaMemcpy(didx, idx.data(), n_idx * sizeof(int), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 3; ++it) {
        cudaEventRecord(e0, s);
        host_gather_kernel<<<148 * 8, 256, 0, s>>>(Yh, n, rows, didx, n_idx, out);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(didx);
    cudaFree(out);
    cudaFreeHost(Yh);
    RUVIII_LAUNCH_CHECK();
    return best;
}

// Copy kernels with the update kernel's access pattern (no arithmetic), to separate "what the memory system gives
// this tiling" from "what the update kernel loses on top".  HINT: streaming cache hints on/off.
template <int ROWS_IN_FLIGHT, bool HINT>
__global__ void __launch_bounds__(128)
tiled_copy_kernel(const double* __restrict__ Y, double* __restrict__ out, int64_t ld, int rows, int n_pairs, int rows_per_cta) {
    const int pair = blockIdx.x * 128 + threadIdx.x;
    if (pair >= n_pairs) return;
    const int64_t col = 2 * (int64_t)pair;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(r0 + rows_per_cta, rows);
    for (int rb = r0; rb < r1; rb += ROWS_IN_FLIGHT) {
        double2 y[ROWS_IN_FLIGHT];
#pragma unroll
        for (int u = 0; u < ROWS_IN_FLIGHT; ++u) {
            if (rb + u < r1) {
                const double* p = Y + (int64_t)(rb + u) * ld + col;
                y[u] = HINT ? ld_stream(p) : *reinterpret_cast<const double2*>(p);
            }
        }
#pragma unroll
        for (int u = 0; u < ROWS_IN_FLIGHT; ++u) {
            if (rb + u < r1) {
                double* p = out + (int64_t)(rb + u) * ld + col;
                if (HINT) st_stream(p, y[u]); else *reinterpret_cast<double2*>(p) = y[u];
            }
        }
    }
}

// GB/s (read + written) of several copy variants over an 11,000 x 20,000 FP64 matrix; results[8]
void probe_copy_variants(double* results, cudaStream_t s) {
    const int rows = 11000, n = 20000;
    const int64_t ld = n;
    const size_t bytes = (size_t)rows * ld * 8;
    double *a = nullptr, *b = nullptr;
    RUVIII_CUDA_CHECK(cudaMalloc(&a, bytes));
    RUVIII_CUDA_CHECK(cudaMalloc(&b, bytes));
    RUVIII_CUDA_CHECK(cudaMemsetAsync(a, 0, bytes, s));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int n_pairs = n / 2, col_tiles = ceil_div(n_pairs, 128);
    auto time_it = [&](auto&& launch) {
        float best = 1e30f;
        for (int it = 0; it < 6; ++it) {
            cudaEventRecord(e0, s);
            launch();
            cudaEventRecord(e1, s);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (it >= 2) best = std::min(best, ms);
        }
        return 2.0 * bytes / (best * 1e-3) / 1e9;
    };
    results[0] = time_it([&] { copy_kernel<<<148 * 16, 256, 0, s>>>((const double2*)a, (double2*)b, bytes / 16); });
    results[1] = time_it([&] { tiled_copy_kernel<4, false><<<dim3(col_tiles, ceil_div(rows, 64)), 128, 0, s>>>(a, b, ld, rows, n_pairs, 64); });
    results[2] = time_it([&] { tiled_copy_kernel<4, true><<<dim3(col_tiles, ceil_div(rows, 64)), 128, 0, s>>>(a, b, ld, rows, n_pairs, 64); });
    results[3] = time_it([&] { tiled_copy_kernel<8, true><<<dim3(col_tiles, ceil_div(rows, 64)), 128, 0, s>>>(a, b, ld, rows, n_pairs, 64); });
    results[4] = time_it([&] { tiled_copy_kernel<4, true><<<dim3(col_tiles, ceil_div(rows, 16)), 128, 0, s>>>(a, b, ld, rows, n_pairs, 16); });
    results[5] = time_it([&] { tiled_copy_kernel<4, true><<<dim3(col_tiles, ceil_div(rows, 1600)), 128, 0, s>>>(a, b, ld, rows, n_pairs, 1600); });
    results[6] = time_it([&] { tiled_copy_kernel<16, true><<<dim3(col_tiles, ceil_div(rows, 64)), 128, 0, s>>>(a, b, ld, rows, n_pairs, 64); });
    results[7] = time_it([&] { tiled_copy_kernel<2, false><<<dim3(col_tiles, ceil_div(rows, 64)), 128, 0, s>>>(a, b, ld, rows, n_pairs, 64); });
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(a);
    cudaFree(b);
    RUVIII_LAUNCH_CHECK();
}

float probe_copy_gbs(size_t bytes, cudaStream_t s) {
    double2 *a = nullptr, *b = nullptr;
    const size_t n2 = bytes / sizeof(double2);
    RUVIII_CUDA_CHECK(cudaMalloc(&a, n2 * sizeof(double2)));
    RUVIII_CUDA_CHECK(cudaMalloc(&b, n2 * sizeof(double2)));
    RUVIII_CUDA_CHECK(cudaMemsetAsync(a, 0, n2 * sizeof(double2), s));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 8; ++it) {
        cudaEventRecord(e0, s);
        copy_kernel<<<148 * 16, 256, 0, s>>>(a, b, n2);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 2) best = std::min(best, ms);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(a);
    cudaFree(b);
    RUVIII_LAUNCH_CHECK();
    return (float)(2.0 * n2 * sizeof(double2) / (best * 1e-3) / 1e9);
}

float probe_dmma_tflops(int n_sms, cudaStream_t s) {
    double* out = nullptr;
    RUVIII_CUDA_CHECK(cudaMalloc(&out, 8));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4096, blocks = n_sms * 4;
    float best = 1e30f;
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(e0, s);
        dmma_peak_kernel<<<blocks, 256, 0, s>>>(out, iters);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 1) best = std::min(best, ms);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    RUVIII_LAUNCH_CHECK();
    const double flops = 2.0 * 256.0 * 16.0 * iters * (double)blocks * 8.0;   // 8 warps x 16 DMMA x 256 FMA
    return (float)(flops / (best * 1e-3) / 1e12);
}

}  // namespace ruviii

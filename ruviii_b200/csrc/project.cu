// K3 post-processing, K4 (projection), K5 (control-gene products, k x k solve).
//
// Reference statements (ruvIII.R):
//   :140  fullalpha <- t(svd(Y0 %*% t(Y0))$u[, 1:r]) %*% Y
//   :142  alpha <- fullalpha[1:min(k, nrow(fullalpha)), ]
//   :143  ac <- alpha[, ctl]
//   :144  W <- Y.stand[, ctl] %*% t(ac) %*% solve(ac %*% t(ac))
// All of these are skinny (k <= 32 rows), so they are bandwidth problems, not tensor problems.
#include "kernels.cuh"

#include <cstdlib>

namespace ruviii {

namespace {

// ---- K3 post: top eigenvectors, descending, deterministic sign -------------------------------------------
// Block j handles component j (0 = largest eigenvalue): Vd[:, j] (column-major, optional) for j < n_keep,
// U[:, j] (row-major R x kp, what the projection kernel stages) for j < kk, zero padding for kk <= j < kp.
__global__ void __launch_bounds__(256)
pick_top_kernel(const double* __restrict__ V, int64_t ldv, const double* __restrict__ w_asc, int n_found,
                int R, int n_keep, int kk, int kp, double* __restrict__ Vd, double* __restrict__ U,
                double* __restrict__ w_desc) {
    const int j = blockIdx.x;
    __shared__ double s_abs[256];
    __shared__ double s_val[256];
    if (j >= n_keep) {
        if (j < kp)
            for (int r = threadIdx.x; r < R; r += blockDim.x) U[(int64_t)r * kp + j] = 0.0;
        return;
    }
    const double* v = V + (int64_t)(n_found - 1 - j) * ldv;
    double best_abs = -1.0, best_val = 0.0;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const double x = v[r], a = fabs(x);
        if (a > best_abs) { best_abs = a; best_val = x; }
    }
    s_abs[threadIdx.x] = best_abs;
    s_val[threadIdx.x] = best_val;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o && s_abs[threadIdx.x + o] > s_abs[threadIdx.x]) {
            s_abs[threadIdx.x] = s_abs[threadIdx.x + o];
            s_val[threadIdx.x] = s_val[threadIdx.x + o];
        }
        __syncthreads();
    }
    const double sgn = s_val[0] < 0.0 ? -1.0 : 1.0;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const double x = sgn * v[r];
        if (Vd) Vd[(int64_t)j * R + r] = x;
        if (j < kp) U[(int64_t)r * kp + j] = (j < kk) ? x : 0.0;
    }
    if (threadIdx.x == 0) w_desc[j] = w_asc[n_found - 1 - j];
}

// ---- K4: part[split][j][g] = sum_{r in chunk} U[r][j] * Y0[r][g] -----------------------------------------
// One thread owns two adjacent genes; a CTA owns 256 genes x one chunk of rows.  U rows are staged through
// shared memory and read as warp-wide broadcasts.
constexpr int kProjThreads = 128;
constexpr int kProjRows = 32;                      // rows staged per step

// KU <= KP: columns of U that are not padding (a multiple of 4): the others are neither accumulated nor read.
// The kernel is a stream over Y0 with 2 KU DFMA per 16 bytes.  Each thread keeps TWO batches of eight rows in registers:
// the loads of the next batch are issued before the current one is used, so a warp always has 8 x 16 B per thread in
// flight.  (Four rows in flight, three CTAs per SM: 2.0 TB/s.  Eight rows, four CTAs, but load -> wait -> compute in turn:
// 3.4 TB/s -- the kernel lasted as long as one warp's chain of 16 exposed memory latencies.)
template <int KP, int KU>
__global__ void __launch_bounds__(kProjThreads, KU <= 12 ? 4 : (KU <= 24 ? 3 : 2))
project_kernel(const double* __restrict__ Y0, int64_t ldY0, int R, int n_pairs, const double* __restrict__ U,
               int rows_per_split, double* __restrict__ part, int64_t ld) {
    __shared__ __align__(16) double sU[kProjRows][KU];
    const int pair = blockIdx.x * kProjThreads + threadIdx.x;
    const bool active = pair < n_pairs;
    const int64_t col = 2 * (int64_t)(active ? pair : 0);
    const int r0 = blockIdx.y * rows_per_split;
    const int r1 = min(r0 + rows_per_split, R);
    double2 acc[KU];
#pragma unroll
    for (int j = 0; j < KU; ++j) acc[j] = make_double2(0.0, 0.0);
    double2 y[8], yn[8];
    auto load8 = [&](double2 (&dst)[8], int row) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
            dst[u] = (row + u < r1) ? ld_stream(Y0 + (int64_t)(row + u) * ldY0 + col) : make_double2(0.0, 0.0);
    };
    if (r0 < r1) load8(y, r0);
    for (int rb = r0; rb < r1; rb += kProjRows) {
        const int nr = min(kProjRows, r1 - rb);
        __syncthreads();
        for (int e = threadIdx.x; e < kProjRows * KU; e += kProjThreads) {
            const int rr = e / KU, j = e - rr * KU;
            sU[rr][j] = rr < nr ? U[(int64_t)(rb + rr) * KP + j] : 0.0;
        }
        __syncthreads();
#pragma unroll 1
        for (int rr = 0; rr < kProjRows; rr += 8) {
            if (rr >= nr) break;
            load8(yn, rb + rr + 8);                               // rows past r1 come back as zeros without a load
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int j = 0; j < KU; j += 2) {
                    const double2 w = *reinterpret_cast<const double2*>(&sU[rr + u][j]);
                    acc[j].x = fma(w.x, y[u].x, acc[j].x);
                    acc[j].y = fma(w.x, y[u].y, acc[j].y);
                    acc[j + 1].x = fma(w.y, y[u].x, acc[j + 1].x);
                    acc[j + 1].y = fma(w.y, y[u].y, acc[j + 1].y);
                }
#pragma unroll
            for (int u = 0; u < 8; ++u) y[u] = yn[u];
        }
    }
    if (active) {
        double* out = part + (int64_t)blockIdx.y * KP * ld + col;
#pragma unroll
        for (int j = 0; j < KP; ++j)
            *reinterpret_cast<double2*>(out + (int64_t)j * ld) = j < KU ? acc[j < KU ? j : 0] : make_double2(0.0, 0.0);
    }
}

__global__ void __launch_bounds__(256)
alpha_finalize_kernel(const double* __restrict__ part, int rsplit, int kp, int64_t ld,
                      double* __restrict__ alpha) {
    // grid.y = j, threads over the padded pitch (pad columns are sums of zeros)
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (g >= ld) return;
    double s = 0.0;
    for (int sp = 0; sp < rsplit; ++sp) s += part[((int64_t)sp * kp + j) * ld + g];
    alpha[(int64_t)j * ld + g] = s;
}

// Gene-side route: with Z'Z = V diag(lambda) V' the left vectors are U = Z V / sigma, so
// fullalpha = t(U) %*% Z = diag(sigma) V':  alpha[j][g] = sqrt(lambda_j) V[g][j].  V is row-major n x kp.
__global__ void __launch_bounds__(256)
alpha_from_v_kernel(const double* __restrict__ V, const double* __restrict__ evals, int n, int kk, int kp, int64_t ld,
                    double* __restrict__ alpha) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (g >= ld) return;
    double v = 0.0;
    if (j < kk && g < n) v = sqrt(fmax(evals[j], 0.0)) * V[g * kp + j];
    alpha[(int64_t)j * ld + g] = v;
}

__global__ void __launch_bounds__(256)
gather_ac_kernel(const double* __restrict__ alpha, int64_t ld, int kp, const int* __restrict__ ctl_idx,
                 int n_ctl, double* __restrict__ acT) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_ctl * kp) return;
    const int c = e / kp, j = e - c * kp;
    acT[e] = alpha[(int64_t)j * ld + ctl_idx[c]];
}

// ---- K5: A[i][j] = sum_c Y[i][ctl_idx[c]] * acT[c][j] -------------------------------------------------------------
// A gather: every control gene of every row costs one 32-byte sector (a 64-byte DRAM granule) for 8 useful bytes, so
// the kernel is bound by sectors in flight, not by arithmetic.  The gather therefore runs through cp.async (8-byte
// copies straight into shared memory, no registers held per outstanding load): a CTA owns 32 rows, tiles of 128
// control genes are double buffered, so 2 x 4,096 gathers per CTA are in flight while the previous tile is
// multiplied against its slice of acT.  Products are plain DFMA (2 m c kp flop is negligible).
constexpr int kCtlThreads = 256;
constexpr int kCtlRowsPerCta = 32;
constexpr int kCtlTile = 128;
constexpr int kCtlYStride = kCtlTile + 1;          // doubles; odd stride: the 32 rows of a column land in distinct banks

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, bool valid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(gsrc), "r"(sz) : "memory");
}

// Yc[r][c] = Y[r][ctl_idx[c]]: the gather itself needs no eigenvector, so the engine runs it on a second stream
// while the (latency-bound, few-CTA) eigensolve occupies the main stream; the products then read Yc contiguously.
constexpr int kGatherRows = 8;
// log_rows > 0: rows [0, log_rows) are raw assay values still to be transformed (the pipelined host path gathers them
// straight out of pinned host memory, before the device copy of the assay exists): apply ruvIII.R:89 on the fly.
__global__ void __launch_bounds__(256)
ctl_gather_kernel(RowSrc src, int rows_w, const int* __restrict__ ctl_idx, int n_ctl, double* __restrict__ Yc, int64_t ldc,
                  int log_rows, double pseudo_count) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= n_ctl) return;
    const int g = ctl_idx[c];
    const int r0 = blockIdx.y * kGatherRows;
    // The gather streams ~1 GB of sectors through L2 while the eigensolver on the main stream lives off a 5 MB
    // L2-resident Gram: mark both sides of the gather evict-first so it does not push that working set out.
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    double y[kGatherRows];
#pragma unroll
    for (int u = 0; u < kGatherRows; ++u) {
        y[u] = 0.0;
        if (r0 + u < rows_w)
            asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;"
                         : "=d"(y[u]) : "l"(src.row(r0 + u) + g), "l"(pol));
    }
#pragma unroll
    for (int u = 0; u < kGatherRows; ++u)
        if (r0 + u < rows_w) {
            const double v = (r0 + u < log_rows) ? log2(y[u] + pseudo_count) : y[u];
            asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(Yc + (int64_t)(r0 + u) * ldc + c), "d"(v), "l"(pol) : "memory");
        }
}

template <int KP>
__global__ void __launch_bounds__(kCtlThreads)
ctl_products_kernel(RowSrc src, int rows_w, const int* __restrict__ ctl_idx, int n_ctl,
                    const double* __restrict__ acT, double* __restrict__ A, const double* __restrict__ Yc, int64_t ldc) {
    constexpr int AS = KP + 4;                       // acT tile row stride (doubles), keeps rows 16-byte aligned
    extern __shared__ __align__(16) double ctl_smem[];
    double* sY = ctl_smem;                           // [2][32][129]
    double* sAc = ctl_smem + 2 * kCtlRowsPerCta * kCtlYStride;   // [2][128][AS]
    const int tid = threadIdx.x;
    const int row0 = blockIdx.x * kCtlRowsPerCta;
    const int n_tiles = (n_ctl + kCtlTile - 1) / kCtlTile;
    // gather role: control gene c = tid % 128 of the tile, rows (tid / 128) + 2 i, i = 0..15
    const int gc = tid & (kCtlTile - 1), gr0 = tid >> 7;
    auto issue_tile = [&](int t, int buf) {
        const int c = t * kCtlTile + gc;
        const bool cok = c < n_ctl;
        double* dY = sY + buf * kCtlRowsPerCta * kCtlYStride + gc;
        if (Yc) {                                      // pre-gathered: contiguous 8-byte copies, whole lines per warp
#pragma unroll
            for (int i = 0; i < kCtlRowsPerCta / 2; ++i) {
                const int r = gr0 + 2 * i;
                cp_async8(dY + r * kCtlYStride, Yc + (int64_t)min(row0 + r, rows_w - 1) * ldc + (cok ? c : 0), cok);
            }
        } else {
            const int g = ctl_idx[cok ? c : 0];
#pragma unroll
            for (int i = 0; i < kCtlRowsPerCta / 2; ++i) {
                const int r = gr0 + 2 * i;
                cp_async8(dY + r * kCtlYStride, src.row(min(row0 + r, rows_w - 1)) + g, cok);
            }
        }
        double* dA = sAc + buf * kCtlTile * AS;
        for (int e = tid; e < kCtlTile * KP; e += kCtlThreads) {
            const int cc = e / KP, j = e - cc * KP;
            const bool ok = t * kCtlTile + cc < n_ctl;
            cp_async8(dA + cc * AS + j, acT + (int64_t)(ok ? t * kCtlTile + cc : 0) * KP + j, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // compute role: row r = tid / 8, columns j = (tid % 8) + 8 u
    constexpr int NJ = (KP + 7) / 8;
    const int cr = tid >> 3, cj = tid & 7;
    double acc[NJ];
#pragma unroll
    for (int u = 0; u < NJ; ++u) acc[u] = 0.0;
    issue_tile(0, 0);
    for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < n_tiles) {
            issue_tile(t + 1, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const double* y = sY + buf * kCtlRowsPerCta * kCtlYStride + cr * kCtlYStride;
        const double* a = sAc + buf * kCtlTile * AS + cj;
#pragma unroll 8
        for (int c = 0; c < kCtlTile; ++c) {
            const double yv = y[c];
#pragma unroll
            for (int u = 0; u < NJ; ++u)
                if (cj + 8 * u < KP) acc[u] = fma(yv, a[c * AS + 8 * u], acc[u]);
        }
        __syncthreads();                              // the buffer is refilled by the next iteration's issue_tile
    }
    if (row0 + cr < rows_w) {
#pragma unroll
        for (int u = 0; u < NJ; ++u)
            if (cj + 8 * u < KP) A[(int64_t)(row0 + cr) * KP + cj + 8 * u] = acc[u];
    }
}

// K5 on the pre-gathered block: A (rows x kp) = Yc (rows x c, dense) . acT (c x kp) as a skinny FP64 GEMM on DMMA.
// The DFMA kernel above spends 1.5 shared-memory loads per FMA (0.09 ms for 98 MB at C3 -- bound by shared-memory
// bandwidth, not by HBM); one m8n8k4 DMMA takes two 8-byte fragment loads per lane for 256 FMAs.  A CTA owns 32 rows,
// tiles of 128 control genes arrive by 16-byte cp.async (double buffered, zero-filled past rows_w / n_ctl), warp w
// owns row tile w % 4 and the column tiles (w / 4) + 2u.  Row stride 132 and acT stride KPE + 4 make the fragment loads
// of a half-warp hit 16 distinct 8-byte banks.
constexpr int kCtlYStrideMma = kCtlTile + 4;

__device__ __forceinline__ void dmma884p(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16z(void* smem_dst, const void* gsrc, int src_bytes) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gsrc), "r"(src_bytes) : "memory");
}

template <int KP>
__global__ void __launch_bounds__(kCtlThreads)
ctl_products_mma_kernel(const double* __restrict__ Yc, int64_t ldc, int rows_w, int n_ctl, const double* __restrict__ acT,
                        double* __restrict__ A) {
    constexpr int KPE = KP < 8 ? 8 : KP;             // column extent in whole 8-wide DMMA tiles
    constexpr int AS = KPE + 4;
    constexpr int NCT = KPE / 8;                     // column tiles
    constexpr int NU = NCT >= 2 ? NCT / 2 : 1;       // column tiles per warp
    extern __shared__ __align__(16) double ctl_smem[];
    double* sY = ctl_smem;                                          // [2][32][132]
    double* sAc = ctl_smem + 2 * kCtlRowsPerCta * kCtlYStrideMma;   // [2][128][AS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = blockIdx.x * kCtlRowsPerCta;
    const int n_tiles = (n_ctl + kCtlTile - 1) / kCtlTile;
    for (int e = tid; e < 2 * kCtlTile * AS; e += kCtlThreads) sAc[e] = 0.0;     // columns >= KP stay zero
    __syncthreads();
    auto issue_tile = [&](int t, int buf) {
        double* dY = sY + buf * kCtlRowsPerCta * kCtlYStrideMma;
#pragma unroll
        for (int i = 0; i < kCtlRowsPerCta * (kCtlTile / 2) / kCtlThreads; ++i) {
            const int e = tid + i * kCtlThreads;
            const int r = e / (kCtlTile / 2), c2 = (e % (kCtlTile / 2)) * 2;
            const int c = t * kCtlTile + c2;
            const int left = (row0 + r < rows_w) ? min(2, max(0, n_ctl - c)) : 0;       // valid doubles of this chunk
            cp_async16z(dY + r * kCtlYStrideMma + c2, Yc + (int64_t)min(row0 + r, rows_w - 1) * ldc + (left ? c : 0), left * 8);
        }
        double* dA = sAc + buf * kCtlTile * AS;
        for (int e = tid; e < kCtlTile * (KP / 2); e += kCtlThreads) {
            const int cc = e / (KP / 2), j = (e - cc * (KP / 2)) * 2;
            const bool ok = t * kCtlTile + cc < n_ctl;
            cp_async16z(dA + cc * AS + j, acT + (int64_t)(ok ? t * kCtlTile + cc : 0) * KP + j, ok ? 16 : 0);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int rt = warp & 3, ct0 = warp >> 2;
    const bool active = ct0 < NCT;
    const int fr = lane >> 2, fk = lane & 3;
    double acc[NU][2];
#pragma unroll
    for (int u = 0; u < NU; ++u) acc[u][0] = acc[u][1] = 0.0;
    issue_tile(0, 0);
    for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < n_tiles) {
            issue_tile(t + 1, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        if (active) {
            const double* y = sY + buf * kCtlRowsPerCta * kCtlYStrideMma + (rt * 8 + fr) * kCtlYStrideMma + fk;
            const double* a = sAc + buf * kCtlTile * AS + fk * AS + ct0 * 8 + fr;
#pragma unroll 8
            for (int kk = 0; kk < kCtlTile; kk += 4) {
                const double av = y[kk];
#pragma unroll
                for (int u = 0; u < NU; ++u) dmma884p(acc[u][0], acc[u][1], av, a[kk * AS + u * 16]);
            }
        }
        __syncthreads();                              // the buffer is refilled by the next iteration's issue_tile
    }
    const int row = row0 + rt * 8 + fr;
    if (active && row < rows_w) {
#pragma unroll
        for (int u = 0; u < NU; ++u) {
            const int col = (ct0 + 2 * u) * 8 + fk * 2;
            if (col < KP) A[(int64_t)row * KP + col] = acc[u][0];
            if (col + 1 < KP) A[(int64_t)row * KP + col + 1] = acc[u][1];
        }
    }
}

// Gm[a][b] = sum_c acT[c][a] acT[c][b]   (kp x kp).  One CTA: acT (128 KB at C3) is brought into shared memory in as few
// chunks as fit (every thread issues all its 16-byte loads before the first wait: one memory round trip per chunk, not one
// per 64 control genes -- that version took 35 us and was the longest thing in the control-block stage); thread (pair e,
// quarter s) sums rows s, s+4, ... and the four quarters are added in a fixed order, so the result is bitwise reproducible.
constexpr int kAcSmemDoubles = 24 * 1024;            // 192 KB of acT per chunk
__global__ void __launch_bounds__(1024)
ac_gram_kernel(const double* __restrict__ acT, int n_ctl, int kp, int chunk_rows, double* __restrict__ Gm) {
    extern __shared__ __align__(16) double ac_smem[];
    double* sA = ac_smem;                                  // chunk_rows x kp
    double* sP = ac_smem + (size_t)chunk_rows * kp;        // 4 x kMaxK^2
    const int tid = threadIdx.x;
    const int npair = kp * kp;
    const int quarter = tid / 256, e = tid % 256;       // kp <= 16: one pair per thread; kp = 32: four pairs per thread
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int c0 = 0; c0 < n_ctl; c0 += chunk_rows) {
        const int nc = min(chunk_rows, n_ctl - c0);
        __syncthreads();
        const double2* src = reinterpret_cast<const double2*>(acT + (int64_t)c0 * kp);      // kp is even, rows are contiguous
        double2* dst = reinterpret_cast<double2*>(sA);
        const int n2 = nc * kp / 2;
#pragma unroll 8
        for (int i = tid; i < n2; i += 1024) dst[i] = src[i];
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int pe = e + u * 256;
            if (pe < npair) {
                const int a = pe / kp, b = pe - a * kp;
                double s0 = 0.0, s1 = 0.0;                    // two chains: the sum order is still fixed
                int r = quarter;
                for (; r + 4 < nc; r += 8) {
                    s0 = fma(sA[r * kp + a], sA[r * kp + b], s0);
                    s1 = fma(sA[(r + 4) * kp + a], sA[(r + 4) * kp + b], s1);
                }
                if (r < nc) s0 = fma(sA[r * kp + a], sA[r * kp + b], s0);
                acc[u] += s0 + s1;
            }
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int pe = e + u * 256;
        if (pe < npair) sP[quarter * kMaxK * kMaxK + pe] = acc[u];
    }
    __syncthreads();
    for (int pe = tid; pe < npair; pe += 1024)
        Gm[pe] = ((sP[pe] + sP[kMaxK * kMaxK + pe]) + sP[2 * kMaxK * kMaxK + pe]) + sP[3 * kMaxK * kMaxK + pe];
}

// ---- K5c: column means of A, Cholesky of Gm[:kk,:kk], inverse of the factor -------------------------------
__global__ void __launch_bounds__(1024)
chol_solve_kernel(const double* __restrict__ A, int rows_w, int mu_rows, const double* __restrict__ Gm, int kk,
                  int kp, double* __restrict__ Linv, double* __restrict__ colmean, int* __restrict__ status) {
    __shared__ double L[kMaxK][kMaxK + 1];
    __shared__ double Li[kMaxK][kMaxK + 1];
    __shared__ double red[1024];
    __shared__ int bad;
    const int tid = threadIdx.x;
    // column means over the rows that define mu (ruvIII.R:118: all m rows; fastruvIII.R:136: the m1 samples)
    {
        const int j = tid % kp, lane_row = tid / kp, stride = 1024 / kp;
        double s = 0.0;
        if (lane_row < stride)
            for (int i = lane_row; i < mu_rows; i += stride) s += A[(int64_t)i * kp + j];
        red[tid] = (lane_row < stride) ? s : 0.0;
        __syncthreads();
        if (tid < kp) {
            double t = 0.0;
            for (int q = 0; q < stride; ++q) t += red[q * kp + tid];
            colmean[tid] = mu_rows > 0 ? t / (double)mu_rows : 0.0;   // mu_rows = 0: no centring (RUV1)
        }
    }
    const int r = tid / kMaxK, c = tid % kMaxK;       // 32 x 32 threads, one per matrix element
    if (tid == 0) bad = 0;
    L[r][c] = (r < kk && c < kk) ? Gm[r * kp + c] : (r == c ? 1.0 : 0.0);
    Li[r][c] = 0.0;
    __syncthreads();
    for (int j = 0; j < kk; ++j) {                    // right-looking Cholesky
        if (tid == 0) {
            // R's solve() (dgesv + rcond check) stops on a computationally singular matrix; an exactly singular Gm leaves a
            // pivot of a few ulps of its diagonal entry with either sign, so "not positive" alone would miss it
            const double d = L[j][j];
            if (!(d > 16.0 * 2.220446049250313e-16 * Gm[j * kp + j])) { bad = j + 1; L[j][j] = 1.0; } else L[j][j] = sqrt(d);
        }
        __syncthreads();
        if (c == j && r > j && r < kk) L[r][j] /= L[j][j];
        __syncthreads();
        if (r > j && c > j && c <= r && r < kk) L[r][c] -= L[r][j] * L[c][j];
        __syncthreads();
    }
    // inverse of the lower-triangular factor, one thread per column, forward substitution
    if (tid < kk) {
        const int col = tid;
        for (int i = col; i < kk; ++i) {
            double s = (i == col) ? 1.0 : 0.0;
            for (int q = col; q < i; ++q) s -= L[i][q] * Li[q][col];
            Li[i][col] = s / L[i][i];
        }
    }
    __syncthreads();
    for (int e = tid; e < kp * kp; e += 1024) {
        const int i = e / kp, j = e - i * kp;
        Linv[e] = (i < kk && j <= i) ? Li[i][j] : 0.0;
    }
    if (tid == 0) *status = bad;
}

__global__ void __launch_bounds__(256)
atilde_kernel(const double* __restrict__ A, int rows_w, const double* __restrict__ colmean,
              const double* __restrict__ Linv, int kk, int kp, double* __restrict__ At) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows_w * kp) return;
    const int i = e / kp, q = e - i * kp;
    double s = 0.0;
    if (q < kk)
        for (int j = 0; j <= q; ++j) s = fma(A[(int64_t)i * kp + j] - colmean[j], Linv[q * kp + j], s);
    At[e] = s;
}

__global__ void __launch_bounds__(256)
alpha_tilde_kernel(const double* __restrict__ alpha, const double* __restrict__ Linv, int kk, int kp, int64_t ld,
                   double* __restrict__ at) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int q = blockIdx.y;
    if (g >= ld) return;
    double s = 0.0;
    if (q < kk)
        for (int j = 0; j <= q; ++j) s = fma(Linv[q * kp + j], alpha[(int64_t)j * ld + g], s);
    at[(int64_t)q * ld + g] = s;
}

__global__ void __launch_bounds__(256)
w_kernel(const double* __restrict__ At, int rows_w, const double* __restrict__ Linv, int k, int kp,
         double* __restrict__ W) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows_w * k) return;
    const int i = e / k, j = e - i * k;
    double s = 0.0;
    for (int q = j; q < k; ++q) s = fma(At[(int64_t)i * kp + q], Linv[q * kp + j], s);
    W[e] = s;
}

}  // namespace

void launch_pick_top(const double* V, int64_t ldv, const double* w_asc, int n_found, int R, int n_keep, int kk,
                     int kp, double* Vd, double* U, double* w_desc, cudaStream_t s) {
    pick_top_kernel<<<std::max(n_keep, kp), 256, 0, s>>>(V, ldv, w_asc, n_found, R, n_keep, kk, kp, Vd, U, w_desc);
    RUVIII_LAUNCH_CHECK();
}

int project_rsplit(int R, int n, int n_sms) {
    const int tiles = ceil_div((n + 1) / 2, kProjThreads);
    int rs = ceil_div(4 * n_sms, tiles);
    rs = std::max(1, std::min(rs, ceil_div(R, kProjRows)));
    return rs;
}

void launch_project(const double* Y0, int64_t ldY0, int R, int n, const double* U, int kp, int rsplit,
                    double* part, int64_t ld, cudaStream_t s, int k_used) {
    const int n_pairs = (n + 1) / 2;
    const int rows_per_split = (int)round_up(ceil_div(R, rsplit), kProjRows);
    dim3 grid(ceil_div(n_pairs, kProjThreads), rsplit);
    const int ku = k_used > 0 ? std::min(kp, (int)round_up(k_used, 4)) : kp;      // columns of U beyond k_used are zero padding
#define RUVIII_PROJ(KP_, KU_) project_kernel<KP_, KU_><<<grid, kProjThreads, 0, s>>>(Y0, ldY0, R, n_pairs, U, rows_per_split, part, ld)
    switch (kp) {
        case 4: RUVIII_PROJ(4, 4); break;
        case 8: if (ku <= 4) RUVIII_PROJ(8, 4); else RUVIII_PROJ(8, 8); break;
        case 16: if (ku <= 12) RUVIII_PROJ(16, 12); else RUVIII_PROJ(16, 16); break;
        default:
            if (ku <= 20) RUVIII_PROJ(32, 20);
            else if (ku <= 24) RUVIII_PROJ(32, 24);
            else if (ku <= 28) RUVIII_PROJ(32, 28);
            else RUVIII_PROJ(32, 32);
            break;
    }
#undef RUVIII_PROJ
    RUVIII_LAUNCH_CHECK();
}

void launch_alpha_finalize(const double* part, int rsplit, int kp, int n, int64_t ld, double* alpha,
                           const int* ctl_idx, int n_ctl, double* acT, cudaStream_t s) {
    dim3 grid(ceil_div(ld, 256), kp);
    alpha_finalize_kernel<<<grid, 256, 0, s>>>(part, rsplit, kp, ld, alpha);
    RUVIII_LAUNCH_CHECK();
    if (n_ctl > 0) {
        gather_ac_kernel<<<ceil_div((int64_t)n_ctl * kp, 256), 256, 0, s>>>(alpha, ld, kp, ctl_idx, n_ctl, acT);
        RUVIII_LAUNCH_CHECK();
    }
}

void launch_alpha_from_v(const double* V, const double* evals, int n, int kk, int kp, int64_t ld, double* alpha,
                         cudaStream_t s) {
    dim3 grid(ceil_div(ld, 256), kp);
    alpha_from_v_kernel<<<grid, 256, 0, s>>>(V, evals, n, kk, kp, ld, alpha);
    RUVIII_LAUNCH_CHECK();
}

void launch_ctl_gather(RowSrc src, int rows_w, const int* ctl_idx, int n_ctl, double* Yc, int64_t ldc, int log_rows,
                       double pseudo_count, cudaStream_t s) {
    if (n_ctl <= 0 || rows_w <= 0) return;
    dim3 grid(ceil_div(n_ctl, 256), ceil_div(rows_w, kGatherRows));
    ctl_gather_kernel<<<grid, 256, 0, s>>>(src, rows_w, ctl_idx, n_ctl, Yc, ldc, log_rows, pseudo_count);
    RUVIII_LAUNCH_CHECK();
}

void launch_ctl_products(RowSrc src, int rows_w, const int* ctl_idx, int n_ctl, const double* acT, int kp,
                         double* A, double* Gm, const double* Yc, int64_t ldc, cudaStream_t s) {
    const int blocks = ceil_div(rows_w, kCtlRowsPerCta);
    static const bool use_mma = [] { const char* v = std::getenv("RUVIII_CTL_MMA"); return !(v && *v == '0'); }();
    if (n_ctl > 0) {
#define RUVIII_CTL(KP)                                                                                                   \
    do {                                                                                                                 \
        if (Yc && use_mma) {                                                                                             \
            const int smem = (2 * kCtlRowsPerCta * kCtlYStrideMma + 2 * kCtlTile * ((KP < 8 ? 8 : KP) + 4)) * 8;         \
            RUVIII_CUDA_CHECK(cudaFuncSetAttribute(ctl_products_mma_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
            ctl_products_mma_kernel<KP><<<blocks, kCtlThreads, smem, s>>>(Yc, ldc, rows_w, n_ctl, acT, A);               \
        } else {                                                                                                         \
            const int smem = (2 * kCtlRowsPerCta * kCtlYStride + 2 * kCtlTile * (KP + 4)) * 8;                           \
            RUVIII_CUDA_CHECK(cudaFuncSetAttribute(ctl_products_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
            ctl_products_kernel<KP><<<blocks, kCtlThreads, smem, s>>>(src, rows_w, ctl_idx, n_ctl, acT, A, Yc, ldc);     \
        }                                                                                                                \
    } while (0)
        switch (kp) {
            case 4: RUVIII_CTL(4); break;
            case 8: RUVIII_CTL(8); break;
            case 16: RUVIII_CTL(16); break;
            default: RUVIII_CTL(32); break;
        }
#undef RUVIII_CTL
    } else {
        RUVIII_CUDA_CHECK(cudaMemsetAsync(A, 0, (size_t)rows_w * kp * sizeof(double), s));
    }
    RUVIII_LAUNCH_CHECK();
    if (Gm) launch_ac_gram(acT, n_ctl, kp, Gm, s);
}

void launch_ac_gram(const double* acT, int n_ctl, int kp, double* Gm, cudaStream_t s) {
    const int chunk_rows = std::max(1, std::min(n_ctl, kAcSmemDoubles / kp));
    const size_t smem = ((size_t)chunk_rows * kp + 4 * kMaxK * kMaxK) * sizeof(double);
    RUVIII_CUDA_CHECK(cudaFuncSetAttribute(ac_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ac_gram_kernel<<<1, 1024, smem, s>>>(acT, n_ctl, kp, chunk_rows, Gm);
    RUVIII_LAUNCH_CHECK();
}

void launch_chol_solve(const double* A, int rows_w, int mu_rows, const double* Gm, int kk, int kp, double* Linv,
                       double* colmean, int* status, cudaStream_t s) {
    chol_solve_kernel<<<1, 1024, 0, s>>>(A, rows_w, mu_rows, Gm, kk, kp, Linv, colmean, status);
    RUVIII_LAUNCH_CHECK();
}

void launch_atilde(const double* A, int rows_w, const double* colmean, const double* Linv, int kk, int kp,
                   double* At, cudaStream_t s) {
    atilde_kernel<<<ceil_div((int64_t)rows_w * kp, 256), 256, 0, s>>>(A, rows_w, colmean, Linv, kk, kp, At);
    RUVIII_LAUNCH_CHECK();
}

void launch_alpha_tilde(const double* alpha, const double* Linv, int kk, int kp, int n, int64_t ld, double* at,
                        cudaStream_t s) {
    dim3 grid(ceil_div(ld, 256), kp);
    alpha_tilde_kernel<<<grid, 256, 0, s>>>(alpha, Linv, kk, kp, ld, at);
    RUVIII_LAUNCH_CHECK();
}

void launch_w_from_atilde(const double* At, int rows_w, const double* Linv, int k, int kp, double* W,
                          cudaStream_t s) {
    if (k <= 0) return;
    w_kernel<<<ceil_div((int64_t)rows_w * k, 256), 256, 0, s>>>(At, rows_w, Linv, k, kp, W);
    RUVIII_LAUNCH_CHECK();
}

}  // namespace ruviii

// K1: segmented replicate-group mean subtraction (residop) and K8: log2 transform.
//
// Reference: ruvIII.R:139 `Y0 = residop(Y, M)`, in-tree statement fastruvIII.R:88-95
//   A - B solve(B'B) B'A  with B = M the one-hot replicate matrix (ruvIII.R:101).
// B'B = diag(group sizes), so row i of the result is Y[i,] - mean_{j in g(i)} Y[j,].  Rows of singleton
// groups are exactly 0 and are never produced: Y0 is stored compactly, one row per member of a
// non-singleton group, members of one group adjacent.
//
// HBM-bound: 16 B per produced cell (8 read + 8 written).  One thread owns two adjacent genes (16-byte
// accesses, a warp covers 512 contiguous bytes of a row); group members stay in registers between the
// summation and the subtraction so every input byte is read once.
#include "kernels.cuh"

#include <algorithm>
#include <cstdlib>

namespace ruviii {

namespace {

constexpr int kResidopThreads = 128;
constexpr int kRegGroup = 8;     // members held in registers; larger groups take the two-pass path

__global__ void __launch_bounds__(kResidopThreads)
residop_kernel(RowSrc src, const int* __restrict__ grp_start, const int* __restrict__ grp_rows,
               int n_groups, int n_pairs, double* __restrict__ Y0, int64_t ldY0) {
    const int pair = blockIdx.x * kResidopThreads + threadIdx.x;
    if (pair >= n_pairs) return;
    const int64_t col = 2 * (int64_t)pair;
    for (int g = blockIdx.y; g < n_groups; g += gridDim.y) {
        const int s = grp_start[g];
        const int sz = grp_start[g + 1] - s;
        const double inv = 1.0 / (double)sz;           // (B'B)^-1 = diag(1/size), fastruvIII.R:90
        if (sz <= kRegGroup) {
            double2 v[kRegGroup];
            double2 sum = make_double2(0.0, 0.0);
#pragma unroll
            for (int t = 0; t < kRegGroup; ++t) {
                if (t < sz) {
                    v[t] = ld_stream(src.row(grp_rows[s + t]) + col);
                    sum.x += v[t].x;
                    sum.y += v[t].y;
                }
            }
            const double2 mean = make_double2(sum.x * inv, sum.y * inv);
#pragma unroll
            for (int t = 0; t < kRegGroup; ++t) {
                if (t < sz)
                    st_stream(Y0 + (int64_t)(s + t) * ldY0 + col, make_double2(v[t].x - mean.x, v[t].y - mean.y));
            }
        } else {
            double2 sum = make_double2(0.0, 0.0);
            for (int t = 0; t < sz; ++t) {
                const double2 v = *reinterpret_cast<const double2*>(src.row(grp_rows[s + t]) + col);
                sum.x += v.x;
                sum.y += v.y;
            }
            const double2 mean = make_double2(sum.x * inv, sum.y * inv);
            for (int t = 0; t < sz; ++t) {
                const double2 v = *reinterpret_cast<const double2*>(src.row(grp_rows[s + t]) + col);
                st_stream(Y0 + (int64_t)(s + t) * ldY0 + col, make_double2(v.x - mean.x, v.y - mean.y));
            }
        }
    }
}

// Helmert contrasts of every replicate group (see kernels.cuh): one pass, 8 B read per member cell and 8 B written
// per contrast cell.  Groups of up to kRegGroup members are loaded in one batch (all loads in flight before the
// first use); larger groups stream with a running sum.
// RG = members held in registers: PRPS sets have 2-4 members, and the 4-member instance needs half the registers of
// the 8-member one (80 -> occupancy 6 CTAs of 128 threads per SM), i.e. twice the loads in flight per SM.
template <int RG>
__global__ void __launch_bounds__(kResidopThreads)
helmert_kernel(RowSrc src, const int* __restrict__ grp_start, const int* __restrict__ grp_rows, int n_groups,
               int n_pairs, double* __restrict__ Z, int64_t ldZ) {
    const int pair = blockIdx.x * kResidopThreads + threadIdx.x;
    if (pair >= n_pairs) return;
    const int64_t col = 2 * (int64_t)pair;
    for (int g = blockIdx.y; g < n_groups; g += gridDim.y) {
        const int s = grp_start[g];
        const int sz = grp_start[g + 1] - s;
        double* out = Z + (int64_t)(s - g) * ldZ + col;        // every earlier group contributed (size - 1) rows
        if (sz <= RG) {
            double2 v[RG];
#pragma unroll
            for (int t = 0; t < RG; ++t)
                if (t < sz) v[t] = ld_stream(src.row(grp_rows[s + t]) + col);
            double2 run = v[0];
#pragma unroll
            for (int j = 1; j < RG; ++j) {
                if (j < sz) {
                    const double c = sqrt((double)j / (double)(j + 1)), inv = 1.0 / (double)j;
                    st_stream(out + (int64_t)(j - 1) * ldZ, make_double2(c * (run.x * inv - v[j].x), c * (run.y * inv - v[j].y)));
                    run.x += v[j].x;
                    run.y += v[j].y;
                }
            }
        } else {
            double2 run = ld_stream(src.row(grp_rows[s]) + col);
            for (int j = 1; j < sz; ++j) {
                const double2 y = ld_stream(src.row(grp_rows[s + j]) + col);
                const double c = sqrt((double)j / (double)(j + 1)), inv = 1.0 / (double)j;
                st_stream(out + (int64_t)(j - 1) * ldZ, make_double2(c * (run.x * inv - y.x), c * (run.y * inv - y.y)));
                run.x += y.x;
                run.y += y.y;
            }
        }
    }
}

// K9  ruvIII.R:148-149  newY <- ((1 / colSums(M)) * t(M)) %*% newY: one output row per replicate group (every
// column of M, singletons included), the mean of the group's rows of newY.  all_start/all_rows: CSR over ALL groups.
__global__ void __launch_bounds__(kResidopThreads)
group_mean_kernel(RowSrc src, const int* __restrict__ all_start, const int* __restrict__ all_rows, int n_groups,
                  int n_pairs, double* __restrict__ out, int64_t ld) {
    const int pair = blockIdx.x * kResidopThreads + threadIdx.x;
    if (pair >= n_pairs) return;
    const int64_t col = 2 * (int64_t)pair;
    for (int g = blockIdx.y; g < n_groups; g += gridDim.y) {
        const int s = all_start[g], e = all_start[g + 1];
        double2 sum = make_double2(0.0, 0.0);
        for (int t = s; t < e; ++t) {
            const double2 v = *reinterpret_cast<const double2*>(src.row(all_rows[t]) + col);
            sum.x += v.x;
            sum.y += v.y;
        }
        const double inv = 1.0 / (double)(e - s);
        st_stream(out + (int64_t)g * ld + col, make_double2(sum.x * inv, sum.y * inv));
    }
}

// out (cols x ld_out, row-major) = in' for in (rows x ld_in, row-major): 32 x 32 tiles through padded shared memory,
// coalesced on both sides.  Used by the gene-side Gram, whose contraction runs over the rows of Z.
__global__ void __launch_bounds__(256)
transpose_kernel(const double* __restrict__ in, int64_t ld_in, int rows, int cols, double* __restrict__ out,
                 int64_t ld_out) {
    __shared__ double t[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8)
        t[i][tx] = (r0 + i < rows && c0 + tx < cols) ? in[(int64_t)(r0 + i) * ld_in + c0 + tx] : 0.0;
    __syncthreads();
    for (int i = ty; i < 32; i += 8)
        if (c0 + i < cols && r0 + tx < rows) out[(int64_t)(c0 + i) * ld_out + r0 + tx] = t[tx][i];
}

__global__ void __launch_bounds__(256)
log2_kernel(double* __restrict__ Y, int64_t ld, int rows, int n, double pc) {
    const int64_t total = (int64_t)rows * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / n, c = i - r * n;
        double* p = Y + r * ld + c;
        *p = log2(*p + pc);
    }
}

}  // namespace

void launch_log2_inplace(double* Y, int64_t ld, int rows, int n, double pc, cudaStream_t s) {
    if (rows <= 0 || n <= 0) return;
    const int64_t total = (int64_t)rows * n;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
    log2_kernel<<<blocks, 256, 0, s>>>(Y, ld, rows, n, pc);
    RUVIII_LAUNCH_CHECK();
}

// grid.y of the segmented kernels.  One CTA per (gene strip, group) is 31,600 CTAs of ~1 us each at C3 -- bound by the
// CTA launch rate, not by HBM (measured 0.63 of peak) -- so the groups are walked by a grid-stride loop instead and the
// grid is sized to `waves` full waves of resident CTAs (16 CTAs of 128 threads per SM).
static int seg_grid_y(int grid_x, int n_groups) {
    static const int waves = [] { const char* v = std::getenv("RUVIII_SEG_WAVES"); return v && *v ? std::max(1, std::atoi(v)) : 2; }();
    static const int sms = [] {
        int dev = 0, n = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        return n;
    }();
    if (waves >= 1000) return std::min(n_groups, 65535);            // one CTA per group (the old launch shape)
    const int target = sms * 16 * waves;
    return std::max(1, std::min(std::min(n_groups, 65535), ceil_div(target, grid_x)));
}

void launch_residop(RowSrc src, const int* grp_start, const int* grp_rows, int n_groups, int /*max_group*/,
                    int n, double* Y0, int64_t ldY0, cudaStream_t s) {
    if (n_groups <= 0 || n <= 0) return;
    const int n_pairs = (n + 1) / 2;          // pitches are even and zero padded, so the odd tail pair is safe
    dim3 grid(ceil_div(n_pairs, kResidopThreads), 1);
    grid.y = seg_grid_y(grid.x, n_groups);
    residop_kernel<<<grid, kResidopThreads, 0, s>>>(src, grp_start, grp_rows, n_groups, n_pairs, Y0, ldY0);
    RUVIII_LAUNCH_CHECK();
}

void launch_group_mean(RowSrc src, const int* all_start, const int* all_rows, int n_groups, int n, double* out, int64_t ld,
                       cudaStream_t s) {
    if (n_groups <= 0 || n <= 0) return;
    const int n_pairs = (n + 1) / 2;
    dim3 grid(ceil_div(n_pairs, kResidopThreads), 1);
    grid.y = seg_grid_y(grid.x, n_groups);
    group_mean_kernel<<<grid, kResidopThreads, 0, s>>>(src, all_start, all_rows, n_groups, n_pairs, out, ld);
    RUVIII_LAUNCH_CHECK();
}

void launch_transpose(const double* in, int64_t ld_in, int rows, int cols, double* out, int64_t ld_out, cudaStream_t s) {
    if (rows <= 0 || cols <= 0) return;
    dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
    transpose_kernel<<<grid, 256, 0, s>>>(in, ld_in, rows, cols, out, ld_out);
    RUVIII_LAUNCH_CHECK();
}

void launch_helmert(RowSrc src, const int* grp_start, const int* grp_rows, int n_groups, int max_group, int n, double* Z,
                    int64_t ldZ, cudaStream_t s) {
    if (n_groups <= 0 || n <= 0) return;
    const int n_pairs = (n + 1) / 2;
    dim3 grid(ceil_div(n_pairs, kResidopThreads), 1);
    grid.y = seg_grid_y(grid.x, n_groups);
    if (max_group > 0 && max_group <= 4)
        helmert_kernel<4><<<grid, kResidopThreads, 0, s>>>(src, grp_start, grp_rows, n_groups, n_pairs, Z, ldZ);
    else
        helmert_kernel<kRegGroup><<<grid, kResidopThreads, 0, s>>>(src, grp_start, grp_rows, n_groups, n_pairs, Z, ldZ);
    RUVIII_LAUNCH_CHECK();
}

}  // namespace ruviii

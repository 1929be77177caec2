// K2: symmetric FP64 Gram  G = Z Z'  on the Blackwell FP64 tensor pipe (DMMA).
//
// Reference: ruvIII.R:140 `Y0 %*% t(Y0)` (a full dgemm; the symmetric half is R(R+1)L flop, which is what
// DESIGN.md counts).  Z is R x K row-major with the contraction (genes) contiguous, so both MMA operands are
// K-major and come from the same matrix.
//
// Shape of the work: in the PRPS regime the output is small (R ~ 10^2..10^3) and the contraction long
// (K = genes of this shard), so the upper-triangular output tiles are additionally split over K to fill 148 SMs;
// partial tiles go to a workspace and a second kernel sums them in a fixed order (bitwise reproducible, and
// symmetric by construction) -- that sum is also what feeds the NCCL allreduce.
//
// Two tile shapes, chosen per problem by padded work (plan_gram):
//   128 x 128, 8 warps as 2 x 4, warp tile 64 x 32 (8 x 4 DMMA tiles, 64 accumulators/thread), 1 CTA per SM;
//    80 x  80, 4 warps as 2 x 2, warp tile 40 x 40 (5 x 5 DMMA tiles, 50 accumulators/thread), 2 CTAs per SM --
//              R = 800 or 1,200 are multiples of 80, so nothing is padded, and two independent CTAs per SM cover
//              each other's barrier stalls.
// impl 1: 4-stage cp.async pipeline, shared-memory rows padded 16 -> 20 doubles so that the 8-byte DMMA
//         fragment loads of a half-warp (4 rows x 4 k) hit 16 distinct 8-byte banks.
// impl 2: TMA (cp.async.bulk.tensor) + mbarrier pipeline, 128-byte swizzle, 128 x 128 tiles (gram_tma.cu).
// impl 0: plain DFMA kernel kept as an in-library cross-check of the MMA fragment layouts.
#include "kernels.cuh"

#include <algorithm>

namespace ruviii {

int launch_gram_tma(const GramPlan& gp, const void* tma_desc, const int2* tile_map, double* workspace,
                    cudaStream_t s);   // gram_tma.cu

namespace {

constexpr int BK = 16, STAGES = 4;
constexpr int SROW = BK + 4;                       // padded smem row, doubles

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;                 // src-size 0 => 16 bytes of zeros
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int BM, int WARPS_M, int WARPS_N>
struct GramCfg {
    static constexpr int THREADS = WARPS_M * WARPS_N * 32;
    static constexpr int FM = BM / WARPS_M / 8, FN = BM / WARPS_N / 8;     // DMMA tiles per warp
    static constexpr int STAGE_DOUBLES = 2 * BM * SROW;
    static constexpr int SMEM_BYTES = STAGES * STAGE_DOUBLES * 8;
    static constexpr int CHUNKS = BM * (BK / 2);                           // 16-byte chunks per operand tile
    static_assert(BM % (WARPS_M * 8) == 0 && BM % (WARPS_N * 8) == 0, "warp tiles are multiples of the 8 x 8 DMMA tile");
    static_assert(CHUNKS % THREADS == 0, "loader assumes an even split of the tile over the threads");
};

// One CTA = one upper-triangular BM x BM tile x one K slice.
template <int BM, int WARPS_M, int WARPS_N, int MIN_CTAS>
__global__ void __launch_bounds__(WARPS_M * WARPS_N * 32, MIN_CTAS)
gram_dmma_kernel(const double* __restrict__ Y0, int64_t ld, int R, int K, int kchunk,
                 const int2* __restrict__ tile_map, int n_tiles, double* __restrict__ ws) {
    using C = GramCfg<BM, WARPS_M, WARPS_N>;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WARPS_N, wn = warp % WARPS_N;
    const int tile = blockIdx.x, split = blockIdx.y;
    const int2 t = tile_map[tile];
    const int rowA0 = t.x * BM, rowB0 = t.y * BM;
    const int k0 = split * kchunk;
    const int k1 = min(k0 + kchunk, K);
    const int nkb = (k1 - k0) / BK;

    auto load_stage = [&](int stage, int kb) {
        double* sA = smem + stage * C::STAGE_DOUBLES;
        double* sB = sA + BM * SROW;
        const int64_t kcol = (int64_t)k0 + (int64_t)kb * BK;
#pragma unroll
        for (int i = 0; i < C::CHUNKS / C::THREADS; ++i) {
            const int c = tid + i * C::THREADS;        // BM rows x 8 chunks of 16 B
            const int r = c >> 3, c2 = (c & 7) * 2;
            const int ga = rowA0 + r, gb = rowB0 + r;
            cp_async16(sA + r * SROW + c2, Y0 + (int64_t)min(ga, R - 1) * ld + kcol + c2, ga < R);
            cp_async16(sB + r * SROW + c2, Y0 + (int64_t)min(gb, R - 1) * ld + kcol + c2, gb < R);
        }
    };

    double acc[C::FM][C::FN][2];
#pragma unroll
    for (int i = 0; i < C::FM; ++i)
#pragma unroll
        for (int j = 0; j < C::FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nkb) load_stage(s, s);
        cp_async_commit();
    }
    const int frag_row = lane >> 2, frag_k = lane & 3;
    for (int kb = 0; kb < nkb; ++kb) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nxt = kb + STAGES - 1;
        if (nxt < nkb) load_stage(nxt % STAGES, nxt);
        cp_async_commit();
        const double* sA = smem + (kb % STAGES) * C::STAGE_DOUBLES + (wm * C::FM * 8 + frag_row) * SROW + frag_k;
        const double* sB = smem + (kb % STAGES) * C::STAGE_DOUBLES + BM * SROW + (wn * C::FN * 8 + frag_row) * SROW + frag_k;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double a[C::FM], b[C::FN];
#pragma unroll
            for (int i = 0; i < C::FM; ++i) a[i] = sA[i * 8 * SROW + kk * 4];
#pragma unroll
            for (int j = 0; j < C::FN; ++j) b[j] = sB[j * 8 * SROW + kk * 4];
#pragma unroll
            for (int i = 0; i < C::FM; ++i)
#pragma unroll
                for (int j = 0; j < C::FN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    double* out = ws + ((int64_t)split * n_tiles + tile) * (BM * BM);
#pragma unroll
    for (int i = 0; i < C::FM; ++i)
#pragma unroll
        for (int j = 0; j < C::FN; ++j) {
            const int r = wm * C::FM * 8 + i * 8 + frag_row;
            const int c = wn * C::FN * 8 + j * 8 + frag_k * 2;
            *reinterpret_cast<double2*>(out + r * BM + c) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
}

// impl 0: 128 x 128 tiles, scalar DFMA, one thread per 8x8 block of the tile.
__global__ void __launch_bounds__(256, 1)
gram_plain_kernel(const double* __restrict__ Y0, int64_t ld, int R, int K, int kchunk,
                  const int2* __restrict__ tile_map, int n_tiles, double* __restrict__ ws) {
    constexpr int BM = 128;
    __shared__ double sA[16][BM + 1], sB[16][BM + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int tile = blockIdx.x, split = blockIdx.y;
    const int2 t = tile_map[tile];
    const int rowA0 = t.x * BM, rowB0 = t.y * BM;
    const int k0 = split * kchunk, k1 = min(k0 + kchunk, K);
    double acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
    for (int kb = k0; kb < k1; kb += 16) {
        for (int e = tid; e < BM * 16; e += 256) {
            const int r = e >> 4, k = e & 15;
            sA[k][r] = (rowA0 + r < R) ? Y0[(int64_t)(rowA0 + r) * ld + kb + k] : 0.0;
            sB[k][r] = (rowB0 + r < R) ? Y0[(int64_t)(rowB0 + r) * ld + kb + k] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            double a[8], b[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = sA[k][ty * 8 + i];
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = sB[k][tx * 8 + j];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    double* out = ws + ((int64_t)split * n_tiles + tile) * (BM * BM);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) out[(ty * 8 + i) * BM + tx * 8 + j] = acc[i][j];
}

// G[i][j] = sum_s ws[s][tile(min,max)][..]; the (i > j) half reads the mirrored element of the same upper tile,
// so G is bitwise symmetric and the summation order over K slices is fixed.
__global__ void __launch_bounds__(256)
gram_reduce_kernel(const double* __restrict__ ws, int R, int bm, int tiles_1d, int n_tiles, int splits,
                   double* __restrict__ G) {
    const int j = blockIdx.x * 32 + (threadIdx.x & 31);
    const int i = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (i >= R || j >= R) return;
    const int lo = min(i, j), hi = max(i, j);
    const int ti = lo / bm, tj = hi / bm;
    const int id = ti * tiles_1d - ti * (ti - 1) / 2 + (tj - ti);
    const int64_t tsz = (int64_t)bm * bm;
    const double* p = ws + (int64_t)id * tsz + (lo % bm) * bm + (hi % bm);
    double s = 0.0;
    for (int k = 0; k < splits; ++k) s += p[(int64_t)k * n_tiles * tsz];
    G[(int64_t)i * R + j] = s;
}

template <int BM, int WARPS_M, int WARPS_N, int MIN_CTAS>
void launch_dmma(const GramPlan& gp, const double* Y0, int64_t ldY0, const int2* tile_map, double* workspace,
                 cudaStream_t s) {
    using C = GramCfg<BM, WARPS_M, WARPS_N>;
    auto kern = gram_dmma_kernel<BM, WARPS_M, WARPS_N, MIN_CTAS>;
    RUVIII_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));  // per device
    dim3 grid(gp.n_tiles, gp.splits);
    kern<<<grid, C::THREADS, C::SMEM_BYTES, s>>>(Y0, ldY0, gp.R, gp.K, gp.kchunk, tile_map, gp.n_tiles, workspace);
}

}  // namespace

// impl: 0 plain DFMA (128), 1 cp.async with the tile shape of least padded work, 2 TMA (128), 4 cp.async 128 forced
GramPlan plan_gram(int R, int K, int n_sms, int impl) {
    GramPlan gp;
    gp.R = R;
    gp.K = K;
    gp.impl = impl;
    auto padded_work = [&](int bm) {
        const int64_t t = ceil_div(R, bm);
        return (double)(t * (t + 1) / 2) * bm * bm;
    };
    gp.bm = 128;
    gp.ctas_per_sm = 1;
    if (impl == 1 && padded_work(80) < 0.97 * padded_work(128)) {      // the 128 tile runs the pipe a little fuller
        gp.bm = 80;
        gp.ctas_per_sm = 2;
    }
    if (impl == 4) gp.impl = 1;
    gp.tiles_1d = ceil_div(R, gp.bm);
    gp.n_tiles = gp.tiles_1d * (gp.tiles_1d + 1) / 2;
    // pick the K split that minimises (waves x per-CTA cost); 192 columns stand for prologue + epilogue
    const int slots = n_sms * gp.ctas_per_sm;
    const int kblocks = K / BK;
    const int max_splits = std::max(1, std::min(kblocks / 8, 64));
    double best = 1e300;
    int best_s = 1;
    for (int s = 1; s <= max_splits; ++s) {
        const int chunk = ceil_div(kblocks, s) * BK;
        const int s_eff = ceil_div(K, chunk);
        const int64_t ctas = (int64_t)gp.n_tiles * s_eff;
        const double waves = (double)((ctas + slots - 1) / slots);
        const double cost = waves * (chunk + 192.0);
        if (cost < best - 1e-9) {
            best = cost;
            best_s = s_eff;
            gp.kchunk = chunk;
        }
    }
    gp.splits = best_s;
    gp.workspace_bytes = (size_t)gp.splits * gp.n_tiles * gp.bm * gp.bm * sizeof(double);
    return gp;
}

void fill_gram_tile_map(const GramPlan& gp, int2* m) {
    int id = 0;
    for (int ti = 0; ti < gp.tiles_1d; ++ti)
        for (int tj = ti; tj < gp.tiles_1d; ++tj) m[id++] = make_int2(ti, tj);
}

int launch_gram(const GramPlan& gp, const double* Y0, int64_t ldY0, const int2* tile_map, double* workspace,
                double* G, const void* tma_desc, cudaStream_t s) {
    if (gp.R <= 0) return 0;
    dim3 grid(gp.n_tiles, gp.splits);
    if (gp.impl == 2) {
        launch_gram_tma(gp, tma_desc, tile_map, workspace, s);
    } else if (gp.impl == 1) {
        if (gp.bm == 80) launch_dmma<80, 2, 2, 2>(gp, Y0, ldY0, tile_map, workspace, s);
        else launch_dmma<128, 2, 4, 1>(gp, Y0, ldY0, tile_map, workspace, s);
    } else {
        gram_plain_kernel<<<grid, 256, 0, s>>>(Y0, ldY0, gp.R, gp.K, gp.kchunk, tile_map, gp.n_tiles, workspace);
    }
    RUVIII_LAUNCH_CHECK();
    dim3 rgrid(ceil_div(gp.R, 32), ceil_div(gp.R, 8));
    gram_reduce_kernel<<<rgrid, 256, 0, s>>>(workspace, gp.R, gp.bm, gp.tiles_1d, gp.n_tiles, gp.splits, G);
    RUVIII_LAUNCH_CHECK();
    return 2;
}

}  // namespace ruviii

// K2, impl 2: TMA (cp.async.bulk.tensor) + mbarrier pipeline feeding the FP64 tensor pipe (DMMA).
//
// Same tiling, split-K and workspace layout as gram.cu impl 1; what changes is how operands reach shared
// memory.  One producer thread issues two TMA box loads per stage (128 rows x 16 doubles = 16 KB each, rows
// beyond R are zero-filled by the TMA unit) into a 6-stage ring guarded by full/empty mbarriers; eight
// consumer warps run DMMA out of the ring.  The boxes are written with the 128-byte swizzle, and the MMA
// row -> tile row assignment interleaves even/odd rows (an 8-row fragment takes rows 0,2,..,14 of a 16-row
// block) so that the four rows a half-warp touches in one 8-byte shared load fall into distinct 32-byte bank
// groups: no padding, no conflicts.  Diagonal tiles load one box and use it for both operands.
//
// FP64 has no tcgen05 path on Blackwell (tcgen05.mma kinds are f16/tf32/f8f6f4/i8/mx*), so the sm_100a
// content of this kernel is the TMA + mbarrier staging; the math is warp-level DMMA.
#include "kernels.cuh"
#include "tma.cuh"

namespace ruviii {

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int STAGES = 6;
constexpr int CONSUMER_WARPS = 8;
constexpr int THREADS = (CONSUMER_WARPS + 4) * 32;       // two consumer warpgroups + one producer warpgroup
constexpr int TILE_BYTES = BM * BK * 8;                 // 16 KB
constexpr int STAGE_BYTES = 2 * TILE_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}
// element (row r, column k) of a 128 x 16-double box stored with the 128-byte swizzle
__device__ __forceinline__ int swz(int r, int k) { return r * 16 + ((((k >> 1) ^ (r & 7)) << 1) | (k & 1)); }

__global__ void __launch_bounds__(THREADS, 1)
gram_tma_kernel(const __grid_constant__ CUtensorMap tmap, int K, int kchunk, const int2* __restrict__ tile_map,
                int n_tiles, double* __restrict__ ws) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(base + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, split = blockIdx.y;
    const int2 t = tile_map[tile];
    const bool diag = t.x == t.y;
    const int k0 = split * kchunk;
    const int k1 = min(k0 + kchunk, K);
    const int nkb = (k1 - k0) / BK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // Registers are allocated per warpgroup (4 warps): 12 warps x 168 registers fill the file, so the producer
    // warpgroup hands its registers back and the two consumer warpgroups grow to what 64 accumulators need.
    if (warp >= CONSUMER_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        // ===== producer: one thread drives the TMA unit =====
        if (warp == CONSUMER_WARPS && lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);               // first pass over the ring returns immediately
                mbar_expect_tx(&full[s], diag ? TILE_BYTES : STAGE_BYTES);
                unsigned char* sA = base + s * STAGE_BYTES;
                tma_load_2d(sA, &tmap, k0 + kb * BK, t.x * BM, &full[s]);
                if (!diag) tma_load_2d(sA + TILE_BYTES, &tmap, k0 + kb * BK, t.y * BN, &full[s]);
            }
        }
        return;
    }

    // ===== consumers: 8 warps as 2 (rows) x 4 (cols), warp tile 64 x 32 =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
    const int wm = warp >> 2, wn = warp & 3;
    const int g = lane >> 2, kq = lane & 3;
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // tile row of fragment i (A) / j (B): interleaved even/odd rows inside 16-row blocks
    int rowA[8], rowB[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) rowA[i] = wm * 64 + (i >> 1) * 16 + 2 * g + (i & 1);
#pragma unroll
    for (int j = 0; j < 4; ++j) rowB[j] = wn * 32 + (j >> 1) * 16 + 2 * g + (j & 1);

    for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full[s], ph);
        const double* sA = reinterpret_cast<const double*>(base + s * STAGE_BYTES);
        const double* sB = diag ? sA : sA + TILE_BYTES / 8;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = sA[swz(rowA[i], kk * 4 + kq)];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[swz(rowB[j], kk * 4 + kq)];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    // accumulator (i, j, e): row rowA[i]; column = tile row of B fragment j at MMA column n = 2*kq + e
    double* out = ws + ((int64_t)split * n_tiles + tile) * (BM * BN);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = wn * 32 + jj * 16 + 2 * (2 * kq + e);
                *reinterpret_cast<double2*>(out + rowA[i] * BN + c) = make_double2(acc[i][2 * jj][e], acc[i][2 * jj + 1][e]);
            }
}

}  // namespace

// Encode the 2-D tensor map of Y0 (inner dimension = padded gene count, outer = R rows) into desc_out (128 B, host).
int launch_gram_tma_prepare(const GramPlan& gp, const double* Y0, int64_t ldY0, void* desc_out) {
    encode_tensor_map_2d(desc_out, Y0, (uint64_t)ldY0, (uint64_t)gp.R, (uint64_t)ldY0, BK, BM, /*swizzle128=*/true);
    return 0;
}

int launch_gram_tma(const GramPlan& gp, const void* tma_desc, const int2* tile_map, double* workspace, cudaStream_t s) {
    CUtensorMap m;
    std::memcpy(&m, tma_desc, sizeof(m));
    RUVIII_CUDA_CHECK(cudaFuncSetAttribute(gram_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    dim3 grid(gp.n_tiles, gp.splits);
    gram_tma_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(m, gp.K, gp.kchunk, tile_map, gp.n_tiles, workspace);
    RUVIII_LAUNCH_CHECK();
    return 1;
}

}  // namespace ruviii

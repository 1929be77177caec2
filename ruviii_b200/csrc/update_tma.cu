// K6, impl 2: the rank-k update with TMA-staged tiles of Y.
//
// Same arithmetic as update.cu (newY_k = Y - sum_{q<k} A~[:, q] (x) a~[q, :], ruvIII.R:145).  What changes is how Y
// reaches the SM: a producer warp streams 4-row x 192-gene boxes (6 KB) through a 6-deep shared-memory ring with
// cp.async.bulk.tensor + mbarrier, so 36 KB per CTA (4 CTAs per SM) are in flight regardless of how many registers the consumer
// threads need for their 2 x kk entries of a~.  Consumers read their 16 bytes per row conflict-free from the ring,
// release the slot as soon as the values sit in registers, and write results straight to HBM with streaming
// stores.  Columns and rows outside the matrix are zero-filled by the TMA unit and never stored.
#include "kernels.cuh"
#include "tma.cuh"

namespace ruviii {

namespace {

constexpr int TR = 4;                    // rows per stage
constexpr int TG = 192;                  // genes per stage (two per consumer thread)
constexpr int STAGES = 6;
constexpr int CONSUMERS = 96;            // 3 consumer warps + 1 producer warp = one register-allocation unit of 4 warps
constexpr int THREADS = CONSUMERS + 32;
constexpr int STAGE_BYTES = TR * TG * 8; // 6 KB
constexpr int ROWS_PER_CTA = 128;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 2 * STAGES * 8;

struct KSlots { signed char slot[kMaxK + 1]; };

__device__ __forceinline__ double2 lds_v2(const void* p) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(smem_u32(p)));
    return v;
}

template <int KP>
__global__ void __launch_bounds__(THREADS, KP <= 16 ? 4 : 2)
update_tma_kernel(const __grid_constant__ CUtensorMap tmap, int rows, int n_pairs, const double* __restrict__ At,
                  const double* __restrict__ at, int64_t ldat, int kk, uint64_t kmask, KSlots slots,
                  double* __restrict__ out, int64_t ldout, int64_t out_stride) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(base + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    __shared__ __align__(16) double sAt[ROWS_PER_CTA][KP];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int col0 = blockIdx.x * TG;
    const int row0 = blockIdx.y * ROWS_PER_CTA;
    const int nrows = min(ROWS_PER_CTA, rows - row0);
    const int ntiles = (nrows + TR - 1) / TR;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CONSUMERS / 32);
        }
        mbar_fence_init();
    }
    for (int e = tid; e < ROWS_PER_CTA * KP; e += THREADS) {
        const int r = e / KP;
        sAt[r][e - r * KP] = r < nrows ? At[(int64_t)(row0 + r) * KP + (e - r * KP)] : 0.0;
    }
    __syncthreads();

    if (warp == CONSUMERS / 32) {
        if (lane == 0) {
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % STAGES;
                mbar_wait(&empty[s], ((t / STAGES) & 1) ^ 1);
                mbar_expect_tx(&full[s], STAGE_BYTES);
                tma_load_2d(base + s * STAGE_BYTES, &tmap, col0, row0 + t * TR, &full[s]);
            }
        }
        return;
    }

    const int pair = blockIdx.x * CONSUMERS + tid;
    const bool active = pair < n_pairs;
    const int64_t col = 2 * (int64_t)pair;
    double2 a[KP];
#pragma unroll
    for (int q = 0; q < KP; ++q)
        a[q] = (active && q < kk) ? *reinterpret_cast<const double2*>(at + (int64_t)q * ldat + col) : make_double2(0.0, 0.0);
    const bool want0 = kmask & 1ull;

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % STAGES;
        mbar_wait(&full[s], (t / STAGES) & 1);
        const unsigned char* st = base + s * STAGE_BYTES + tid * 16;
        double2 y[TR];
#pragma unroll
        for (int u = 0; u < TR; ++u) y[u] = lds_v2(st + u * (TG * 8));
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (!active) continue;
        const int rb = t * TR;
        if (want0) {
            double* o = out + (int64_t)slots.slot[0] * out_stride + col;
#pragma unroll
            for (int u = 0; u < TR; ++u)
                if (rb + u < nrows) st_stream(o + (int64_t)(row0 + rb + u) * ldout, y[u]);
        }
#pragma unroll
        for (int q = 0; q < KP; q += 2) {
            if (q < kk) {
                double2 w[TR];
#pragma unroll
                for (int u = 0; u < TR; ++u) w[u] = lds_v2(&sAt[rb + u][q]);
#pragma unroll
                for (int u = 0; u < TR; ++u) {
                    y[u].x = fma(-w[u].x, a[q].x, y[u].x);
                    y[u].y = fma(-w[u].x, a[q].y, y[u].y);
                }
                if ((kmask >> (q + 1)) & 1ull) {
                    double* o = out + (int64_t)slots.slot[q + 1] * out_stride + col;
#pragma unroll
                    for (int u = 0; u < TR; ++u)
                        if (rb + u < nrows) st_stream(o + (int64_t)(row0 + rb + u) * ldout, y[u]);
                }
                if (q + 1 < kk) {
#pragma unroll
                    for (int u = 0; u < TR; ++u) {
                        y[u].x = fma(-w[u].y, a[q + 1].x, y[u].x);
                        y[u].y = fma(-w[u].y, a[q + 1].y, y[u].y);
                    }
                    if ((kmask >> (q + 2)) & 1ull) {
                        double* o = out + (int64_t)slots.slot[q + 2] * out_stride + col;
#pragma unroll
                        for (int u = 0; u < TR; ++u)
                            if (rb + u < nrows) st_stream(o + (int64_t)(row0 + rb + u) * ldout, y[u]);
                    }
                }
            }
        }
    }
}

}  // namespace

void update_tma_prepare(const double* Y, int64_t ldY, int rows, void* desc_out) {
    encode_tensor_map_2d(desc_out, Y, (uint64_t)ldY, (uint64_t)std::max(rows, 1), (uint64_t)ldY, TG, TR, /*swizzle128=*/false);
}

void launch_update_tma(const void* tmap_host, int rows, int n, const double* At_rows, int kp, const double* at,
                       int64_t ldat, int kk, uint64_t kmask, const int* out_slot_of_k, double* out, int64_t ldout,
                       int64_t out_stride, cudaStream_t s) {
    if (rows <= 0 || n <= 0) return;
    CUtensorMap m;
    std::memcpy(&m, tmap_host, sizeof(m));
    KSlots slots;
    for (int i = 0; i <= kMaxK; ++i) slots.slot[i] = (signed char)out_slot_of_k[i];
    const int n_pairs = (n + 1) / 2;
    dim3 grid(ceil_div(n_pairs, CONSUMERS), ceil_div(rows, ROWS_PER_CTA));
#define RUVIII_UPD(KP)                                                                                               \
    do {                                                                                                             \
        RUVIII_CUDA_CHECK(cudaFuncSetAttribute(update_tma_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                               SMEM_BYTES));                                                         \
        update_tma_kernel<KP><<<grid, THREADS, SMEM_BYTES, s>>>(m, rows, n_pairs, At_rows, at, ldat, kk, kmask, slots, \
                                                                out, ldout, out_stride);                             \
    } while (0)
    switch (kp) {
        case 4: RUVIII_UPD(4); break;
        case 8: RUVIII_UPD(8); break;
        case 16: RUVIII_UPD(16); break;
        default: RUVIII_UPD(32); break;
    }
#undef RUVIII_UPD
    RUVIII_LAUNCH_CHECK();
}

}  // namespace ruviii

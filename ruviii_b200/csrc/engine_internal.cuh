// Internal definitions shared by the engine's translation units (engine.cu, engine_pipeline.cu, engine_aux.cu):
// the lazily loaded NCCL entry points, device buffers, the per-GPU shard state and the handle behind the C ABI.
#pragma once
#include "../../include/ruviii.h"
#include "kernels.cuh"

#include <cublas_v2.h>
#include <cuda.h>
#include <cusolverDn.h>
#include <dlfcn.h>
#include <nccl.h>   // types only: the library is dlopen()ed, see NcclApi
#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges cost nothing unless a profiler is attached

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

namespace ruviii {

int launch_gram_tma_prepare(const GramPlan& gp, const double* Y0, int64_t ldY0, void* desc_out);  // gram_tma.cu

// NCCL entry points, resolved at run time.  If the process already holds a libnccl.so.2 (torch imports its own
// 2.28 build) that one is reused; otherwise the system library is loaded.  Linking libnccl directly would let
// whichever of the two builds is loaded first shadow the other under the shared soname.
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    static NcclApi& get() {
        static NcclApi api;
        static std::once_flag once;
        static std::string err;
        std::call_once(once, [] {
            void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
            if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
            if (!lib) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
            auto sym = [&](const char* name) {
                void* p = dlsym(lib, name);
                if (!p && err.empty()) err = std::string("libnccl lacks ") + name;
                return p;
            };
            api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
            api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
            api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
            api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
            api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
            api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
            api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
            api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
            api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
            api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
            api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
            api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        });
        if (!err.empty()) throw ::ruviii::Error(3, err);
        return api;
    }
};
#define ncclGetUniqueId ::ruviii::NcclApi::get().GetUniqueId
#define ncclCommInitRank ::ruviii::NcclApi::get().CommInitRank
#define ncclCommInitAll ::ruviii::NcclApi::get().CommInitAll
#define ncclCommDestroy ::ruviii::NcclApi::get().CommDestroy
#define ncclAllReduce ::ruviii::NcclApi::get().AllReduce
#define ncclBroadcast ::ruviii::NcclApi::get().Broadcast
#define ncclAllGather ::ruviii::NcclApi::get().AllGather
#define ncclSend ::ruviii::NcclApi::get().Send
#define ncclRecv ::ruviii::NcclApi::get().Recv
#define ncclGroupStart ::ruviii::NcclApi::get().GroupStart
#define ncclGroupEnd ::ruviii::NcclApi::get().GroupEnd
#define ncclGetErrorString ::ruviii::NcclApi::get().GetErrorString

#define RUVIII_NCCL_CHECK(expr)                                                                         \
    do {                                                                                                \
        ncclResult_t _r = (expr);                                                                       \
        if (_r != ncclSuccess)                                                                          \
            throw ::ruviii::Error(RUVIII_ERR_NCCL, std::string(#expr) + ": " + ncclGetErrorString(_r)); \
    } while (0)
#define RUVIII_SOLVER_CHECK(expr)                                                                        \
    do {                                                                                                 \
        cusolverStatus_t _r = (expr);                                                                    \
        if (_r != CUSOLVER_STATUS_SUCCESS)                                                               \
            throw ::ruviii::Error(RUVIII_ERR_SOLVER, std::string(#expr) + ": cusolver status " +         \
                                                         std::to_string((int)_r));                       \
    } while (0)

// NVTX range around the host-side enqueue of one stage of the pass (K1..K6 of DESIGN.md): a timeline tool (Nsight Systems)
// shows the stages by name above the kernels they launch.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

inline int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v && *v ? std::atoi(v) : dflt;
}

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t count = 0;
    bool owned = true;            // false: p points into memory owned by somebody else (the peer-mapped exchange buffer)
    void alias(T* q, size_t n) {
        release();
        p = q;
        count = n;
        owned = false;
    }
    // The clear runs on the legacy default stream, which does NOT order against the engine's non-blocking streams: wait for
    // it, otherwise an asynchronous copy queued right after alloc() can be overwritten by the late memset.
    static void clear(void* q, size_t bytes) {
        RUVIII_CUDA_CHECK(cudaMemset(q, 0, bytes));
        RUVIII_CUDA_CHECK(cudaStreamSynchronize(cudaStreamLegacy));
    }
    void alloc(size_t n, bool zero = true) {
        if (!owned) { p = nullptr; count = 0; owned = true; }
        if (n <= count && p) {
            if (zero) clear(p, n * sizeof(T));
            return;
        }
        release();
        if (n == 0) return;
        RUVIII_CUDA_CHECK(cudaMalloc(&p, n * sizeof(T)));
        count = n;
        if (zero) clear(p, n * sizeof(T));
    }
    void release() {
        if (p && owned) cudaFree(p);
        p = nullptr;
        count = 0;
        owned = true;
    }
};

struct StageState;
void stage_state_destroy(StageState* st);
struct P2PState;
void p2p_state_destroy(P2PState* p);

struct Shard {
    int device = 0;
    int g0 = 0, n = 0;            // first gene (in the caller's matrix) and gene count of this shard
    int64_t ld = 0;               // device row pitch (doubles) of every genes-contiguous matrix
    int n_ctl = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr, d2h_stream = nullptr;
    std::vector<cudaEvent_t> ev_chunk, ev_upd, ev_d2h;   // host-to-host paths: chunk uploaded / updated / downloaded
    std::vector<int> ctl_idx_host;               // this shard's control-gene columns (local indices), host copy
    ncclComm_t comm = nullptr;
    std::vector<cudaEvent_t> ev;
    cudaEvent_t ev_fork = nullptr, ev_gather = nullptr;   // main -> aux stream fork, aux -> main join (control-gene gather)
    cudaEvent_t ev_acgram = nullptr;                      // aux -> main join of ac %*% t(ac)
    cudaEvent_t ev_g0 = nullptr, ev_g1 = nullptr;         // timing of the control-gene gather on the auxiliary stream
    int64_t ldc = 0;
    DevBuf<double> Y, PS, Y0, G, gram_ws, Vd, U, evals, part, alpha, acT, AG, Linv, colmean, At, at, out, out_ps, W,
        evals_asc, falpha, topk_buf, Zt, avg, Yc, Zstage;
    DevBuf<double> solve_ws;                    // partial column sums + barrier counters of the fused solve kernel
    DevBuf<double> eta, eta_t, etacT, AG1, At1, Linv1, cm1, Yc1, PS1;   // ruv::RUV1 pre-step (eta != NULL)
    DevBuf<double> pX, pG, pWs, pU, pEv, pPart, pAlpha, pStat, pMean, pSd, pTopk, pEvAsc;   // computePCA fast path (ruviii_pca)
    DevBuf<double> gX, gPart, gCov, gF, gR;     // gene statistics (ruviii_gene_stats)
    DevBuf<int> gStart, gRows, gGroup;
    DevBuf<double> rD, rW, rQ, rPart, rAlpha, gReg, gReg2;   // regress-out (ruviii_regress_out): design, basis, residuals
    DevBuf<int> rRank;
    int reg_g0 = 0, reg_n = 0;                  // genes [reg_g0, reg_g0 + reg_n) of the regressed matrix live in gReg
    int64_t reg_ld = 0;
    DevBuf<int2> pTiles;
    DevBuf<int> pState;
    TopkWork ptopk{};
    DevBuf<int2> tile_map;
    DevBuf<int> ctl_idx, grp_start, grp_rows, status, topk_state, all_start, all_rows, status1, set_start, set_rows;
    TopkWork topk{};
    double* persist_ws = nullptr;          // workspace of the persistent eigensolver (inside topk_buf)
    alignas(64) unsigned char tmap[128];   // host copy of the CUtensorMap of Y0 (gram impl 2)
    alignas(64) unsigned char tmap_y[128]; // host copy of the CUtensorMap of Y (update impl 2)
    GramPlan gp;
    int rsplit = 1;
};

constexpr int kHostFlagInts = 64 + 16 * 16;
enum Ev { E_START, E_LOG, E_RESIDOP, E_GRAM, E_AR1, E_EIG, E_BCAST, E_PROJECT, E_CTL, E_AR2, E_SOLVE, E_UPDATE, E_COUNT };

}  // namespace ruviii

struct ruviii_handle {
    std::vector<ruviii::Shard> shards;
    int rank0 = 0, world = 1;           // global rank of shards[0], total shards over all processes
    bool multi_process = false;
    std::string err;
    int n_sms = 148;

    // solver state lives on the device of shards[0] (used only when rank0 == 0)
    cusolverDnHandle_t solver = nullptr;
    cublasHandle_t blas = nullptr;
    ruviii::DevBuf<double> solver_work;
    ruviii::DevBuf<int> solver_info;
    int* h_flags = nullptr;             // pinned: [0] cusolver devInfo, [1..] cholesky status per shard, [60..63] multi-kernel
                                        // eigensolver state, [64 + 16 i ..] persistent eigensolver state of shard i
    double* h_dflag = nullptr;          // pinned: [i] = allreduced "some rank's eigensolver did not converge" flag of shard i

    // plan
    bool planned = false, uploaded = false, ran = false, logged = false, have_user_alpha = false;
    ruviii_params prm{};
    int m1 = 0, n_ps = 0, m = 0, n_local_total = 0;
    int p = 0, R = 0, r = 0, n_groups_active = 0, max_group = 0;
    int active_rows = 0;                // rows in non-singleton groups; R = active_rows - n_groups_active (Helmert rows)
    int64_t n_ctl_global = 0;
    std::vector<int> ks, k_eff;
    int kmax = 0, kk = 0, kp = 4;       // kk = min(kmax, r); kp = padded register extent
    int n_alpha_rows = 0, n_keep = 0;   // rows of fullalpha produced / eigenvectors kept
    int rows_w = 0, mu_rows = 0;
    bool no_replicates = false;
    uint64_t kmask = 0;
    int slot_of_k[ruviii::kMaxK + 1];
    std::vector<std::pair<int, int>> dup_slots;   // (dst slot, src slot) for k values with the same effective k
    int launches = 0;
    ruviii_info info{};
    int upd_impl = 3;
    int gram_impl = 1, eig_impl = 0;   // eig_impl: 0 auto, 1 syevd, 2 syevdx, 3 subspace iteration only
    int rank_bound = 0, eig_solver_used = 0, eig_iterations = 0;
    bool use_topk = false;
    int persist_rb = 0;                 // rows of G per CTA of the persistent eigensolver; 0 = multi-kernel form / cuSOLVER
    bool active_all_ps = false;         // every non-singleton row is a PRPS row (the PRPS regime)
    int pipelined_last = 0;             // 1 when the last ruviii_normalise took the pipelined host path
    bool gene_side = false;             // Gram = Z'Z (n x n) instead of Z Z' (R x R): taken when R > n
    int gorder = 0;                     // order of the Gram that is formed and decomposed
    // gene-side Gram over several shards: gene count / first gene of every GLOBAL rank, and the split of
    // the R Helmert rows over the ranks after the all-to-all (rank j contracts rows [zrow[j], zrow[j+1]))
    std::vector<int> gn, goff, zrow;
    int n_total = 0;                    // genes over all shards of all processes
    // ruv::RUV1 pre-step (ruvIII.R:116): eta as handed over by ruviii_set_eta (q x n row-major, host copy), sticky
    std::vector<double> eta_host;
    int eta_q = 0, eta_n = 0;
    bool eta_applied = false;           // once per upload, like the log transform
    // PRPS construction on the device (prpsForCategoricalUV.R:123-139): CSR of the sample sets handed over by
    // ruviii_set_prps_sets (sticky); replicate.data is then built from the uploaded assay when PS == NULL
    std::vector<int> prps_start, prps_rows;
    int prps_n = 0;
    bool ps_pending = false;            // the PRPS rows still have to be built from the resident assay
    bool ps_from_sets = false;          // the resident PRPS rows come from ruviii_set_prps_sets, not from the host
    bool reg_valid = false;             // gReg holds the residuals of the last ruviii_regress_out (slot = -2 of the statistics)
    int reg_m1 = 0, reg_n = 0;
    bool yc_valid = false;              // the control-gene block Yc holds Y[, ctl] of the CURRENT upload (gathered once per upload)
    double eig_tol = 1e-13;
    ruviii::P2PState* p2p = nullptr;       // peer-memory exchange buffers of the one-process-per-GPU allreduce (p2p.cu)
    ruviii::StageState* stage = nullptr;   // pool of host threads + pinned bounce buffers of the staged host path

    ~ruviii_handle();
};

extern thread_local std::string g_create_error;

namespace ruviii {

// engine.cu
void allreduce_sum(ruviii_handle* h, std::vector<double*> bufs, size_t count);
void broadcast0(ruviii_handle* h, std::vector<double*> bufs, size_t count);
void spin_sync(cudaStream_t st);
void sync_all(ruviii_handle* h);
void do_plan(ruviii_handle* h, const ruviii_params* prm, int m1, int n, int n_ps, const int32_t* group_of_row,
             const uint8_t* ctl, const int32_t* ks, int nk);
void ensure_logged(ruviii_handle* h);
void do_upload(ruviii_handle* h, const double* Y, int64_t ldY, const double* PS, int64_t ldPS);
void eigensolve(ruviii_handle* h, bool allow_subspace = true);
void do_run(ruviii_handle* h, ruviii_info* info_out);
void launch_solve(ruviii_handle* h, Shard& s);      // K5c + A~ + a~ on the shard's main stream
void launch_w_blocks(ruviii_handle* h, Shard& s);   // W_k for every requested k (only when W is downloaded)
void do_download(ruviii_handle* h, double* newY, double* newY_ps, double* fullalpha, double* W, double* const* newY_blocks = nullptr);
// engine_pipeline.cu
bool can_pipeline(ruviii_handle* h, const double* Y, const double* PS, const double* newY_out, const double* newY_ps_out);
void do_normalise_pipelined(ruviii_handle* h, const double* Y, int64_t ldY, const double* PS, int64_t ldPS, double* newY_out,
                            double* fullalpha_out, double* W_out);
// p2p.cu
bool p2p_setup(ruviii_handle* h, size_t inbox_doubles, size_t mat_doubles);
double* p2p_carve(ruviii_handle* h, size_t count);
bool p2p_allreduce(ruviii_handle* h, double* buf, size_t count);
void p2p_check(ruviii_handle* h);
// engine_stage.cu
bool can_stage(ruviii_handle* h, const double* PS, const double* newY_ps_out);
void do_normalise_staged(ruviii_handle* h, const double* Y, int64_t ldY, const double* PS, int64_t ldPS, double* const* out_blocks,
                         double* fullalpha_out, double* W_out);
// engine_aux.cu
void apply_ruv1(ruviii_handle* h);
void do_pca(ruviii_handle* h, const double* X, int64_t ldX, int m1, int n, int slot, int apply_log, double pseudo_count,
            int center, int scale, int k, double* u_out, double* d_out, double* v_out, float* stage_ms);
void do_regress_out(ruviii_handle* h, const double* X, int64_t ldX, int m1, int n, int slot, int apply_log, double pseudo_count,
                    const double* design, int p, double* out, int64_t ld_out, int* rank_out, float* ms_out);
void do_gene_stats(ruviii_handle* h, const double* X, int64_t ldX, int m1, int n, int slot, int apply_log, double pseudo_count,
                   const int32_t* group, const double* covariate, double* F_out, double* r_out, float* ms_out);

// After a failure nothing of this handle may still be running when the caller touches (or frees) its buffers again.  Only
// for handles that own all their shards: with one process per GPU a peer may be sitting in a collective this rank never
// reached, and waiting for the device would turn an error return into a hang.
inline void quiesce(ruviii_handle* h) {
    if (!h || h->multi_process) return;
    for (auto& s : h->shards) {
        if (cudaSetDevice(s.device) == cudaSuccess) cudaDeviceSynchronize();
        cudaGetLastError();
    }
}

template <class F>
int guarded(ruviii_handle* h, F&& fn) {
    try {
        fn();
        return RUVIII_OK;
    } catch (const Error& e) {
        if (h) h->err = e.what(); else g_create_error = e.what();
        cudaGetLastError();
        quiesce(h);
        return e.code;
    } catch (const std::exception& e) {
        if (h) h->err = e.what(); else g_create_error = e.what();
        quiesce(h);
        return RUVIII_ERR_INVALID;
    }
}

}  // namespace ruviii

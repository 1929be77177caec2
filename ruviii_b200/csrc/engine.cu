// Engine: handle, plan, gene-sharded pipeline, NCCL exchange steps, C ABI (include/ruviii.h).
//
// Pipeline.  Every GPU ("shard") owns a contiguous gene range of Y, PS, ctl and of
// every output; the only exchanges are two small allreduces and one broadcast:
//   K1 residop -> K2 partial Gram -> allreduce(G) -> K3 eigensolve on global rank 0 -> bcast(U)
//   -> K4 projection -> K5 control-gene products -> allreduce([A | Gm]) -> Cholesky -> K6 update.
// One process may drive several shards (an R session: ruviii_create) or exactly one (torchrun:
// ruviii_create_rank); the code below is the same loop over `shards_` in both cases.
// Companion translation units: engine_pipeline.cu (overlapped host-to-host path of ruviii_normalise), engine_aux.cu (RUV1
// pre-step, PRPS construction, gene statistics, fast PCA, probes); shared definitions in engine_internal.cuh.
#include "engine_internal.cuh"

using namespace ruviii;

thread_local std::string g_create_error;

ruviii_handle::~ruviii_handle() {
    for (auto& s : shards) {
        cudaSetDevice(s.device);
        cudaDeviceSynchronize();
        if (s.comm) ncclCommDestroy(s.comm);
        for (auto e : s.ev) cudaEventDestroy(e);
        if (s.ev_fork) cudaEventDestroy(s.ev_fork);
        if (s.ev_gather) cudaEventDestroy(s.ev_gather);
        if (s.ev_acgram) cudaEventDestroy(s.ev_acgram);
        if (s.ev_g0) cudaEventDestroy(s.ev_g0);
        if (s.ev_g1) cudaEventDestroy(s.ev_g1);
        DevBuf<double>* bufs[] = {&s.Y, &s.PS, &s.Y0, &s.G, &s.gram_ws, &s.Vd, &s.U, &s.evals, &s.part, &s.alpha, &s.acT,
                                  &s.AG, &s.Linv, &s.colmean, &s.At, &s.at, &s.out, &s.out_ps, &s.W, &s.evals_asc,
                                  &s.falpha, &s.topk_buf, &s.Zt, &s.avg, &s.Yc, &s.Zstage, &s.solve_ws,
                                  &s.eta, &s.eta_t, &s.etacT, &s.AG1, &s.At1, &s.Linv1, &s.cm1, &s.Yc1, &s.PS1,
                                  &s.pX, &s.pG, &s.pWs, &s.pU, &s.pEv, &s.pPart, &s.pAlpha, &s.pStat, &s.pMean, &s.pSd, &s.pTopk,
                                  &s.pEvAsc, &s.gX, &s.gPart, &s.gCov, &s.gF, &s.gR};
        for (auto b : bufs) b->release();
        s.tile_map.release();
        s.ctl_idx.release();
        s.grp_start.release();
        s.grp_rows.release();
        s.status.release();
        s.topk_state.release();
        s.all_start.release();
        s.all_rows.release();
        s.status1.release();
        s.set_start.release();
        s.set_rows.release();
        s.pTiles.release();
        s.pState.release();
        s.gStart.release();
        s.gRows.release();
        s.gGroup.release();
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.copy_stream) cudaStreamDestroy(s.copy_stream);
        if (s.d2h_stream) cudaStreamDestroy(s.d2h_stream);
        for (auto e : s.ev_chunk) cudaEventDestroy(e);
        for (auto e : s.ev_upd) cudaEventDestroy(e);
        for (auto e : s.ev_d2h) cudaEventDestroy(e);
    }
    if (!shards.empty()) cudaSetDevice(shards[0].device);
    solver_work.release();
    solver_info.release();
    if (solver) cusolverDnDestroy(solver);
    if (blas) cublasDestroy(blas);
    if (h_flags) cudaFreeHost(h_flags);
    if (h_dflag) cudaFreeHost(h_dflag);
    if (stage) ruviii::stage_state_destroy(stage);
    if (p2p) { if (!shards.empty()) cudaSetDevice(shards[0].device); ruviii::p2p_state_destroy(p2p); }
}

namespace ruviii {

namespace {

void check_device(int dev) {
    cudaDeviceProp prop;
    RUVIII_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        throw Error(RUVIII_ERR_CUDA, "device " + std::to_string(dev) + " (" + prop.name + ", sm_" +
                                         std::to_string(prop.major) + std::to_string(prop.minor) +
                                         ") is not a Blackwell sm_100 GPU; this library has no other code path");
}

void init_shard(ruviii_handle* h, Shard& s, int device) {
    s.device = device;
    check_device(device);
    RUVIII_CUDA_CHECK(cudaSetDevice(device));
    // The main stream carries the critical path (including the tiny latency-bound eigensolver kernels); the auxiliary
    // streams carry the control-gene gather and the copies.  Priorities make the block scheduler hand freed SM slots to
    // the main stream first, otherwise the 6,000-CTA gather delays every few-CTA eigensolver launch behind it.
    int prio_lo = 0, prio_hi = 0;
    RUVIII_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    RUVIII_CUDA_CHECK(cudaStreamCreateWithPriority(&s.stream, cudaStreamNonBlocking, prio_hi));
    RUVIII_CUDA_CHECK(cudaStreamCreateWithPriority(&s.copy_stream, cudaStreamNonBlocking, prio_lo));
    RUVIII_CUDA_CHECK(cudaStreamCreateWithPriority(&s.d2h_stream, cudaStreamNonBlocking, prio_lo));
    s.ev.resize(E_COUNT);
    for (auto& e : s.ev) RUVIII_CUDA_CHECK(cudaEventCreate(&e));
    RUVIII_CUDA_CHECK(cudaEventCreateWithFlags(&s.ev_fork, cudaEventDisableTiming));
    RUVIII_CUDA_CHECK(cudaEventCreateWithFlags(&s.ev_gather, cudaEventDisableTiming));
    RUVIII_CUDA_CHECK(cudaEventCreateWithFlags(&s.ev_acgram, cudaEventDisableTiming));
    RUVIII_CUDA_CHECK(cudaEventCreate(&s.ev_g0));
    RUVIII_CUDA_CHECK(cudaEventCreate(&s.ev_g1));
    // RUVIII_L2_GRAN=32/64/128: cap on the DRAM -> L2 fetch size (the control-gene gather uses one 8-byte element per sector)
    if (const int gran = env_int("RUVIII_L2_GRAN", 0)) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    h->n_sms = sms;
}

void finish_create(ruviii_handle* h) {
    RUVIII_CUDA_CHECK(cudaSetDevice(h->shards[0].device));
    RUVIII_CUDA_CHECK(cudaHostAlloc((void**)&h->h_flags, kHostFlagInts * sizeof(int), cudaHostAllocDefault));
    std::memset(h->h_flags, 0, kHostFlagInts * sizeof(int));
    RUVIII_CUDA_CHECK(cudaHostAlloc((void**)&h->h_dflag, 64 * sizeof(double), cudaHostAllocDefault));
    std::memset(h->h_dflag, 0, 64 * sizeof(double));
    h->gram_impl = env_int("RUVIII_GRAM_IMPL", 1);
    h->eig_impl = env_int("RUVIII_EIG_IMPL", 0);
    h->upd_impl = env_int("RUVIII_UPD_IMPL", 3);   // 1 a~ in registers, 2 TMA ring, 3 a~ in shared memory (default)
    if (const char* t = std::getenv("RUVIII_EIG_TOL")) h->eig_tol = std::atof(t);
}

}  // namespace

// ---- collectives over all shards of all processes ------------------------------------------------------
void allreduce_sum(ruviii_handle* h, std::vector<double*> bufs, size_t count) {
    if (h->world == 1) return;
    // one process per GPU: the exchange runs as one kernel over NVLink peer memory (p2p.cu) when the buffers are mapped
    if (h->multi_process && p2p_allreduce(h, bufs[0], count)) { h->launches += 1; return; }
    RUVIII_NCCL_CHECK(ncclGroupStart());
    for (size_t i = 0; i < h->shards.size(); ++i)
        RUVIII_NCCL_CHECK(ncclAllReduce(bufs[i], bufs[i], count, ncclDouble, ncclSum, h->shards[i].comm,
                                        h->shards[i].stream));
    RUVIII_NCCL_CHECK(ncclGroupEnd());
}
void broadcast0(ruviii_handle* h, std::vector<double*> bufs, size_t count) {
    if (h->world == 1) return;
    RUVIII_NCCL_CHECK(ncclGroupStart());
    for (size_t i = 0; i < h->shards.size(); ++i)
        RUVIII_NCCL_CHECK(ncclBroadcast(bufs[i], bufs[i], count, ncclDouble, 0, h->shards[i].comm, h->shards[i].stream));
    RUVIII_NCCL_CHECK(ncclGroupEnd());
}

// Wait for a stream by polling instead of cudaStreamSynchronize: the driver's blocking wait puts a host thread that has
// waited for more than a moment to sleep, and its wake-up latency (measured: 0.1 - 3 ms per pass on an 8-rank run,
// showing up as the other ranks waiting in the first allreduce) is larger than the whole 1 ms pass.  Spin for up to
// 200 ms, then fall back to the blocking call.
void spin_sync(cudaStream_t st) {
    const auto t0 = std::chrono::steady_clock::now();
    for (int it = 0;; ++it) {
        const cudaError_t q = cudaStreamQuery(st);
        if (q == cudaSuccess) return;
        if (q != cudaErrorNotReady) RUVIII_CUDA_CHECK(q);
        if ((it & 1023) == 1023 &&
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 0.2) break;
    }
    RUVIII_CUDA_CHECK(cudaStreamSynchronize(st));
}

void sync_all(ruviii_handle* h) {
    for (auto& s : h->shards) {
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        spin_sync(s.stream);
        spin_sync(s.copy_stream);
        spin_sync(s.d2h_stream);
    }
}

// ---- plan ----------------------------------------------------------------------------------------------
void do_plan(ruviii_handle* h, const ruviii_params* prm, int m1, int n, int n_ps, const int32_t* group_of_row,
             const uint8_t* ctl, const int32_t* ks, int nk) {
    if (!prm || m1 < 0 || n <= 0 || n_ps < 0 || nk <= 0 || !ks || !ctl)
        throw Error(RUVIII_ERR_INVALID, "ruviii_plan: null or non-positive argument");
    if (prm->mode != RUVIII_MODE_STACKED && prm->mode != RUVIII_MODE_FAST)
        throw Error(RUVIII_ERR_INVALID, "ruviii_plan: unknown mode");
    if (prm->average_rep && prm->mode == RUVIII_MODE_FAST && m1 != n_ps)
        throw Error(RUVIII_ERR_INVALID, "apply.average.rep in fastruvIII multiplies t(M) (n_ps rows) with newY (m1 rows): "
                                        "non-conformable arguments (fastruvIII.R:179-180)");
    if (m1 + n_ps <= 0) throw Error(RUVIII_ERR_INVALID, "ruviii_plan: no rows");
    // keep_resident: the matrices of the previous plan stay where they are (same shapes and the same log transform, no RUV1
    // pre-step applied to them yet): only the ctl- / k- / grouping-dependent state is rebuilt
    const bool keep = prm->keep_resident != 0;
    if (keep) {
        if (!h->uploaded) throw Error(RUVIII_ERR_STATE, "keep_resident: nothing has been uploaded under a previous plan");
        if (h->m1 != m1 || h->n_local_total != n || h->n_ps != n_ps)
            throw Error(RUVIII_ERR_STATE, "keep_resident: the shapes differ from those of the resident matrices");
        if (h->prm.apply_log != prm->apply_log || (prm->apply_log && h->prm.pseudo_count != prm->pseudo_count))
            throw Error(RUVIII_ERR_STATE, "keep_resident: apply.log / pseudo.count differ from those of the resident assay");
        if (h->eta_applied || h->eta_q > 0)
            throw Error(RUVIII_ERR_STATE, "keep_resident: a RUV1 pre-step (eta) rewrites the resident assay; upload again");
    }
    const bool kept_logged = keep && h->logged, kept_ps_pending = keep && h->ps_pending;
    h->prm = *prm;
    h->planned = h->uploaded = h->ran = h->logged = h->have_user_alpha = h->eta_applied = h->ps_pending = h->yc_valid = false;
    if (h->eta_q > 0 && h->eta_n != n)
        throw Error(RUVIII_ERR_INVALID, "eta has " + std::to_string(h->eta_n) + " columns, the assay has " + std::to_string(n) +
                                            " genes (ruviii_set_eta)");
    h->m1 = m1;
    h->n_ps = n_ps;
    h->m = m1 + n_ps;
    const bool fast = prm->mode == RUVIII_MODE_FAST;
    // rows that carry replicate structure: all m rows (ruvIII.R:101) or the PRPS rows only (fastruvIII.R:117-118)
    const int n_grp_rows = fast ? n_ps : h->m;
    const int row_offset = fast ? m1 : 0;          // stacked row index of group_of_row[0]
    const int m_for_M = fast ? n_ps : h->m;         // the `m` of `ncol(M) >= m` (fastruvIII.R:112,146)
    if (n_grp_rows > 0 && !group_of_row) throw Error(RUVIII_ERR_INVALID, "ruviii_plan: group_of_row is NULL");

    // ---- replicate structure: CSR of non-singleton groups (host, O(m log m)) ----
    std::map<int32_t, std::vector<int>> groups;
    for (int i = 0; i < n_grp_rows; ++i) {
        if (group_of_row[i] < 0) throw Error(RUVIII_ERR_INVALID, "ruviii_plan: negative group id");
        groups[group_of_row[i]].push_back(i + row_offset);
    }
    h->p = (int)groups.size();
    std::vector<int> grp_start{0}, grp_rows;
    h->max_group = 0;
    for (auto& kv : groups) {
        if (kv.second.size() < 2) continue;
        for (int r : kv.second) grp_rows.push_back(r);
        grp_start.push_back((int)grp_rows.size());
        h->max_group = std::max(h->max_group, (int)kv.second.size());
    }
    std::vector<int> all_start{0}, all_rows;               // every group, id order = factor level order (K9)
    if (prm->average_rep) {
        for (auto& kv : groups) {
            for (int r : kv.second) all_rows.push_back(fast ? r - row_offset : r);   // fast: row i of newY <-> PS row i
            all_start.push_back((int)all_rows.size());
        }
    }
    h->n_groups_active = (int)grp_start.size() - 1;
    h->active_rows = (int)grp_rows.size();
    h->active_all_ps = true;
    for (int r : grp_rows) h->active_all_ps = h->active_all_ps && r >= m1;
    h->R = h->active_rows - h->n_groups_active;     // order of the Gram: sum(size - 1) = m - ncol(M)
    h->no_replicates = h->p >= m_for_M;             // ruvIII.R:128-129

    // ---- k values ----
    h->ks.assign(ks, ks + nk);
    h->kmax = 0;
    for (int k : h->ks) {
        if (k < 0) throw Error(RUVIII_ERR_INVALID, "k must be >= 0");
        h->kmax = std::max(h->kmax, k);
    }

    // ---- gene shards ----
    const int nsh = (int)h->shards.size();
    int64_t local_ctl = 0;
    for (int g = 0; g < n; ++g) local_ctl += ctl[g] != 0;
    h->n_local_total = n;
    for (int i = 0; i < nsh; ++i) {
        Shard& s = h->shards[i];
        s.g0 = (int)((int64_t)n * i / nsh);
        s.n = (int)((int64_t)n * (i + 1) / nsh) - s.g0;
        s.ld = round_up(std::max(s.n, 1), kRowAlign);
    }
    // total number of control genes over all processes (r depends on it, ruvIII.R:140) and the gene count of every
    // global rank (the gene-side Gram needs to know where each peer's genes sit in the n_total x n_total matrix)
    h->n_ctl_global = local_ctl;
    h->gn.assign(h->world, 0);
    if (h->multi_process) {
        Shard& s = h->shards[0];
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        std::vector<double> v(h->world + 1, 0.0);
        v[0] = (double)local_ctl;
        v[1 + h->rank0] = (double)n;
        s.AG.alloc(v.size() + 16);
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(s.AG.p, v.data(), v.size() * 8, cudaMemcpyHostToDevice, s.stream));
        RUVIII_NCCL_CHECK(ncclAllReduce(s.AG.p, s.AG.p, v.size(), ncclDouble, ncclSum, s.comm, s.stream));
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(v.data(), s.AG.p, v.size() * 8, cudaMemcpyDeviceToHost, s.stream));
        RUVIII_CUDA_CHECK(cudaStreamSynchronize(s.stream));
        h->n_ctl_global = (int64_t)(v[0] + 0.5);
        for (int j = 0; j < h->world; ++j) h->gn[j] = (int)(v[1 + j] + 0.5);
    } else {
        for (int i = 0; i < nsh; ++i) h->gn[i] = h->shards[i].n;
    }
    h->goff.assign(h->world + 1, 0);
    for (int j = 0; j < h->world; ++j) h->goff[j + 1] = h->goff[j] + h->gn[j];
    h->n_total = h->goff[h->world];
    // r = min(m - ncol(M), sum(ctl))  (ruvIII.R:140); fastruvIII keeps k vectors of the n_ps x n residual
    const int rank_bound = m_for_M - h->p;
    h->rank_bound = rank_bound;
    h->r = (int)std::max<int64_t>(0, std::min<int64_t>(rank_bound, fast ? (int64_t)rank_bound : h->n_ctl_global));
    h->kk = std::min(h->kmax, h->r);
    if (h->kk > kMaxK) throw Error(RUVIII_ERR_UNSUPPORTED, "k > 32 is not supported by the fused update");
    // register / buffer extent: by max(k), not by min(max(k), r), so that a user-supplied fullalpha with more rows than r
    // (ruvIII.R:142 takes fullalpha[1:min(k, nrow(fullalpha)), ]) still fits
    h->kp = pad_k(std::max(std::min(h->kmax, kMaxK), 1));
    const bool skip = h->no_replicates || h->kk == 0;
    h->n_alpha_rows = skip ? 0 : (prm->n_alpha_rows < 0 ? h->r : std::min(h->r, std::max(h->kk, prm->n_alpha_rows)));
    h->n_keep = skip ? 0 : std::min(h->R, std::max(h->n_alpha_rows, h->kk + 1));   // one extra: the gap lambda_kk+1
    h->rows_w = fast ? m1 : h->m;
    h->mu_rows = fast ? m1 : h->m;                  // ruvIII.R:118 vs fastruvIII.R:136
    // Gram side: Z Z' (R x R) or Z'Z (n x n), whichever is smaller unless forced
    h->gene_side = !skip && (prm->gram_side == 2 || (prm->gram_side == 0 && h->R > h->n_total));
    if (h->gene_side) {
        if (h->n_alpha_rows > h->kk)
            throw Error(RUVIII_ERR_UNSUPPORTED, "gene-side Gram produces the leading max(k) rows of fullalpha only");
        if (h->kk + 1 > h->n_total) throw Error(RUVIII_ERR_INVALID, "k exceeds the number of genes");
    }
    h->gorder = h->gene_side ? h->n_total : h->R;
    if (!skip) h->n_keep = std::min(h->gorder, std::max(h->n_alpha_rows, h->kk + 1));
    // Gene side over several shards: Z'Z couples the gene shards, so the contraction is re-split over the ROWS of Z:
    // an all-to-all hands rank j the rows [zrow[j], zrow[j+1]) of every shard's Z, rank j forms the full
    // n_total x n_total partial Gram of those rows (balanced, symmetric half only) and the partials are allreduced.
    h->zrow.assign(h->world + 1, 0);
    for (int j = 0; j <= h->world; ++j) h->zrow[j] = (int)((int64_t)h->R * j / h->world);

    // ---- output slots ----
    h->k_eff.resize(nk);
    h->kmask = 0;
    h->dup_slots.clear();
    for (int i = 0; i <= kMaxK; ++i) h->slot_of_k[i] = -1;
    for (int i = 0; i < nk; ++i) {
        const int ke = skip ? 0 : std::min(h->ks[i], h->r);
        h->k_eff[i] = ke;
        if (h->slot_of_k[ke] < 0) {
            h->slot_of_k[ke] = i;
            h->kmask |= 1ull << ke;
        } else {
            h->dup_slots.emplace_back(i, h->slot_of_k[ke]);
        }
    }
    for (int i = 0; i <= kMaxK; ++i)
        if (h->slot_of_k[i] < 0) h->slot_of_k[i] = 0;

    // ---- device buffers ----
    const int kp = h->kp, R = h->R;
    // Peer-memory allreduce (one process per GPU, p2p.cu): the two matrices that are summed over the ranks -- the Gram and
    // [A | ac ac' | flag] -- are carved out of an exchange buffer every peer has mapped, so that the owner of a slice can
    // store its sum straight into every rank's copy.  Latency-bound sizes only: a 20,000^2 gene-side Gram is bandwidth-bound
    // and stays with NCCL's ring.
    bool p2p_g = false, p2p_ag = false;
    {
        const bool skip0 = h->no_replicates || h->kk == 0;
        if (!skip0 && h->multi_process) {
            const size_t lim = (size_t)8 << 20, cg = (size_t)h->gorder * h->gorder + 2, ca = (size_t)h->rows_w * kp + (size_t)kp * kp + 4;
            p2p_g = cg <= lim;
            p2p_ag = ca <= lim;
            if (p2p_g || p2p_ag) {
                const bool ok = p2p_setup(h, std::max(p2p_g ? cg : 0, p2p_ag ? ca : 0), (p2p_g ? cg + 64 : 0) + (p2p_ag ? ca + 64 : 0));
                p2p_g = p2p_g && ok;
                p2p_ag = p2p_ag && ok;
            }
        }
    }
    const bool want_ps = (prm->want_ps_rows || prm->average_rep) && !fast && n_ps > 0;
    size_t w_total = 0;
    for (int i = 0; i < nk; ++i) w_total += (size_t)h->rows_w * h->k_eff[i];
    for (int i = 0; i < nsh; ++i) {
        Shard& s = h->shards[i];
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        std::vector<int> cidx;
        for (int g = 0; g < s.n; ++g)
            if (ctl[s.g0 + g]) cidx.push_back(g);
        s.n_ctl = (int)cidx.size();
        s.ctl_idx_host = cidx;
        s.Y.alloc((size_t)std::max(m1, 1) * s.ld, !keep);
        s.PS.alloc((size_t)std::max(n_ps, 1) * s.ld, !keep);
        s.out.alloc((size_t)nk * std::max(m1, 1) * s.ld, false);
        if (want_ps) s.out_ps.alloc((size_t)nk * n_ps * s.ld, false);
        s.status.alloc(4);
        if (prm->average_rep) {
            s.all_start.alloc(all_start.size());
            s.all_rows.alloc(std::max<size_t>(all_rows.size(), 1));
            RUVIII_CUDA_CHECK(cudaMemcpy(s.all_start.p, all_start.data(), all_start.size() * sizeof(int), cudaMemcpyHostToDevice));
            if (!all_rows.empty())
                RUVIII_CUDA_CHECK(cudaMemcpy(s.all_rows.p, all_rows.data(), all_rows.size() * sizeof(int), cudaMemcpyHostToDevice));
            s.avg.alloc((size_t)nk * h->p * s.ld, false);
        }
        s.ldc = round_up(std::max(s.n_ctl, 1), kRowAlign);
        if (!skip || h->eta_q > 0) {
            s.ctl_idx.alloc(std::max<size_t>(cidx.size(), 1));
            if (!cidx.empty())
                RUVIII_CUDA_CHECK(cudaMemcpy(s.ctl_idx.p, cidx.data(), cidx.size() * sizeof(int), cudaMemcpyHostToDevice));
        }
        if (h->prps_n > 0 && h->prps_n == n_ps) {
            for (int r : h->prps_rows)
                if (r >= m1) throw Error(RUVIII_ERR_INVALID, "a PRPS set refers to sample " + std::to_string(r) + " but the assay has " +
                                                                 std::to_string(m1) + " samples");
            s.set_start.alloc(h->prps_start.size());
            s.set_rows.alloc(std::max<size_t>(h->prps_rows.size(), 1));
            RUVIII_CUDA_CHECK(cudaMemcpy(s.set_start.p, h->prps_start.data(), h->prps_start.size() * sizeof(int), cudaMemcpyHostToDevice));
            RUVIII_CUDA_CHECK(cudaMemcpy(s.set_rows.p, h->prps_rows.data(), h->prps_rows.size() * sizeof(int), cudaMemcpyHostToDevice));
        }
        if (h->eta_q > 0) {
            const int q = h->eta_q, qp = pad_k(q);
            s.eta.alloc((size_t)qp * s.ld);
            RUVIII_CUDA_CHECK(cudaMemcpy2D(s.eta.p, s.ld * 8, h->eta_host.data() + s.g0, (size_t)n * 8, (size_t)s.n * 8, q,
                                           cudaMemcpyHostToDevice));
            s.eta_t.alloc((size_t)qp * s.ld);
            s.etacT.alloc((size_t)std::max(s.n_ctl, 1) * qp);
            s.AG1.alloc((size_t)h->m * qp + (size_t)qp * qp);
            s.At1.alloc((size_t)h->m * qp);
            s.Linv1.alloc((size_t)qp * qp);
            s.cm1.alloc(qp);
            s.Yc1.alloc((size_t)h->m * s.ldc, false);
            if (n_ps > 0) s.PS1.alloc((size_t)n_ps * s.ld, false);
            s.status1.alloc(4);
        }
        if (!skip) {
            s.grp_start.alloc(grp_start.size());
            s.grp_rows.alloc(std::max<size_t>(grp_rows.size(), 1));
            RUVIII_CUDA_CHECK(cudaMemcpy(s.grp_start.p, grp_start.data(), grp_start.size() * sizeof(int), cudaMemcpyHostToDevice));
            if (!grp_rows.empty())
                RUVIII_CUDA_CHECK(cudaMemcpy(s.grp_rows.p, grp_rows.data(), grp_rows.size() * sizeof(int), cudaMemcpyHostToDevice));
            s.Y0.alloc((size_t)R * s.ld);
            const int GO = h->gorder;
            const int64_t ldR = round_up(R, kRowAlign);
            int64_t ldZt = ldR;                        // pitch of t(Z): the rows of Z this shard contracts, padded
            if (h->gene_side) {
                const int j = h->rank0 + i;
                const int rows_me = h->zrow[j + 1] - h->zrow[j];
                ldZt = round_up(std::max(rows_me, 1), kRowAlign);
                s.Zt.alloc((size_t)h->n_total * ldZt);                     // zeroed: the padding columns stay zero
                if (h->world > 1) {
                    size_t stage = 0;
                    for (int q = 0; q < h->world; ++q) stage += (size_t)round_up(std::max(h->gn[q], 1), kRowAlign);
                    s.Zstage.alloc((size_t)std::max(rows_me, 1) * stage, false);
                }
            }
            s.gp = h->gene_side ? plan_gram(GO, (int)ldZt, h->n_sms, h->gram_impl) : plan_gram(R, (int)s.ld, h->n_sms, h->gram_impl);
            s.gram_ws.alloc(s.gp.workspace_bytes / sizeof(double), false);
            std::vector<int2> tm(s.gp.n_tiles);
            fill_gram_tile_map(s.gp, tm.data());
            s.tile_map.alloc(tm.size());
            RUVIII_CUDA_CHECK(cudaMemcpy(s.tile_map.p, tm.data(), tm.size() * sizeof(int2), cudaMemcpyHostToDevice));
            {
                double* q = p2p_g ? p2p_carve(h, (size_t)GO * GO + 2) : nullptr;
                if (q) s.G.alias(q, (size_t)GO * GO + 2); else s.G.alloc((size_t)GO * GO + 2);
            }
            s.evals_asc.alloc(GO);
            s.evals.alloc(std::max(h->n_keep, kp) + 2);
            s.U.alloc((size_t)GO * kp);
            if (h->n_alpha_rows > h->kk) s.Vd.alloc((size_t)GO * h->n_keep);
            s.rsplit = project_rsplit(R, s.n, h->n_sms);
            s.part.alloc((size_t)s.rsplit * kp * s.ld);
            s.alpha.alloc((size_t)kp * s.ld);
            s.at.alloc((size_t)kp * s.ld);
            s.acT.alloc((size_t)std::max(s.n_ctl, 1) * kp);
            {                                                            // [A | ac ac' | eigensolver flag]
                double* q = p2p_ag ? p2p_carve(h, (size_t)h->rows_w * kp + (size_t)kp * kp + 4) : nullptr;
                if (q) s.AG.alias(q, (size_t)h->rows_w * kp + (size_t)kp * kp + 4); else s.AG.alloc((size_t)h->rows_w * kp + (size_t)kp * kp + 4);
            }
            s.Yc.alloc((size_t)h->rows_w * s.ldc, false);
            s.At.alloc((size_t)h->rows_w * kp);
            s.Linv.alloc((size_t)kp * kp);
            s.colmean.alloc(kp);
            s.solve_ws.alloc(solve_fused_ws_doubles(h->n_sms));
            if (h->n_alpha_rows > h->kk) s.falpha.alloc((size_t)h->n_alpha_rows * s.ld);
            s.W.alloc(std::max<size_t>(w_total, 1));
            if (s.gp.impl == 2) {
                launch_gram_tma_prepare(s.gp, h->gene_side ? s.Zt.p : s.Y0.p, h->gene_side ? ldZt : s.ld, s.tmap);
            }
            if (h->upd_impl == 2 && m1 > 0) update_tma_prepare(s.Y.p, s.ld, m1, s.tmap_y);
        }
    }
    // ---- eigensolver: subspace iteration on every shard that runs it, cuSOLVER on global rank 0 only ----
    h->use_topk = false;
    h->persist_rb = 0;
    if (!skip && R > 0) {
        const int GO = h->gorder;
        // block subspace iteration when only the leading kk rows of fullalpha are wanted and the block (32) leaves
        // room for oversampling; cuSOLVER otherwise and as the fallback
        const bool topk_ok = h->n_alpha_rows == h->kk && h->kk <= kTopkMaxK && GO >= 4 * kTopkBlock &&
                             h->rank_bound >= 2 * kTopkBlock;
        if (h->eig_impl == 3 && !topk_ok)
            throw Error(RUVIII_ERR_UNSUPPORTED, "RUVIII_EIG_IMPL=3 but the problem is outside the subspace-iteration range");
        h->use_topk = topk_ok && (h->eig_impl == 0 || h->eig_impl == 3);
        if (h->use_topk) {
            // The persistent single-kernel form runs redundantly on EVERY shard (the allreduced Gram is the same bits
            // everywhere), which removes the broadcast of U; the multi-kernel form (large orders) runs on global rank 0.
            h->persist_rb = topk_persist_rows(GO, h->n_sms);
            for (int i = 0; i < nsh; ++i) {
                if (!h->persist_rb && !(i == 0 && h->rank0 == 0)) continue;
                Shard& s = h->shards[i];
                RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
                const size_t base = topk_work_doubles(GO, kTopkBlock);
                s.topk_buf.alloc(base + topk_persist_doubles(h->n_sms), false);
                s.topk_state.alloc(16);
                topk_bind(s.topk, s.topk_buf.p, s.topk_state.p, GO, kTopkBlock);
                s.persist_ws = s.topk_buf.p + base;
            }
        }
        if (h->rank0 == 0) {
            Shard& s = h->shards[0];
            RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
            if (!h->solver) {
                RUVIII_SOLVER_CHECK(cusolverDnCreate(&h->solver));
                RUVIII_SOLVER_CHECK(cusolverDnSetStream(h->solver, s.stream));
            }
            int lwork = 0, meig = 0;
            if (h->eig_impl == 1 || h->n_keep >= GO) {
                RUVIII_SOLVER_CHECK(cusolverDnDsyevd_bufferSize(h->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, GO,
                                                                s.G.p, GO, s.evals_asc.p, &lwork));
            } else {
                RUVIII_SOLVER_CHECK(cusolverDnDsyevdx_bufferSize(h->solver, CUSOLVER_EIG_MODE_VECTOR, CUSOLVER_EIG_RANGE_I,
                                                                 CUBLAS_FILL_MODE_UPPER, GO, s.G.p, GO, 0.0, 0.0,
                                                                 GO - h->n_keep + 1, GO, &meig, s.evals_asc.p, &lwork));
            }
            h->solver_work.alloc((size_t)lwork, false);
            h->solver_info.alloc(4);
        }
    }
    h->planned = true;
    if (keep) {
        h->uploaded = true;
        h->logged = kept_logged;
        // PRPS rows: those uploaded from the host stay; those built from registered sets are rebuilt on the next run (K0 is
        // idempotent and costs 0.15 ms)
        h->ps_pending = kept_ps_pending || h->ps_from_sets;
    }
}

// ruvIII.R:88-89 on the resident assay, once per upload (ruviii_run does the same on its first pass)
void ensure_logged(ruviii_handle* h) {
    if (!h->prm.apply_log || h->logged) return;
    for (auto& s : h->shards) {
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        launch_log2_inplace(s.Y.p, s.ld, h->m1, s.n, h->prm.pseudo_count, s.stream);
    }
    h->logged = true;
}

void do_upload(ruviii_handle* h, const double* Y, int64_t ldY, const double* PS, int64_t ldPS) {
    if (!h->planned) throw Error(RUVIII_ERR_STATE, "ruviii_upload before ruviii_plan");
    // PS == NULL with registered PRPS sets: replicate.data is built on the device from the assay (K0, first run)
    const bool build_ps = h->n_ps > 0 && !PS && h->prps_n == h->n_ps;
    if ((h->m1 > 0 && (!Y || ldY < h->n_local_total)) || (h->n_ps > 0 && !build_ps && (!PS || ldPS < h->n_local_total)))
        throw Error(RUVIII_ERR_INVALID, "ruviii_upload: NULL matrix or row stride < n");
    h->ps_pending = build_ps;
    h->ps_from_sets = build_ps;
    for (auto& s : h->shards) {
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_START], s.stream));
        if (h->n_ps > 0 && !build_ps)
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(s.PS.p, s.ld * 8, PS + s.g0, ldPS * 8, (size_t)s.n * 8, h->n_ps,
                                                cudaMemcpyHostToDevice, s.stream));
        if (h->m1 > 0)
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(s.Y.p, s.ld * 8, Y + s.g0, ldY * 8, (size_t)s.n * 8, h->m1,
                                                cudaMemcpyHostToDevice, s.stream));
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_LOG], s.stream));
    }
    sync_all(h);
    float ms = 0.f;
    RUVIII_CUDA_CHECK(cudaSetDevice(h->shards[0].device));
    cudaEventElapsedTime(&ms, h->shards[0].ev[E_START], h->shards[0].ev[E_LOG]);
    h->info.ms_h2d = ms;
    h->uploaded = true;
    h->logged = false;
    h->eta_applied = false;
    h->yc_valid = false;
    h->ran = false;
}

// Products with G the subspace iteration may spend before cuSOLVER takes over.  In the multi-kernel form a pass costs ~60 us
// against ~15 ms of syevdx at R ~ 10^3; in the persistent form a filtered product costs a few microseconds.
static int eig_max_products(const ruviii_handle* h) {
    if (h->persist_rb) return env_int("RUVIII_EIG_MAXIT", 400);
    return env_int("RUVIII_EIG_MAXIT", h->kk > kTopkBlock / 2 ? kTopkPatience : 150);
}

// Global rank 0 decomposes the Gram on shards[0]; the caller broadcasts.  This entry WAITS for the verdict of the subspace
// iteration (the pipelined host path uses it, and do_run when its optimistic persistent solve did not converge:
// allow_subspace = false goes straight to cuSOLVER).
void eigensolve(ruviii_handle* h, bool allow_subspace) {
    Shard& s = h->shards[0];
    const int R = h->gorder;            // order of the Gram being decomposed
    RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
    h->h_flags[0] = 0;
    h->eig_iterations = 0;
    if (h->use_topk && allow_subspace) {
        TopkResult tr;
        if (h->persist_rb)
            tr = topk_eigen_persistent(s.G.p, R, h->kk, h->kp, h->n_keep, s.topk, s.persist_ws, s.U.p, s.evals.p, nullptr, h->eig_tol,
                                       eig_max_products(h), h->n_sms, &h->h_flags[64], true, s.stream);
        else
            tr = topk_eigen(s.G.p, R, h->kk, h->kp, h->n_keep, kTopkBlock, s.topk, s.U.p, s.evals.p, h->eig_tol, eig_max_products(h),
                            h->n_sms, &h->h_flags[60], s.stream);
        h->launches += tr.launches;
        h->eig_iterations = tr.iterations;
        if (tr.converged) {
            h->eig_solver_used = 3;
            return;
        }
        if (h->eig_impl == 3)
            throw Error(RUVIII_ERR_SOLVER, "subspace iteration did not converge in " + std::to_string(tr.iterations) +
                                               " passes (RUVIII_EIG_IMPL=3 forbids the cuSOLVER fallback)");
    }
    int n_found = R;
    h->eig_solver_used = (h->eig_impl == 1 || h->n_keep >= R) ? 1 : 2;
    if (h->eig_impl == 1 || h->n_keep >= R) {
        RUVIII_SOLVER_CHECK(cusolverDnDsyevd(h->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, R, s.G.p, R,
                                             s.evals_asc.p, h->solver_work.p, (int)h->solver_work.count,
                                             h->solver_info.p));
    } else {
        int meig = 0;
        RUVIII_SOLVER_CHECK(cusolverDnDsyevdx(h->solver, CUSOLVER_EIG_MODE_VECTOR, CUSOLVER_EIG_RANGE_I,
                                              CUBLAS_FILL_MODE_UPPER, R, s.G.p, R, 0.0, 0.0, R - h->n_keep + 1, R, &meig,
                                              s.evals_asc.p, h->solver_work.p, (int)h->solver_work.count,
                                              h->solver_info.p));
        n_found = meig;
        if (meig != h->n_keep) throw Error(RUVIII_ERR_SOLVER, "syevdx returned an unexpected number of eigenpairs");
    }
    RUVIII_CUDA_CHECK(cudaMemcpyAsync(&h->h_flags[0], h->solver_info.p, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    launch_pick_top(s.G.p, R, s.evals_asc.p, n_found, R, h->n_keep, h->kk, h->kp, h->n_alpha_rows > h->kk ? s.Vd.p : nullptr,
                    s.U.p, s.evals.p, s.stream);
    h->launches += 1;
}

void do_run(ruviii_handle* h, ruviii_info* info_out) {
    if (!h->planned || !h->uploaded) throw Error(RUVIII_ERR_STATE, "ruviii_run needs ruviii_plan and ruviii_upload first");
    const bool fast = h->prm.mode == RUVIII_MODE_FAST;
    const bool skip = h->no_replicates || h->kk == 0;
    const int kp = h->kp, kk = h->kk, R = h->R, nk = (int)h->ks.size();
    const bool want_ps = (h->prm.want_ps_rows || h->prm.average_rep) && !fast && h->n_ps > 0;
    h->launches = 0;
    const double host_t_enter = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    auto rec = [&](int e) {
        for (auto& s : h->shards) {
            RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
            RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[e], s.stream));
        }
    };
    auto for_shards = [&](auto&& fn) {
        for (auto& s : h->shards) {
            RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
            fn(s);
        }
    };
    NvtxRange nvtx_run("ruviii_run");
    rec(E_START);
    if (h->prm.apply_log && !h->logged) {             // ruvIII.R:88-89, once per upload
        ensure_logged(h);
        h->launches += 1;
    }
    if (h->ps_pending) {
        // K0  prpsForCategoricalUV.R:123-139 / prpsForContinuousUV.R:190-201: pseudo-sample j = rowMeans(expre.data[, set_j]),
        // taken from the (already log-transformed) resident assay; replicate.data never crosses PCIe
        for_shards([&](Shard& s) {
            RowSrc src{s.Y.p, s.ld, nullptr, 0, h->m1};
            launch_group_mean(src, s.set_start.p, s.set_rows.p, h->n_ps, s.n, s.PS.p, s.ld, s.stream);
        });
        h->ps_pending = false;
        h->launches += 1;
    }
    if (h->eta_q > 0 && !h->eta_applied) {
        apply_ruv1(h);
        h->eta_applied = true;
    }
    rec(E_LOG);
    if (skip) {
        // ncol(M) >= m (ruvIII.R:128-129) or k == 0 (:134-136): newY = Y
        for_shards([&](Shard& s) {
            for (int i = 0; i < nk; ++i) {
                RUVIII_CUDA_CHECK(cudaMemcpyAsync(s.out.p + (size_t)i * h->m1 * s.ld, s.Y.p, (size_t)h->m1 * s.ld * 8,
                                                  cudaMemcpyDeviceToDevice, s.stream));
                if (want_ps)
                    RUVIII_CUDA_CHECK(cudaMemcpyAsync(s.out_ps.p + (size_t)i * h->n_ps * s.ld, s.PS.p,
                                                      (size_t)h->n_ps * s.ld * 8, cudaMemcpyDeviceToDevice, s.stream));
            }
        });
        for (int e = E_RESIDOP; e < E_COUNT; ++e) rec(e);
    } else {
        // Control-gene gather on the auxiliary stream.  It depends on Y only; it is forked AFTER the (bandwidth- and
        // tensor-bound) residop and Gram so that it fills the otherwise idle GPU during the eigensolve, whose kernels
        // are latency-bound and occupy a handful of SMs (forked at the start it only stole HBM bandwidth from K1).
        // The control-gene block Yc = Y[, ctl] depends on the uploaded matrices only (after their log transform, PRPS
        // construction and RUV1 pre-step, all of which run once per upload), so it is gathered once per upload and every later
        // pass on the resident assay reuses it: 5 % extra storage instead of 1.1 GB of sector traffic per pass.
        bool defer_gather = false;
        auto fork_gather = [&](bool record = true) {
            if (h->yc_valid) {
                for_shards([&](Shard& s) {
                    if (record) RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_fork, s.stream));
                    RUVIII_CUDA_CHECK(cudaStreamWaitEvent(s.copy_stream, s.ev_fork, 0));
                    RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_g0, s.copy_stream));
                    RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_g1, s.copy_stream));
                    RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_gather, s.copy_stream));
                });
                return;
            }
            h->yc_valid = env_int("RUVIII_YC_CACHE", 1) != 0;
            for_shards([&](Shard& s) {
                RowSrc src{s.Y.p, s.ld, s.PS.p, s.ld, h->m1};
                if (record) RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_fork, s.stream));
                RUVIII_CUDA_CHECK(cudaStreamWaitEvent(s.copy_stream, s.ev_fork, 0));
                RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_g0, s.copy_stream));
                launch_ctl_gather(src, h->rows_w, s.ctl_idx.p, s.n_ctl, s.Yc.p, s.ldc, 0, 0.0, s.copy_stream);
                RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_g1, s.copy_stream));
                RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_gather, s.copy_stream));
            });
            h->launches += 1;
        };
        // RUVIII_GATHER_AT: 0 = fork after the Gram (default), 1 = after the residual (overlaps the Gram instead)
        static const int gather_at = env_int("RUVIII_GATHER_AT", 0);
        if (h->have_user_alpha) fork_gather();
        if (!h->have_user_alpha) {
            NvtxRange nvtx_k12("K1 residop + K2 Gram + allreduce");
            for_shards([&](Shard& s) {
                RowSrc src{s.Y.p, s.ld, s.PS.p, s.ld, h->m1};
                launch_helmert(src, s.grp_start.p, s.grp_rows.p, h->n_groups_active, h->max_group, s.n, s.Y0.p, s.ld, s.stream);
            });
            h->launches += 1;
            rec(E_RESIDOP);
            if (gather_at == 1) fork_gather();
            const int GO = h->gorder;
            if (h->gene_side) {
                // t(Z) restricted to the rows this shard contracts.  One shard: all of Z.  Several: all-to-all of row
                // blocks first (every shard sends rows [zrow[q], zrow[q+1]) of its gene slab to rank q).
                if (h->world > 1) {
                    RUVIII_NCCL_CHECK(ncclGroupStart());
                    for (size_t i = 0; i < h->shards.size(); ++i) {
                        Shard& s = h->shards[i];
                        const int j = h->rank0 + (int)i;
                        const int rows_me = h->zrow[j + 1] - h->zrow[j];
                        size_t off = 0;
                        for (int q = 0; q < h->world; ++q) {
                            const int rows_q = h->zrow[q + 1] - h->zrow[q];
                            const int64_t ld_q = round_up(std::max(h->gn[q], 1), kRowAlign);
                            if (rows_q > 0)
                                RUVIII_NCCL_CHECK(ncclSend(s.Y0.p + (size_t)h->zrow[q] * s.ld, (size_t)rows_q * s.ld, ncclDouble, q,
                                                           s.comm, s.stream));
                            if (rows_me > 0)
                                RUVIII_NCCL_CHECK(ncclRecv(s.Zstage.p + off, (size_t)rows_me * ld_q, ncclDouble, q, s.comm, s.stream));
                            off += (size_t)rows_me * ld_q;
                        }
                    }
                    RUVIII_NCCL_CHECK(ncclGroupEnd());
                }
                for (size_t i = 0; i < h->shards.size(); ++i) {
                    Shard& s = h->shards[i];
                    RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
                    const int j = h->rank0 + (int)i;
                    const int rows_me = h->zrow[j + 1] - h->zrow[j];
                    const int64_t ldZt = round_up(std::max(rows_me, 1), kRowAlign);
                    if (h->world == 1) {
                        launch_transpose(s.Y0.p, s.ld, R, s.n, s.Zt.p, ldZt, s.stream);
                    } else {
                        size_t off = 0;
                        for (int q = 0; q < h->world; ++q) {
                            const int64_t ld_q = round_up(std::max(h->gn[q], 1), kRowAlign);
                            launch_transpose(s.Zstage.p + off, ld_q, rows_me, h->gn[q], s.Zt.p + (size_t)h->goff[q] * ldZt, ldZt,
                                             s.stream);
                            off += (size_t)rows_me * ld_q;
                        }
                    }
                    launch_gram(s.gp, s.Zt.p, ldZt, s.tile_map.p, s.gram_ws.p, s.G.p, s.tmap, s.stream);
                }
                h->launches += 2 + h->world;
            } else {
                for_shards([&](Shard& s) { launch_gram(s.gp, s.Y0.p, s.ld, s.tile_map.p, s.gram_ws.p, s.G.p, s.tmap, s.stream); });
                h->launches += 2;
            }
            rec(E_GRAM);
            // Persistent eigensolver: its ~100 CTAs must all become resident, and a gather that is already running (6,000 short
            // CTAs on a low-priority stream) keeps back-filling every slot that frees up (measured: the solve then starts
            // ~0.17 ms late, when the gather is done).  So the fork is only RECORDED here; the gather itself is launched
            // right after the solver, and fills what the solver leaves free.
            defer_gather = gather_at == 0 && !h->have_user_alpha && h->use_topk && h->persist_rb > 0;
            if (gather_at == 0) {
                if (defer_gather) for_shards([&](Shard& s) { RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_fork, s.stream)); });
                else fork_gather();
            }
            {
                std::vector<double*> b;
                for (auto& s : h->shards) b.push_back(s.G.p);
                allreduce_sum(h, b, (size_t)GO * GO);
            }
            rec(E_AR1);
        }
        // Everything behind the Gram.  classic = false: the persistent eigensolver is enqueued on every shard WITHOUT waiting
        // for its verdict and the rest of the pass is queued behind it; its "not converged" flag travels in the [A | ac ac']
        // allreduce, so every rank learns after the final synchronisation whether ANY rank has to fall back, and then all of
        // them run this tail once more with classic = true (cuSOLVER on global rank 0 + broadcast).
        bool persist_used = false, eig_status_pending = false;
        std::string eig_fail_local;
        int eig_fail_code = RUVIII_ERR_SOLVER;
        auto check_eig_status = [&]() {          // after a synchronisation: did global rank 0's decomposition fail?
            if (!eig_status_pending) return;
            eig_status_pending = false;
            if (!eig_fail_local.empty()) throw Error(eig_fail_code, eig_fail_local);
            for (size_t i = 0; i < h->shards.size(); ++i) {
                if (h->h_dflag[16 + 2 * i] != 0.0)
                    throw Error(RUVIII_ERR_SOLVER, "the eigensolve failed on global rank 0 (its own error message has the details)");
                int devinfo = 0;
                std::memcpy(&devinfo, &h->h_dflag[16 + 2 * i + 1], sizeof(int));
                if (devinfo != 0) throw Error(RUVIII_ERR_SOLVER, "cuSOLVER syevd devInfo = " + std::to_string(devinfo) + " on global rank 0");
            }
        };
        auto tail = [&](bool classic) {
        const int GO = h->gorder;
        const size_t flag_off = (size_t)h->rows_w * kp + (size_t)kp * kp;
        persist_used = !classic && !h->have_user_alpha && h->use_topk && h->persist_rb > 0;
        NvtxRange nvtx_tail("K3 eigensolve + K4 projection + K5 control products + solve + K6 update");
        if (!h->have_user_alpha) {
            if (persist_used) {
                for (size_t i = 0; i < h->shards.size(); ++i) {
                    Shard& s = h->shards[i];
                    RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
                    topk_eigen_persistent(s.G.p, GO, kk, kp, h->n_keep, s.topk, s.persist_ws, s.U.p, s.evals.p, s.AG.p + flag_off,
                                          h->eig_tol, eig_max_products(h), h->n_sms, &h->h_flags[64 + 16 * i], false, s.stream);
                }
                h->launches += 1;
                h->eig_solver_used = 3;
                if (defer_gather) { fork_gather(false); defer_gather = false; }
                rec(E_EIG);
                rec(E_BCAST);
            } else {
                for_shards([&](Shard& s) { RUVIII_CUDA_CHECK(cudaMemsetAsync(s.AG.p + flag_off, 0, sizeof(double), s.stream)); });
                // Global rank 0 decomposes; everybody else waits in the broadcast.  A failure on rank 0 (cuSOLVER status, syevdx
                // eigenpair count, no convergence under RUVIII_EIG_IMPL=3, devInfo != 0) must not leave the peers blocked or
                // let them return garbage: the two spare doubles behind the eigenvalues carry [host-side failure, devInfo bits]
                // through the broadcast, and every rank raises after the pass has drained.
                const size_t ev_tail = (size_t)std::max(h->n_keep, kp);
                eig_fail_local.clear();
                if (h->rank0 == 0) {
                    Shard& s0 = h->shards[0];
                    RUVIII_CUDA_CHECK(cudaSetDevice(s0.device));
                    RUVIII_CUDA_CHECK(cudaMemsetAsync(s0.evals.p + ev_tail, 0, 2 * sizeof(double), s0.stream));
                    try {
                        eigensolve(h, !classic);
                    } catch (const Error& e) {
                        eig_fail_local = e.what();
                        eig_fail_code = e.code;
                        cudaGetLastError();
                    }
                    h->h_dflag[48] = eig_fail_local.empty() ? 0.0 : 1.0;
                    RUVIII_CUDA_CHECK(cudaMemcpyAsync(s0.evals.p + ev_tail, &h->h_dflag[48], sizeof(double), cudaMemcpyHostToDevice, s0.stream));
                    if (h->eig_solver_used != 3 && h->solver_info.p)
                        RUVIII_CUDA_CHECK(cudaMemcpyAsync(s0.evals.p + ev_tail + 1, h->solver_info.p, sizeof(int), cudaMemcpyDeviceToDevice,
                                                          s0.stream));
                }
                rec(E_EIG);
                {
                    std::vector<double*> b;
                    for (auto& s : h->shards) b.push_back(s.U.p);
                    broadcast0(h, b, (size_t)GO * kp);
                    b.clear();
                    for (auto& s : h->shards) b.push_back(s.evals.p);
                    broadcast0(h, b, ev_tail + 2);
                    for (size_t i = 0; i < h->shards.size(); ++i) {
                        Shard& s = h->shards[i];
                        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
                        RUVIII_CUDA_CHECK(cudaMemcpyAsync(&h->h_dflag[16 + 2 * i], s.evals.p + ev_tail, 2 * sizeof(double),
                                                          cudaMemcpyDeviceToHost, s.stream));
                    }
                    eig_status_pending = true;
                    if (h->n_alpha_rows > kk) {
                        b.clear();
                        for (auto& s : h->shards) b.push_back(s.Vd.p);
                        broadcast0(h, b, (size_t)GO * h->n_keep);
                    }
                }
                rec(E_BCAST);
            }
            for_shards([&](Shard& s) {
                if (h->gene_side) {
                    const int j = h->rank0 + (int)(&s - h->shards.data());      // this shard's rows of V
                    launch_alpha_from_v(s.U.p + (size_t)h->goff[j] * kp, s.evals.p, s.n, kk, kp, s.ld, s.alpha.p, s.stream);
                    launch_alpha_finalize(s.alpha.p, 1, kp, s.n, s.ld, s.alpha.p, s.ctl_idx.p, s.n_ctl, s.acT.p, s.stream);
                } else {
                    launch_project(s.Y0.p, s.ld, R, s.n, s.U.p, kp, s.rsplit, s.part.p, s.ld, s.stream, kk);
                    launch_alpha_finalize(s.part.p, s.rsplit, kp, s.n, s.ld, s.alpha.p, s.ctl_idx.p, s.n_ctl, s.acT.p, s.stream);
                }
            });
            h->launches += 3;
            if (h->n_alpha_rows > kk) {
                // rows kk.. of fullalpha (return.info = TRUE, ruvIII.R:187-188): a plain library GEMM off the hot path
                if (!h->blas) {
                    RUVIII_CUDA_CHECK(cudaSetDevice(h->shards[0].device));
                    if (cublasCreate(&h->blas) != CUBLAS_STATUS_SUCCESS) throw Error(RUVIII_ERR_CUDA, "cublasCreate failed");
                }
                if (h->shards.size() != 1)
                    throw Error(RUVIII_ERR_UNSUPPORTED, "n_alpha_rows = -1 needs one GPU per process");
                Shard& s = h->shards[0];
                const double one = 1.0, zero = 0.0;
                cublasSetStream(h->blas, s.stream);
                if (cublasDgemm(h->blas, CUBLAS_OP_N, CUBLAS_OP_N, (int)s.ld, h->n_alpha_rows, R, &one, s.Y0.p, (int)s.ld, s.Vd.p,
                                R, &zero, s.falpha.p, (int)s.ld) != CUBLAS_STATUS_SUCCESS)
                    throw Error(RUVIII_ERR_CUDA, "cublasDgemm (fullalpha) failed");
            }
        } else {
            for (int e = E_RESIDOP; e <= E_BCAST; ++e) rec(e);
            for_shards([&](Shard& s) {
                RUVIII_CUDA_CHECK(cudaMemsetAsync(s.AG.p + flag_off, 0, sizeof(double), s.stream));
                launch_alpha_finalize(s.alpha.p, 1, kp, s.n, s.ld, s.alpha.p, s.ctl_idx.p, s.n_ctl, s.acT.p, s.stream);
            });
            h->launches += 2;
        }
        rec(E_PROJECT);
        for_shards([&](Shard& s) {
            RowSrc src{s.Y.p, s.ld, s.PS.p, s.ld, h->m1};
            // ac %*% t(ac) (one latency-bound CTA) runs on the auxiliary stream beside the products
            RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_fork, s.stream));
            RUVIII_CUDA_CHECK(cudaStreamWaitEvent(s.copy_stream, s.ev_fork, 0));
            launch_ac_gram(s.acT.p, s.n_ctl, kp, s.AG.p + (size_t)h->rows_w * kp, s.copy_stream);
            RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_acgram, s.copy_stream));
            RUVIII_CUDA_CHECK(cudaStreamWaitEvent(s.stream, s.ev_gather, 0));
            launch_ctl_products(src, h->rows_w, s.ctl_idx.p, s.n_ctl, s.acT.p, kp, s.AG.p, nullptr, s.n_ctl > 0 ? s.Yc.p : nullptr,
                                s.ldc, s.stream);
            RUVIII_CUDA_CHECK(cudaStreamWaitEvent(s.stream, s.ev_acgram, 0));
        });
        h->launches += 2;
        rec(E_CTL);
        {
            std::vector<double*> b;
            for (auto& s : h->shards) b.push_back(s.AG.p);
            allreduce_sum(h, b, flag_off + 1);
        }
        rec(E_AR2);
        for (size_t i = 0; i < h->shards.size(); ++i) {
            Shard& s = h->shards[i];
            RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
            RUVIII_CUDA_CHECK(cudaMemcpyAsync(&h->h_dflag[i], s.AG.p + flag_off, sizeof(double), cudaMemcpyDeviceToHost, s.stream));
        }
        for_shards([&](Shard& s) { launch_solve(h, s); });
        rec(E_SOLVE);
        for_shards([&](Shard& s) {
            const int64_t stride = (int64_t)h->m1 * s.ld;
            if (h->upd_impl == 3)
                launch_update_smem(s.Y.p, s.ld, h->m1, s.n, s.At.p, kp, s.at.p, s.ld, kk, h->kmask, h->slot_of_k, s.out.p, s.ld,
                                   stride, s.stream);
            else if (h->upd_impl == 2)
                launch_update_tma(s.tmap_y, h->m1, s.n, s.At.p, kp, s.at.p, s.ld, kk, h->kmask, h->slot_of_k, s.out.p, s.ld,
                                  stride, s.stream);
            else
                launch_update(s.Y.p, s.ld, h->m1, s.n, s.At.p, kp, s.at.p, s.ld, kk, h->kmask, h->slot_of_k, s.out.p, s.ld,
                              stride, s.stream);
            if (want_ps)
                launch_update(s.PS.p, s.ld, h->n_ps, s.n, s.At.p + (size_t)h->m1 * kp, kp, s.at.p, s.ld, kk, h->kmask,
                              h->slot_of_k, s.out_ps.p, s.ld, (int64_t)h->n_ps * s.ld, s.stream);
            for (auto& d : h->dup_slots) {
                RUVIII_CUDA_CHECK(cudaMemcpyAsync(s.out.p + (size_t)d.first * stride, s.out.p + (size_t)d.second * stride,
                                                  (size_t)stride * 8, cudaMemcpyDeviceToDevice, s.stream));
                if (want_ps)
                    RUVIII_CUDA_CHECK(cudaMemcpyAsync(s.out_ps.p + (size_t)d.first * h->n_ps * s.ld,
                                                      s.out_ps.p + (size_t)d.second * h->n_ps * s.ld,
                                                      (size_t)h->n_ps * s.ld * 8, cudaMemcpyDeviceToDevice, s.stream));
            }
        });
        h->launches += 1 + (want_ps ? 1 : 0);
        rec(E_UPDATE);
        for (size_t i = 0; i < h->shards.size(); ++i) {
            Shard& s = h->shards[i];
            RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
            RUVIII_CUDA_CHECK(cudaMemcpyAsync(&h->h_flags[1 + i], s.status.p, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
        }
        };   // tail
        tail(false);
        if (eig_status_pending) {
            sync_all(h);
            check_eig_status();
        }
        if (persist_used) {
            sync_all(h);
            bool redo = false;
            for (size_t i = 0; i < h->shards.size(); ++i) redo = redo || h->h_dflag[i] != 0.0;
            h->eig_iterations = h->h_flags[64 + 1];
            if (redo) {
                if (h->eig_impl == 3)
                    throw Error(RUVIII_ERR_SOLVER, "subspace iteration did not converge in " + std::to_string(h->eig_iterations) +
                                                       " products (RUVIII_EIG_IMPL=3 forbids the cuSOLVER fallback)");
                tail(true);
                sync_all(h);
                check_eig_status();
            }
        }
    }
    if (h->prm.average_rep) {                               // ruvIII.R:148-149, after either branch
        for_shards([&](Shard& s) {
            for (int i = 0; i < nk; ++i) {
                RowSrc src{s.out.p + (size_t)i * h->m1 * s.ld, s.ld,
                           want_ps ? s.out_ps.p + (size_t)i * h->n_ps * s.ld : nullptr, s.ld, h->m1};
                launch_group_mean(src, s.all_start.p, s.all_rows.p, h->p, s.n, s.avg.p + (size_t)i * h->p * s.ld, s.ld, s.stream);
            }
        });
        h->launches += nk;
    }
    sync_all(h);
    const double host_t_done = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    if (h->multi_process) p2p_check(h);
    if (!skip) {
        if (h->rank0 == 0 && !h->have_user_alpha && h->h_flags[0] != 0)
            throw Error(RUVIII_ERR_SOLVER, "cuSOLVER syevd devInfo = " + std::to_string(h->h_flags[0]));
        for (size_t i = 0; i < h->shards.size(); ++i)
            if (h->h_flags[1 + i] != 0)
                throw Error(RUVIII_ERR_SINGULAR,
                            "ac %*% t(ac) is not positive definite at pivot " + std::to_string(h->h_flags[1 + i]) +
                                " (solve() in ruvIII.R:144 would fail)");
    }
    // ---- info ----
    ruviii_info& I = h->info;
    const float keep_h2d = I.ms_h2d, keep_d2h = I.ms_d2h;
    std::memset(&I, 0, sizeof(I));
    I.struct_size = sizeof(ruviii_info);
    I.ms_h2d = keep_h2d;
    I.ms_d2h = keep_d2h;
    I.m = h->m; I.m1 = h->m1; I.n_ps = h->n_ps; I.n = h->n_local_total; I.p = h->p; I.active_rows = h->active_rows; I.r = h->r; I.gram_order = h->gorder;
    I.n_ctl = (int)h->n_ctl_global; I.kmax_eff = h->kk; I.no_replicates = h->no_replicates; I.gram_side = h->gene_side ? 2 : 1;
    I.n_shards = (int)h->shards.size(); I.kernel_launches = h->launches;
    I.eig_solver = (skip || h->have_user_alpha) ? 0 : h->eig_solver_used; I.eig_iterations = h->eig_iterations;
    I.host_t_enter = host_t_enter; I.host_t_done = host_t_done;
    {
        Shard& s = h->shards[0];
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        auto el = [&](int a, int b) { float ms = 0.f; cudaEventElapsedTime(&ms, s.ev[a], s.ev[b]); return ms; };
        I.ms_log = el(E_START, E_LOG);
        I.ms_residop = el(E_LOG, E_RESIDOP);
        I.ms_gram = el(E_RESIDOP, E_GRAM);
        I.ms_allreduce_gram = el(E_GRAM, E_AR1);
        I.ms_eigen = el(E_AR1, E_EIG);
        I.ms_bcast = el(E_EIG, E_BCAST);
        I.ms_project = el(E_BCAST, E_PROJECT);
        I.ms_ctl = el(E_PROJECT, E_CTL);
        I.ms_allreduce_ctl = el(E_CTL, E_AR2);
        I.ms_solve = el(E_AR2, E_SOLVE);
        I.ms_update = el(E_SOLVE, E_UPDATE);
        I.ms_total = el(E_START, E_UPDATE);
        if (!skip) { float ms = 0.f; if (cudaEventElapsedTime(&ms, s.ev_g0, s.ev_g1) == cudaSuccess) I.ms_gather = ms; else cudaGetLastError(); }
        if (!skip && !h->have_user_alpha && h->eig_solver_used == 3 && h->persist_rb) {
            I.eig_rr = h->h_flags[64 + 4];
            I.eig_chol2 = h->h_flags[64 + 5];
            I.eig_filter_products = h->h_flags[64 + 6];
        }
        if (!skip && !h->have_user_alpha) {
            std::vector<double> ev(h->kk + 1, 0.0);
            const int have = std::min(h->kk + 1, h->n_keep);
            RUVIII_CUDA_CHECK(cudaMemcpy(ev.data(), s.evals.p, have * sizeof(double), cudaMemcpyDeviceToHost));
            I.eig_top = ev[0];
            I.eig_kmax = ev[h->kk - 1];
            I.eig_next = have > h->kk ? ev[h->kk] : 0.0;
        }
    }
    h->ran = true;
    if (info_out) *info_out = I;
}

// K5c ruvIII.R:144 solve(ac %*% t(ac)) + centring (ruvIII.R:118-120) + A~ = A_c L^-T + a~ = L^-1 alpha.  One cooperative kernel
// (solve.cu); RUVIII_SOLVE_FUSED=0 selects the three-kernel form.
void launch_solve(ruviii_handle* h, Shard& s) {
    const int kp = h->kp, kk = h->kk;
    const double* A = s.AG.p;
    const double* Gm = s.AG.p + (size_t)h->rows_w * kp;
    static const bool fused = env_int("RUVIII_SOLVE_FUSED", 1) != 0;
    if (fused) {
        launch_solve_fused(A, h->rows_w, h->mu_rows, Gm, kk, kp, s.alpha.p, s.n, s.ld, s.solve_ws.p, s.Linv.p, s.colmean.p, s.status.p,
                           s.At.p, s.at.p, h->n_sms, s.stream);
        h->launches += 1;
    } else {
        launch_chol_solve(A, h->rows_w, h->mu_rows, Gm, kk, kp, s.Linv.p, s.colmean.p, s.status.p, s.stream);
        launch_atilde(A, h->rows_w, s.colmean.p, s.Linv.p, kk, kp, s.At.p, s.stream);
        launch_alpha_tilde(s.alpha.p, s.Linv.p, kk, kp, s.n, s.ld, s.at.p, s.stream);
        h->launches += 3;
    }
}

void launch_w_blocks(ruviii_handle* h, Shard& s) {
    size_t off = 0;
    for (size_t i = 0; i < h->ks.size(); ++i) {
        launch_w_from_atilde(s.At.p, h->rows_w, s.Linv.p, h->k_eff[i], h->kp, s.W.p + off, s.stream);
        off += (size_t)h->rows_w * h->k_eff[i];
    }
}

void do_download(ruviii_handle* h, double* newY, double* newY_ps, double* fullalpha, double* W, double* const* newY_blocks) {
    if (!h->ran) throw Error(RUVIII_ERR_STATE, "ruviii_download before ruviii_run");
    const bool fast = h->prm.mode == RUVIII_MODE_FAST;
    const bool skip = h->no_replicates || h->kk == 0;
    const int nk = (int)h->ks.size();
    const int64_t n = h->n_local_total;
    for (auto& s : h->shards) {
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_START], s.stream));
        if ((newY || newY_blocks) && h->m1 > 0)
            for (int i = 0; i < nk; ++i)
                RUVIII_CUDA_CHECK(cudaMemcpy2DAsync((newY_blocks ? newY_blocks[i] : newY + (size_t)i * h->m1 * n) + s.g0, n * 8,
                                                    s.out.p + (size_t)i * h->m1 * s.ld, s.ld * 8, (size_t)s.n * 8, h->m1,
                                                    cudaMemcpyDeviceToHost, s.stream));
        if (newY_ps) {
            if (fast || !h->prm.want_ps_rows) throw Error(RUVIII_ERR_INVALID, "newY_ps_out needs STACKED mode with want_ps_rows");
            for (int i = 0; i < nk; ++i)
                RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(newY_ps + (size_t)i * h->n_ps * n + s.g0, n * 8,
                                                    s.out_ps.p + (size_t)i * h->n_ps * s.ld, s.ld * 8, (size_t)s.n * 8,
                                                    h->n_ps, cudaMemcpyDeviceToHost, s.stream));
        }
        if (fullalpha && !skip) {
            const double* src = (!h->have_user_alpha && h->n_alpha_rows > h->kk) ? s.falpha.p : s.alpha.p;
            const int rows = h->have_user_alpha ? h->kk : h->n_alpha_rows;
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(fullalpha + s.g0, n * 8, src, s.ld * 8, (size_t)s.n * 8, rows,
                                                cudaMemcpyDeviceToHost, s.stream));
        }
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_LOG], s.stream));
    }
    if (W && !skip) {
        Shard& s = h->shards[0];
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        launch_w_blocks(h, s);           // W_k = A~[:, :k] L^-1[:k, :k]: formed only when somebody asks for it (return.info)
        size_t w_total = 0;
        for (int i = 0; i < nk; ++i) w_total += (size_t)h->rows_w * h->k_eff[i];
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(W, s.W.p, w_total * 8, cudaMemcpyDeviceToHost, s.stream));
    }
    sync_all(h);
    float ms = 0.f;
    RUVIII_CUDA_CHECK(cudaSetDevice(h->shards[0].device));
    cudaEventElapsedTime(&ms, h->shards[0].ev[E_START], h->shards[0].ev[E_LOG]);
    h->info.ms_d2h = ms;
}

}  // namespace ruviii

// ======================================== C ABI ==========================================================
#pragma GCC visibility push(default)
extern "C" {

int ruviii_abi_version(void) { return RUVIII_ABI_VERSION; }

int ruviii_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.major == 10) ++ok;
    }
    return ok;
}

int ruviii_create(ruviii_handle** out, const int32_t* devices, int32_t ndev) {
    if (!out || ndev <= 0) { g_create_error = "ruviii_create: bad arguments"; return RUVIII_ERR_INVALID; }
    *out = nullptr;
    auto h = std::make_unique<ruviii_handle>();
    int rc = guarded(nullptr, [&] {
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
            throw Error(RUVIII_ERR_CUDA, "no CUDA device: the RUV-III engine has no CPU fallback");
        h->shards.resize(ndev);
        std::vector<int> devs(ndev);
        for (int i = 0; i < ndev; ++i) {
            devs[i] = devices ? devices[i] : i;
            if (devs[i] < 0 || devs[i] >= count) throw Error(RUVIII_ERR_INVALID, "device index out of range");
            init_shard(h.get(), h->shards[i], devs[i]);
        }
        h->world = ndev;
        h->rank0 = 0;
        if (ndev > 1) {
            std::vector<ncclComm_t> comms(ndev);
            RUVIII_NCCL_CHECK(ncclCommInitAll(comms.data(), ndev, devs.data()));
            for (int i = 0; i < ndev; ++i) h->shards[i].comm = comms[i];
        }
        finish_create(h.get());
    });
    if (rc == RUVIII_OK) *out = h.release();
    return rc;
}

int ruviii_nccl_unique_id(void* out128) {
    if (!out128) return RUVIII_ERR_INVALID;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    return guarded(nullptr, [&] {
        ncclUniqueId id;
        RUVIII_NCCL_CHECK(ncclGetUniqueId(&id));
        std::memcpy(out128, &id, sizeof(id));
    });
}

int ruviii_create_rank(ruviii_handle** out, int32_t device, int32_t rank, int32_t nranks, const void* id128) {
    if (!out || nranks <= 0 || rank < 0 || rank >= nranks || (nranks > 1 && !id128)) {
        g_create_error = "ruviii_create_rank: bad arguments";
        return RUVIII_ERR_INVALID;
    }
    *out = nullptr;
    auto h = std::make_unique<ruviii_handle>();
    int rc = guarded(nullptr, [&] {
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
            throw Error(RUVIII_ERR_CUDA, "no CUDA device: the RUV-III engine has no CPU fallback");
        if (device < 0 || device >= count) throw Error(RUVIII_ERR_INVALID, "device index out of range");
        h->shards.resize(1);
        init_shard(h.get(), h->shards[0], device);
        h->world = nranks;
        h->rank0 = rank;
        h->multi_process = nranks > 1;
        if (nranks > 1) {
            ncclUniqueId id;
            std::memcpy(&id, id128, sizeof(id));
            RUVIII_NCCL_CHECK(ncclCommInitRank(&h->shards[0].comm, nranks, id, rank));
        }
        finish_create(h.get());
    });
    if (rc == RUVIII_OK) *out = h.release();
    return rc;
}

void ruviii_destroy(ruviii_handle* h) { delete h; }

const char* ruviii_last_error(const ruviii_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int ruviii_plan(ruviii_handle* h, const ruviii_params* params, int32_t m1, int32_t n, int32_t n_ps,
                const int32_t* group_of_row, const uint8_t* ctl, const int32_t* ks, int32_t nk) {
    if (!h) return RUVIII_ERR_INVALID;
    return guarded(h, [&] { do_plan(h, params, m1, n, n_ps, group_of_row, ctl, ks, nk); });
}

int ruviii_upload(ruviii_handle* h, const double* Y, int64_t ldY, const double* PS, int64_t ldPS) {
    if (!h) return RUVIII_ERR_INVALID;
    return guarded(h, [&] { do_upload(h, Y, ldY, PS, ldPS); });
}

int ruviii_run(ruviii_handle* h, ruviii_info* info) {
    if (!h) return RUVIII_ERR_INVALID;
    return guarded(h, [&] { do_run(h, info); });
}

int ruviii_download(ruviii_handle* h, double* newY_out, double* newY_ps_out, double* fullalpha_out, double* W_out) {
    if (!h) return RUVIII_ERR_INVALID;
    return guarded(h, [&] { do_download(h, newY_out, newY_ps_out, fullalpha_out, W_out); });
}

int ruviii_download_averaged(ruviii_handle* h, double* out) {
    if (!h || !out) return RUVIII_ERR_INVALID;
    return guarded(h, [&] {
        if (!h->ran) throw Error(RUVIII_ERR_STATE, "ruviii_download_averaged before ruviii_run");
        if (!h->prm.average_rep) throw Error(RUVIII_ERR_STATE, "the plan was made without average_rep");
        const int nk = (int)h->ks.size();
        const int64_t n = h->n_local_total;
        for (auto& s : h->shards) {
            RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
            for (int i = 0; i < nk; ++i)
                RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(out + (size_t)i * h->p * n + s.g0, n * 8, s.avg.p + (size_t)i * h->p * s.ld,
                                                    s.ld * 8, (size_t)s.n * 8, h->p, cudaMemcpyDeviceToHost, s.stream));
        }
        sync_all(h);
    });
}

int ruviii_download_replicate_data(ruviii_handle* h, double* PS_out, int64_t ldPS) {
    if (!h || !PS_out) return RUVIII_ERR_INVALID;
    return guarded(h, [&] {
        if (!h->planned || !h->uploaded) throw Error(RUVIII_ERR_STATE, "ruviii_download_replicate_data needs a plan and an upload");
        if (h->n_ps <= 0) return;
        if (ldPS < h->n_local_total) throw Error(RUVIII_ERR_INVALID, "row stride < n");
        if (h->ps_pending) {               // the sets are registered but no run has built the rows yet
            ensure_logged(h);
            for (auto& s : h->shards) {
                RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
                RowSrc src{s.Y.p, s.ld, nullptr, 0, h->m1};
                launch_group_mean(src, s.set_start.p, s.set_rows.p, h->n_ps, s.n, s.PS.p, s.ld, s.stream);
            }
            h->ps_pending = false;
        }
        for (auto& s : h->shards) {
            RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(PS_out + s.g0, ldPS * 8, s.PS.p, s.ld * 8, (size_t)s.n * 8, h->n_ps, cudaMemcpyDeviceToHost,
                                                s.stream));
        }
        sync_all(h);
    });
}

int ruviii_eigenvalues(ruviii_handle* h, double* out, int32_t count) {
    if (!h || !out) return RUVIII_ERR_INVALID;
    return guarded(h, [&] {
        if (!h->ran || h->no_replicates || h->kk == 0 || h->have_user_alpha)
            throw Error(RUVIII_ERR_STATE, "no eigenvalues: run did not decompose anything");
        if (count < 0 || count > h->n_keep) throw Error(RUVIII_ERR_INVALID, "count exceeds the eigenvalues kept");
        RUVIII_CUDA_CHECK(cudaSetDevice(h->shards[0].device));
        RUVIII_CUDA_CHECK(cudaMemcpy(out, h->shards[0].evals.p, count * sizeof(double), cudaMemcpyDeviceToHost));
    });
}

int ruviii_set_fullalpha(ruviii_handle* h, const double* fullalpha, int32_t rows) {
    if (!h) return RUVIII_ERR_INVALID;
    return guarded(h, [&] {
        if (!h->planned) throw Error(RUVIII_ERR_STATE, "ruviii_set_fullalpha before ruviii_plan");
        if (!fullalpha || rows <= 0) throw Error(RUVIII_ERR_INVALID, "ruviii_set_fullalpha: bad arguments");
        if (h->no_replicates || h->kmax == 0) return;
        // alpha <- fullalpha[1:min(k, nrow(fullalpha)), ]  (ruvIII.R:142)
        const int kk = std::min(h->kmax, (int)rows);
        if (kk > kMaxK) throw Error(RUVIII_ERR_UNSUPPORTED, "k > 32");
        if (h->kk == 0)
            throw Error(RUVIII_ERR_UNSUPPORTED, "the plan found no usable factor (r = 0: no control genes?), so nothing was set up for "
                                                "a supplied fullalpha");
        h->kk = kk;
        const int64_t n = h->n_local_total;
        for (auto& s : h->shards) {
            RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
            RUVIII_CUDA_CHECK(cudaMemsetAsync(s.alpha.p, 0, (size_t)h->kp * s.ld * 8, s.stream));
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(s.alpha.p, s.ld * 8, fullalpha + s.g0, n * 8, (size_t)s.n * 8, kk,
                                                cudaMemcpyHostToDevice, s.stream));
            RUVIII_CUDA_CHECK(cudaStreamSynchronize(s.stream));
        }
        // effective k of every request shrinks accordingly
        const int nk = (int)h->ks.size();
        h->kmask = 0;
        h->dup_slots.clear();
        for (int i = 0; i <= kMaxK; ++i) h->slot_of_k[i] = -1;
        for (int i = 0; i < nk; ++i) {
            const int ke = std::min(h->ks[i], kk);
            h->k_eff[i] = ke;
            if (h->slot_of_k[ke] < 0) { h->slot_of_k[ke] = i; h->kmask |= 1ull << ke; }
            else h->dup_slots.emplace_back(i, h->slot_of_k[ke]);
        }
        for (int i = 0; i <= kMaxK; ++i) if (h->slot_of_k[i] < 0) h->slot_of_k[i] = 0;
        size_t w_total = 0;
        for (int i = 0; i < nk; ++i) w_total += (size_t)h->rows_w * h->k_eff[i];
        for (auto& s : h->shards) {
            RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
            s.W.alloc(std::max<size_t>(w_total, 1));
        }
        h->have_user_alpha = true;
    });
}

// RUVIII_PIPELINE: 0 / 1 = auto (staged path when eligible), 2 = never overlap (upload -> run -> download in sequence),
// 3 = the round-1 pipeline with the zero-copy control-gene gather (needs pinned buffers)
static int normalise_any(ruviii_handle* h, const ruviii_params* params, const double* Y, int64_t ldY, int32_t m1, int32_t n,
                         const double* PS, int64_t ldPS, int32_t n_ps, const int32_t* group_of_row, const uint8_t* ctl,
                         const int32_t* ks, int32_t nk, double* newY_out, double* const* newY_blocks, double* newY_ps_out,
                         double* fullalpha_out, double* W_out, ruviii_info* info) {
    if (!h) return RUVIII_ERR_INVALID;
    return guarded(h, [&] {
        if (!newY_out && !newY_blocks) throw Error(RUVIII_ERR_INVALID, "newY_out is NULL");
        const auto tp0 = std::chrono::steady_clock::now();
        do_plan(h, params, m1, n, n_ps, group_of_row, ctl, ks, nk);
        if (env_int("RUVIII_PIPE_TRACE", 0))
            fprintf(stderr, "[pipe trace] plan=%.2f ms\n",
                    1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - tp0).count());
        h->pipelined_last = 0;
        const int mode = env_int("RUVIII_PIPELINE", env_int("RUVIII_NO_PIPELINE", 0) ? 2 : 0);
        std::vector<double*> blocks(nk);
        for (int i = 0; i < nk; ++i) {
            blocks[i] = newY_blocks ? newY_blocks[i] : newY_out + (size_t)i * m1 * n;
            if (!blocks[i]) throw Error(RUVIII_ERR_INVALID, "a newY block is NULL");
        }
        if ((m1 > 0 && ldY < n) || (n_ps > 0 && PS && ldPS < n)) throw Error(RUVIII_ERR_INVALID, "row stride < n");
        if (mode != 2 && mode != 3 && can_stage(h, PS, newY_ps_out)) {
            do_normalise_staged(h, Y, ldY, PS, ldPS, blocks.data(), fullalpha_out, W_out);
        } else if (mode == 3 && !newY_blocks && can_pipeline(h, Y, PS, newY_out, newY_ps_out)) {
            do_normalise_pipelined(h, Y, ldY, PS, ldPS, newY_out, fullalpha_out, W_out);
        } else {
            do_upload(h, Y, ldY, PS, ldPS);
            do_run(h, nullptr);
            do_download(h, nullptr, newY_ps_out, fullalpha_out, W_out, blocks.data());
        }
        if (info) *info = h->info;
    });
}

int ruviii_normalise(ruviii_handle* h, const ruviii_params* params, const double* Y, int64_t ldY, int32_t m1, int32_t n,
                     const double* PS, int64_t ldPS, int32_t n_ps, const int32_t* group_of_row, const uint8_t* ctl,
                     const int32_t* ks, int32_t nk, double* newY_out, double* newY_ps_out, double* fullalpha_out,
                     double* W_out, ruviii_info* info) {
    return normalise_any(h, params, Y, ldY, m1, n, PS, ldPS, n_ps, group_of_row, ctl, ks, nk, newY_out, nullptr, newY_ps_out,
                         fullalpha_out, W_out, info);
}

int ruviii_normalise_blocks(ruviii_handle* h, const ruviii_params* params, const double* Y, int64_t ldY, int32_t m1, int32_t n,
                            const double* PS, int64_t ldPS, int32_t n_ps, const int32_t* group_of_row, const uint8_t* ctl,
                            const int32_t* ks, int32_t nk, double* const* newY_blocks, double* newY_ps_out, double* fullalpha_out,
                            double* W_out, ruviii_info* info) {
    if (!newY_blocks) return RUVIII_ERR_INVALID;
    return normalise_any(h, params, Y, ldY, m1, n, PS, ldPS, n_ps, group_of_row, ctl, ks, nk, nullptr, newY_blocks, newY_ps_out,
                         fullalpha_out, W_out, info);
}

void* ruviii_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void ruviii_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int ruviii_barrier(ruviii_handle* h) {
    if (!h) return RUVIII_ERR_INVALID;
    return guarded(h, [&] {
        if (h->world > 1) {
            std::vector<double*> ptrs;
            for (auto& s : h->shards) {
                RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
                s.status.alloc(4);
                ptrs.push_back(reinterpret_cast<double*>(s.status.p));     // 16 bytes, used as one double
            }
            allreduce_sum(h, ptrs, 1);
        }
        sync_all(h);
    });
}

}  // extern "C"
#pragma GCC visibility pop


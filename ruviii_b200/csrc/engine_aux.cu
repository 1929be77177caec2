// The steps either side of the path and the RUV1 pre-step: engine-level drivers and their C ABI.
#include "engine_internal.cuh"

namespace ruviii {

// ruv::RUV1 (ruvIII.R:116; fastruvIII.R:133-134 applies it to Y and to replicate.data): every row y of the stacked matrix
// becomes y - y[ctl] eta_c' (eta_c eta_c')^-1 eta.  Same shape as the RUV-III update itself with eta in the role of alpha
// and no centring, so it reuses K5/K6: A1 = Y[, ctl] eta_c' and eta_c eta_c' (allreduced over gene shards), Cholesky,
// then an in-place rank-q update of the device copies of Y and replicate.data.
void apply_ruv1(ruviii_handle* h) {
    const int q = h->eta_q, qp = pad_k(q), m = h->m, m1 = h->m1;
    int slot0[kMaxK + 1] = {0};
    for (auto& s : h->shards) {
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        RowSrc src{s.Y.p, s.ld, s.PS.p, s.ld, m1};
        launch_alpha_finalize(s.eta.p, 1, qp, s.n, s.ld, s.eta.p, s.ctl_idx.p, s.n_ctl, s.etacT.p, s.stream);
        launch_ctl_gather(src, m, s.ctl_idx.p, s.n_ctl, s.Yc1.p, s.ldc, 0, 0.0, s.stream);
        launch_ctl_products(src, m, s.ctl_idx.p, s.n_ctl, s.etacT.p, qp, s.AG1.p, s.AG1.p + (size_t)m * qp,
                            s.n_ctl > 0 ? s.Yc1.p : nullptr, s.ldc, s.stream);
    }
    {
        std::vector<double*> b;
        for (auto& s : h->shards) b.push_back(s.AG1.p);
        allreduce_sum(h, b, (size_t)m * qp + (size_t)qp * qp);
    }
    for (size_t i = 0; i < h->shards.size(); ++i) {
        Shard& s = h->shards[i];
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        const double* A = s.AG1.p;
        launch_chol_solve(A, m, 0, s.AG1.p + (size_t)m * qp, q, qp, s.Linv1.p, s.cm1.p, s.status1.p, s.stream);   // mu_rows = 0: no centring
        launch_atilde(A, m, s.cm1.p, s.Linv1.p, q, qp, s.At1.p, s.stream);
        launch_alpha_tilde(s.eta.p, s.Linv1.p, q, qp, s.n, s.ld, s.eta_t.p, s.stream);
        // The update kernels read their input through the non-coherent path (ld.global.nc) and declare input and output
        // __restrict__, so they must not run in place: the result goes to scratch (the first output block / PS1) and is
        // copied back.  Once per upload, and only when eta != NULL.
        if (m1 > 0) {
            launch_update_smem(s.Y.p, s.ld, m1, s.n, s.At1.p, qp, s.eta_t.p, s.ld, q, 1ull << q, slot0, s.out.p, s.ld, 0, s.stream);
            RUVIII_CUDA_CHECK(cudaMemcpyAsync(s.Y.p, s.out.p, (size_t)m1 * s.ld * 8, cudaMemcpyDeviceToDevice, s.stream));
        }
        if (h->n_ps > 0) {
            launch_update(s.PS.p, s.ld, h->n_ps, s.n, s.At1.p + (size_t)m1 * qp, qp, s.eta_t.p, s.ld, q, 1ull << q, slot0, s.PS1.p,
                          s.ld, 0, s.stream);
            RUVIII_CUDA_CHECK(cudaMemcpyAsync(s.PS.p, s.PS1.p, (size_t)h->n_ps * s.ld * 8, cudaMemcpyDeviceToDevice, s.stream));
        }
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(&h->h_flags[32 + i], s.status1.p, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    }
    h->launches += 8 + (h->n_ps > 0 ? 1 : 0);
    for (size_t i = 0; i < h->shards.size(); ++i) {
        RUVIII_CUDA_CHECK(cudaSetDevice(h->shards[i].device));
        spin_sync(h->shards[i].stream);
        if (h->h_flags[32 + i] != 0)
            throw Error(RUVIII_ERR_SINGULAR, "etac %*% t(etac) is not positive definite at pivot " +
                                                 std::to_string(h->h_flags[32 + i]) + " (solve() in ruv::RUV1 would fail)");
    }
}

// The k > 32 tail of ruviii_pca: the k leading eigenpairs sit in pG (column-major, ascending) / pEvAsc on rank 0; per chunk of
// 32: sign-fixed descending columns (pick_top) -> broadcast -> K4 projection -> v = t(alpha) / d.
static void do_pca_wide_tail(ruviii_handle* h, bool resident, int m1, int n, int k, double* u_out, double* d_out, double* v_out) {
    const int nsh = (int)h->shards.size(), kp = kMaxK;
    std::vector<double> ev(k);
    for (int c0 = 0; c0 < k; c0 += kp) {
        const int kc = std::min(kp, k - c0);
        if (h->rank0 == 0) {
            Shard& s = h->shards[0];
            RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
            launch_pick_top(s.pG.p, m1, s.pEvAsc.p, k - c0, m1, kc, kc, kp, nullptr, s.pU.p, s.pEv.p, s.stream);
            h->launches += 1;
        }
        if (c0 == 0)
            for (auto& s : h->shards) {
                RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
                RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_EIG], s.stream));
            }
        {
            std::vector<double*> b;
            for (auto& s : h->shards) b.push_back(s.pU.p);
            broadcast0(h, b, (size_t)m1 * kp);
            b.clear();
            for (auto& s : h->shards) b.push_back(s.pEv.p);
            broadcast0(h, b, (size_t)kp);
        }
        for (int i = 0; i < nsh; ++i) {
            Shard& s = h->shards[i];
            RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
            const int g0 = resident ? s.g0 : (int)((int64_t)n * i / nsh);
            const int ng = resident ? s.n : (int)((int64_t)n * (i + 1) / nsh) - g0;
            const int64_t ld = resident ? s.ld : round_up(std::max(ng, 1), kRowAlign);
            const int rsplit = project_rsplit(m1, ng, h->n_sms);
            launch_project(s.pX.p, ld, m1, ng, s.pU.p, kp, rsplit, s.pPart.p, ld, s.stream, kc);
            launch_alpha_finalize(s.pPart.p, rsplit, kp, ng, ld, s.pAlpha.p, nullptr, 0, nullptr, s.stream);
            launch_v_from_alpha(s.pAlpha.p, s.pEv.p, kc, ld, nullptr, s.stream);
            h->launches += 3;
            if (v_out && ng > 0)
                RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(v_out + (size_t)c0 * n + g0, (size_t)n * 8, s.pAlpha.p, ld * 8, (size_t)ng * 8, kc,
                                                    cudaMemcpyDeviceToHost, s.stream));
            if (i == 0) {
                RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(u_out + c0, (size_t)k * 8, s.pU.p, (size_t)kp * 8, (size_t)kc * 8, m1,
                                                    cudaMemcpyDeviceToHost, s.stream));
                RUVIII_CUDA_CHECK(cudaMemcpyAsync(ev.data() + c0, s.pEv.p, (size_t)kc * 8, cudaMemcpyDeviceToHost, s.stream));
            }
        }
        sync_all(h);                      // pU / pEv / pAlpha are reused by the next chunk
    }
    for (auto& s : h->shards) {
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_PROJECT], s.stream));
    }
    sync_all(h);
    for (int j = 0; j < k; ++j) d_out[j] = std::sqrt(std::max(ev[j], 0.0));
}

// ---- computePCA fast path --------------------------------------------------------------------------
// computePCA.R:120-134: runSVD(t(assay), k = nb.pcs, center, scale) -- the top-k singular triplets of the centred samples x
// genes matrix.  Sample-side route with the engine's own kernels: K10 centre -> K2 Gram Xc Xc' (genes sharded, allreduced)
// -> K3 top-k eigenpairs (u, d^2) on rank 0 -> broadcast -> K4 projection t(u) Xc = diag(d) t(v).
// source: X != NULL host matrix; X == NULL and slot >= 0: output block `slot` of the last run, already on the device
// (normAssessment.R:113 runs computePCA on every RUV_K* assay right after normalise()); slot == -1: the resident input.
void do_pca(ruviii_handle* h, const double* X, int64_t ldX, int m1, int n, int slot, int apply_log, double pseudo_count,
            int center, int scale, int k, double* u_out, double* d_out, double* v_out, float* stage_ms) {
    const bool resident = X == nullptr;
    if (resident) {
        if (!h->planned || !h->uploaded) throw Error(RUVIII_ERR_STATE, "ruviii_pca on resident data needs a plan and an upload");
        if (slot >= 0 && (!h->ran || slot >= (int)h->ks.size())) throw Error(RUVIII_ERR_STATE, "ruviii_pca: no such output block");
        m1 = h->m1;
        n = h->n_local_total;
        if (slot < 0) { ensure_logged(h); apply_log = 0; }       // the resident assay after the plan's own log transform
    }
    if (m1 < 2 || n <= 0 || k <= 0 || !u_out || !d_out) throw Error(RUVIII_ERR_INVALID, "ruviii_pca: bad arguments");
    if (!resident && ldX < n) throw Error(RUVIII_ERR_INVALID, "ruviii_pca: row stride < n");
    // k <= 32: the fast path of computePCA (runSVD, :120-134).  Wider requests -- up to min(m1, n), which is the full svd() of
    // computePCA.R:150-190 (fast.pca = FALSE) -- take the same route with the library eigensolver for the k leading pairs
    // and the projection in chunks of 32 columns.
    const bool wide = k > kMaxK;
    if (k > std::min(m1, n) || (!wide && k >= m1))
        throw Error(RUVIII_ERR_UNSUPPORTED, "ruviii_pca: nb.pcs must be < the number of samples (<= min(samples, genes) beyond 32)");
    const int kp = wide ? kMaxK : pad_k(k), nsh = (int)h->shards.size();
    const int n_keep = wide ? k : std::min(m1, k + 1);
    const bool topk_ok = !wide && k <= kTopkMaxK && m1 >= 4 * kTopkBlock && h->eig_impl != 1 && h->eig_impl != 2;
    h->launches = 0;
    std::vector<GramPlan> gps(nsh);
    for (int i = 0; i < nsh; ++i) {
        Shard& s = h->shards[i];
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        const int g0 = resident ? s.g0 : (int)((int64_t)n * i / nsh);
        const int ng = resident ? s.n : (int)((int64_t)n * (i + 1) / nsh) - g0;
        const int64_t ld = resident ? s.ld : round_up(std::max(ng, 1), kRowAlign);
        const int blocks = colstat_blocks(m1, ng, h->n_sms);
        s.pX.alloc((size_t)m1 * ld, false);
        s.pStat.alloc((size_t)(blocks + 1) * ld, false);
        s.pMean.alloc(ld);
        s.pSd.alloc(ld);
        GramPlan& gp = gps[i];
        gp = plan_gram(m1, (int)ld, h->n_sms, h->gram_impl == 2 ? 1 : h->gram_impl);     // cp.async kernels (no tensor map to build)
        s.pWs.alloc(gp.workspace_bytes / sizeof(double), false);
        std::vector<int2> tm(gp.n_tiles);
        fill_gram_tile_map(gp, tm.data());
        s.pTiles.alloc(tm.size());
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(s.pTiles.p, tm.data(), tm.size() * sizeof(int2), cudaMemcpyHostToDevice, s.stream));
        s.pG.alloc((size_t)m1 * m1, false);
        s.pU.alloc((size_t)m1 * kp);
        s.pEv.alloc(std::max(n_keep, kp) + 2);
        const int rsplit = project_rsplit(m1, ng, h->n_sms);
        s.pPart.alloc((size_t)rsplit * kp * ld, false);
        s.pAlpha.alloc((size_t)kp * ld);
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_START], s.stream));
        const double* src = nullptr;
        if (!resident) {
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(s.pX.p, ld * 8, X + g0, ldX * 8, (size_t)ng * 8, m1, cudaMemcpyHostToDevice, s.stream));
            if (apply_log) { launch_log2_inplace(s.pX.p, ld, m1, ng, pseudo_count, s.stream); h->launches += 1; }   // computePCA.R:124-127
            src = s.pX.p;
        } else {
            src = slot >= 0 ? s.out.p + (size_t)slot * m1 * s.ld : s.Y.p;
            if (apply_log) {                          // a resident block is copied before it is transformed: outputs stay intact
                RUVIII_CUDA_CHECK(cudaMemcpyAsync(s.pX.p, src, (size_t)m1 * ld * 8, cudaMemcpyDeviceToDevice, s.stream));
                launch_log2_inplace(s.pX.p, ld, m1, ng, pseudo_count, s.stream);
                h->launches += 1;
                src = s.pX.p;
            }
        }
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_LOG], s.stream));
        h->launches += launch_center_columns(src, ld, m1, ng, center, scale, blocks, s.pStat.p, s.pMean.p, s.pSd.p, s.pX.p, ld, s.stream);
        // the pitch padding takes part in the Gram's contraction: clear it after the centring (an output block's padding
        // is never written by the update kernel, so what the centring read there is arbitrary)
        if (ld > ng)
            RUVIII_CUDA_CHECK(cudaMemset2DAsync(s.pX.p + ng, ld * 8, 0, (size_t)(ld - ng) * 8, m1, s.stream));
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_RESIDOP], s.stream));
        h->launches += launch_gram(gp, s.pX.p, ld, s.pTiles.p, s.pWs.p, s.pG.p, nullptr, s.stream);
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_GRAM], s.stream));
    }
    {
        std::vector<double*> b;
        for (auto& s : h->shards) b.push_back(s.pG.p);
        allreduce_sum(h, b, (size_t)m1 * m1);
    }
    if (h->rank0 == 0) {
        Shard& s = h->shards[0];
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        bool done = false;
        if (topk_ok) {
            s.pTopk.alloc(topk_work_doubles(m1, kTopkBlock), false);
            s.pState.alloc(4);
            topk_bind(s.ptopk, s.pTopk.p, s.pState.p, m1, kTopkBlock);
            TopkResult tr = topk_eigen(s.pG.p, m1, k, kp, n_keep, kTopkBlock, s.ptopk, s.pU.p, s.pEv.p, h->eig_tol,
                                       env_int("RUVIII_EIG_MAXIT", k > kTopkBlock / 2 ? kTopkPatience : 150), h->n_sms,
                                       &h->h_flags[60], s.stream);
            h->launches += tr.launches;
            h->eig_iterations = tr.iterations;
            done = tr.converged;
        }
        if (!done) {                                   // cuSOLVER: k > 28, small matrices, or no convergence
            if (!h->solver) {
                RUVIII_SOLVER_CHECK(cusolverDnCreate(&h->solver));
                RUVIII_SOLVER_CHECK(cusolverDnSetStream(h->solver, s.stream));
            }
            s.pEvAsc.alloc(m1);
            int lwork = 0, meig = 0;
            RUVIII_SOLVER_CHECK(cusolverDnDsyevdx_bufferSize(h->solver, CUSOLVER_EIG_MODE_VECTOR, CUSOLVER_EIG_RANGE_I,
                                                             CUBLAS_FILL_MODE_UPPER, m1, s.pG.p, m1, 0.0, 0.0, m1 - n_keep + 1, m1,
                                                             &meig, s.pEvAsc.p, &lwork));
            h->solver_work.alloc((size_t)lwork, false);
            h->solver_info.alloc(4);
            RUVIII_SOLVER_CHECK(cusolverDnDsyevdx(h->solver, CUSOLVER_EIG_MODE_VECTOR, CUSOLVER_EIG_RANGE_I, CUBLAS_FILL_MODE_UPPER,
                                                  m1, s.pG.p, m1, 0.0, 0.0, m1 - n_keep + 1, m1, &meig, s.pEvAsc.p,
                                                  h->solver_work.p, (int)h->solver_work.count, h->solver_info.p));
            if (meig != n_keep) throw Error(RUVIII_ERR_SOLVER, "syevdx returned an unexpected number of eigenpairs");
            if (!wide) {
                launch_pick_top(s.pG.p, m1, s.pEvAsc.p, meig, m1, n_keep, k, kp, nullptr, s.pU.p, s.pEv.p, s.stream);
                h->launches += 1;
            }
        }
    }
    if (wide) {
        do_pca_wide_tail(h, resident, m1, n, k, u_out, d_out, v_out);
        if (stage_ms) {
            Shard& s = h->shards[0];
            RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
            auto el = [&](int a, int b) { float ms = 0.f; cudaEventElapsedTime(&ms, s.ev[a], s.ev[b]); return ms; };
            stage_ms[0] = el(E_START, E_LOG); stage_ms[1] = el(E_LOG, E_RESIDOP); stage_ms[2] = el(E_RESIDOP, E_GRAM);
            stage_ms[3] = el(E_GRAM, E_EIG); stage_ms[4] = el(E_EIG, E_PROJECT); stage_ms[5] = el(E_START, E_PROJECT);
        }
        return;
    }
    for (auto& s : h->shards) {
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_EIG], s.stream));
    }
    {
        std::vector<double*> b;
        for (auto& s : h->shards) b.push_back(s.pU.p);
        broadcast0(h, b, (size_t)m1 * kp);
        b.clear();
        for (auto& s : h->shards) b.push_back(s.pEv.p);
        broadcast0(h, b, (size_t)std::max(n_keep, kp) + 2);
    }
    std::vector<double> ev(k);
    for (int i = 0; i < nsh; ++i) {
        Shard& s = h->shards[i];
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        const int g0 = resident ? s.g0 : (int)((int64_t)n * i / nsh);
        const int ng = resident ? s.n : (int)((int64_t)n * (i + 1) / nsh) - g0;
        const int64_t ld = resident ? s.ld : round_up(std::max(ng, 1), kRowAlign);
        const int rsplit = project_rsplit(m1, ng, h->n_sms);
        launch_project(s.pX.p, ld, m1, ng, s.pU.p, kp, rsplit, s.pPart.p, ld, s.stream, k);
        launch_alpha_finalize(s.pPart.p, rsplit, kp, ng, ld, s.pAlpha.p, nullptr, 0, nullptr, s.stream);
        launch_v_from_alpha(s.pAlpha.p, s.pEv.p, k, ld, nullptr, s.stream);
        h->launches += 3;
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_PROJECT], s.stream));
        if (v_out)
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(v_out + g0, (size_t)n * 8, s.pAlpha.p, ld * 8, (size_t)ng * 8, k, cudaMemcpyDeviceToHost,
                                                s.stream));
        if (i == 0) {
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(u_out, (size_t)k * 8, s.pU.p, (size_t)kp * 8, (size_t)k * 8, m1, cudaMemcpyDeviceToHost,
                                                s.stream));
            RUVIII_CUDA_CHECK(cudaMemcpyAsync(ev.data(), s.pEv.p, (size_t)k * 8, cudaMemcpyDeviceToHost, s.stream));
        }
    }
    sync_all(h);
    for (int j = 0; j < k; ++j) d_out[j] = std::sqrt(std::max(ev[j], 0.0));
    if (stage_ms) {
        Shard& s = h->shards[0];
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        auto el = [&](int a, int b) { float ms = 0.f; cudaEventElapsedTime(&ms, s.ev[a], s.ev[b]); return ms; };
        stage_ms[0] = el(E_START, E_LOG);       // upload (+ log)
        stage_ms[1] = el(E_LOG, E_RESIDOP);     // K10 centre / scale
        stage_ms[2] = el(E_RESIDOP, E_GRAM);    // K2 Gram
        stage_ms[3] = el(E_GRAM, E_EIG);        // allreduce + K3 eigensolve
        stage_ms[4] = el(E_EIG, E_PROJECT);     // broadcast + K4 projection
        stage_ms[5] = el(E_START, E_PROJECT);
    }
}

// ---- regress.out.*  ------------------------------------------------------------------------------------
// supervisedFindNCG.R:111-133: y <- lm(t(expr.data) ~ variables); expr.data.reg <- t(y$residuals).  The model matrix (m1 x p:
// intercept, treatment dummies of the factors, the continuous variables) comes from the host; K12 turns it into an orthonormal
// basis Q with lm's rank rule, then per chunk of <= 32 columns K4 gives Q'Y and K6 gives Y - Q (Q'Y).  The residuals stay on the
// device (the statistics address them as slot = -2) and are downloaded only when `out` is given.
void do_regress_out(ruviii_handle* h, const double* X, int64_t ldX, int m1, int n, int slot, int apply_log, double pseudo_count,
                    const double* design, int p, double* out, int64_t ld_out, int* rank_out, float* ms_out) {
    const bool resident = X == nullptr;
    if (resident) {
        if (!h->planned || !h->uploaded) throw Error(RUVIII_ERR_STATE, "ruviii_regress_out on resident data needs a plan and an upload");
        if (slot >= 0 && (!h->ran || slot >= (int)h->ks.size())) throw Error(RUVIII_ERR_STATE, "ruviii_regress_out: no such output block");
        if (slot < -1) throw Error(RUVIII_ERR_INVALID, "ruviii_regress_out: slot must be -1 or an output block");
        m1 = h->m1;
        n = h->n_local_total;
        if (slot < 0) { ensure_logged(h); apply_log = 0; }       // the resident assay after the plan's own log transform
    }
    if (m1 < 2 || n <= 0 || !design || p <= 0) throw Error(RUVIII_ERR_INVALID, "ruviii_regress_out: bad arguments");
    if (p > 256 || p >= m1) throw Error(RUVIII_ERR_UNSUPPORTED, "ruviii_regress_out: the model matrix must have <= 256 columns and fewer columns than samples");
    if (!resident && ldX < n) throw Error(RUVIII_ERR_INVALID, "ruviii_regress_out: row stride < n");
    if (out && ld_out < n) throw Error(RUVIII_ERR_INVALID, "ruviii_regress_out: output row stride < n");
    const int nsh = (int)h->shards.size(), n_chunks = ceil_div(p, 32);
    int slot0[kMaxK + 1] = {0};
    h->reg_valid = false;
    h->launches = 0;
    for (int i = 0; i < nsh; ++i) {
        Shard& s = h->shards[i];
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        const int g0 = resident ? s.g0 : (int)((int64_t)n * i / nsh);
        const int ng = resident ? s.n : (int)((int64_t)n * (i + 1) / nsh) - g0;
        const int64_t ld = resident ? s.ld : round_up(std::max(ng, 1), kRowAlign);
        s.reg_g0 = g0; s.reg_n = ng; s.reg_ld = ld;
        if (ng <= 0) continue;
        s.rD.alloc((size_t)m1 * p, false);
        s.rW.alloc((size_t)m1 * p, false);
        s.rQ.alloc((size_t)n_chunks * m1 * 32, false);
        s.rRank.alloc(1);
        const int rsplit = project_rsplit(m1, ng, h->n_sms);
        s.rPart.alloc((size_t)rsplit * 32 * ld, false);
        s.rAlpha.alloc((size_t)32 * ld, false);
        s.gReg.alloc((size_t)m1 * ld, false);
        if (n_chunks > 1) s.gReg2.alloc((size_t)m1 * ld, false);
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(s.rD.p, design, (size_t)m1 * p * 8, cudaMemcpyHostToDevice, s.stream));
        const double* src;
        if (!resident) {
            s.gX.alloc((size_t)m1 * ld, false);
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(s.gX.p, ld * 8, X + g0, ldX * 8, (size_t)ng * 8, m1, cudaMemcpyHostToDevice, s.stream));
            if (apply_log) { launch_log2_inplace(s.gX.p, ld, m1, ng, pseudo_count, s.stream); h->launches += 1; }
            src = s.gX.p;
        } else {
            src = slot >= 0 ? s.out.p + (size_t)slot * m1 * s.ld : s.Y.p;
            if (apply_log) {                              // an output block is copied before it is transformed
                s.gX.alloc((size_t)m1 * ld, false);
                RUVIII_CUDA_CHECK(cudaMemcpyAsync(s.gX.p, src, (size_t)m1 * ld * 8, cudaMemcpyDeviceToDevice, s.stream));
                launch_log2_inplace(s.gX.p, ld, m1, ng, pseudo_count, s.stream);
                h->launches += 1;
                src = s.gX.p;
            }
        }
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_START], s.stream));
        launch_design_qr(s.rD.p, p, m1, p, 1e-7, s.rW.p, s.rQ.p, s.rRank.p, s.stream);          // lm's default tol = 1e-07
        h->launches += 1;
        for (int c = 0; c < n_chunks; ++c) {
            const int kc = std::min(32, p - 32 * c), kp = pad_k(kc);
            const double* q = s.rQ.p + (size_t)c * m1 * 32;
            double* dst = ((n_chunks - 1 - c) % 2 == 0) ? s.gReg.p : s.gReg2.p;                  // the last chunk lands in gReg
            launch_project(src, ld, m1, ng, q, kp, rsplit, s.rPart.p, ld, s.stream, kc);
            launch_alpha_finalize(s.rPart.p, rsplit, kp, ng, ld, s.rAlpha.p, nullptr, 0, nullptr, s.stream);
            launch_update_smem(src, ld, m1, ng, q, kp, s.rAlpha.p, ld, kc, 1ull << kc, slot0, dst, ld, 0, s.stream);
            h->launches += 3;
            src = dst;
        }
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_LOG], s.stream));
        if (out)
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(out + g0, (size_t)ld_out * 8, s.gReg.p, ld * 8, (size_t)ng * 8, m1, cudaMemcpyDeviceToHost,
                                                s.stream));
        if (i == 0) RUVIII_CUDA_CHECK(cudaMemcpyAsync(&h->h_flags[56], s.rRank.p, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    }
    sync_all(h);
    h->reg_valid = true;
    h->reg_m1 = m1;
    h->reg_n = n;
    if (rank_out) *rank_out = h->h_flags[56];
    if (ms_out) {
        RUVIII_CUDA_CHECK(cudaSetDevice(h->shards[0].device));
        cudaEventElapsedTime(ms_out, h->shards[0].ev[E_START], h->shards[0].ev[E_LOG]);
    }
}

// ---- negative-control-gene statistics ---------------------------------------------------------------
// supervisedFindNCG.R:179-191,326-341 (row_oneway_equalvar(x, g)$statistic) and :214-229,270-285 (correls(y, x, "pearson")).
// One streaming pass per call (K11, stats.cu).  Samples with group id < 0 are left out (the reference drops levels with
// fewer than three samples, :183-185); without groups every sample takes part and only the correlation is produced.
void do_gene_stats(ruviii_handle* h, const double* X, int64_t ldX, int m1, int n, int slot, int apply_log, double pseudo_count,
                   const int32_t* group, const double* covariate, double* F_out, double* r_out, float* ms_out) {
    const bool resident = X == nullptr;
    const bool regressed = resident && slot == -2;               // the residuals the last ruviii_regress_out left on the device
    if (regressed) {
        if (!h->reg_valid) throw Error(RUVIII_ERR_STATE, "ruviii_gene_stats: slot -2 needs a ruviii_regress_out first");
        m1 = h->reg_m1;
        n = h->reg_n;
        apply_log = 0;
    } else if (resident) {
        if (!h->planned || !h->uploaded) throw Error(RUVIII_ERR_STATE, "ruviii_gene_stats on resident data needs a plan and an upload");
        if (slot >= 0 && (!h->ran || slot >= (int)h->ks.size())) throw Error(RUVIII_ERR_STATE, "ruviii_gene_stats: no such output block");
        m1 = h->m1;
        n = h->n_local_total;
        if (slot < 0) { ensure_logged(h); apply_log = 0; }       // the resident assay after the plan's own log transform
    }
    if (m1 < 2 || n <= 0 || (!group && !covariate) || (group && !F_out) || (covariate && !r_out))
        throw Error(RUVIII_ERR_INVALID, "ruviii_gene_stats: bad arguments");
    if (!resident && ldX < n) throw Error(RUVIII_ERR_INVALID, "ruviii_gene_stats: row stride < n");
    // ---- pieces: groups in ascending id order, each cut into runs of at most piece_rows samples ----
    std::map<int32_t, std::vector<int>> groups;
    if (group) {
        for (int i = 0; i < m1; ++i)
            if (group[i] >= 0) groups[group[i]].push_back(i);
    } else {
        for (int i = 0; i < m1; ++i) groups[0].push_back(i);
    }
    int n_rows = 0;
    for (auto& kv : groups) n_rows += (int)kv.second.size();
    const int n_groups = (int)groups.size();
    if (group && (n_groups < 2 || n_rows <= n_groups)) throw Error(RUVIII_ERR_INVALID, "one-way ANOVA needs >= 2 groups and N > G");
    if (n_rows < 2) throw Error(RUVIII_ERR_INVALID, "ruviii_gene_stats: fewer than two samples");
    const int nsh = (int)h->shards.size();
    const int piece_rows = gene_stats_piece_rows(std::max(1, n / nsh), n_rows, h->n_sms);
    std::vector<int> pstart{0}, prow, pgroup;
    int gi = 0;
    for (auto& kv : groups) {
        const std::vector<int>& rows = kv.second;
        for (size_t a = 0; a < rows.size(); a += piece_rows) {
            const size_t b = std::min(rows.size(), a + piece_rows);
            prow.insert(prow.end(), rows.begin() + a, rows.begin() + b);
            pstart.push_back((int)prow.size());
            pgroup.push_back(gi);
        }
        ++gi;
    }
    const int n_pieces = (int)pgroup.size();
    // centred covariate over the samples that take part (Pearson: sum c = 0)
    std::vector<double> c(m1, 0.0);
    double cov_ss = 1.0;
    if (covariate) {
        long double mean = 0.0L;
        for (int r : prow) mean += covariate[r];
        mean /= n_rows;
        long double ss = 0.0L;
        for (int r : prow) { c[r] = (double)(covariate[r] - mean); ss += (long double)c[r] * c[r]; }
        cov_ss = (double)ss;
    }
    h->launches = 0;
    for (int i = 0; i < nsh; ++i) {
        Shard& s = h->shards[i];
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        const int g0 = regressed ? s.reg_g0 : resident ? s.g0 : (int)((int64_t)n * i / nsh);
        const int ng = regressed ? s.reg_n : resident ? s.n : (int)((int64_t)n * (i + 1) / nsh) - g0;
        if (ng <= 0) continue;
        const int64_t ld = regressed ? s.reg_ld : resident ? s.ld : round_up(ng, kRowAlign);
        s.gStart.alloc(pstart.size(), false);
        s.gRows.alloc(prow.size(), false);
        s.gGroup.alloc(pgroup.size(), false);
        s.gPart.alloc((size_t)n_pieces * 3 * ld, false);
        s.gF.alloc(ld, false);
        s.gR.alloc(ld, false);
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(s.gStart.p, pstart.data(), pstart.size() * sizeof(int), cudaMemcpyHostToDevice, s.stream));
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(s.gRows.p, prow.data(), prow.size() * sizeof(int), cudaMemcpyHostToDevice, s.stream));
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(s.gGroup.p, pgroup.data(), pgroup.size() * sizeof(int), cudaMemcpyHostToDevice, s.stream));
        if (covariate) {
            s.gCov.alloc(m1, false);
            RUVIII_CUDA_CHECK(cudaMemcpyAsync(s.gCov.p, c.data(), (size_t)m1 * 8, cudaMemcpyHostToDevice, s.stream));
        }
        const double* src;
        if (!resident) {
            s.gX.alloc((size_t)m1 * ld, false);
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(s.gX.p, ld * 8, X + g0, ldX * 8, (size_t)ng * 8, m1, cudaMemcpyHostToDevice, s.stream));
            src = s.gX.p;
        } else {
            src = regressed ? s.gReg.p : slot >= 0 ? s.out.p + (size_t)slot * m1 * s.ld : s.Y.p;
        }
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_START], s.stream));
        launch_gene_stats(src, ld, ng, s.gStart.p, s.gRows.p, s.gGroup.p, n_pieces, n_groups, n_rows, covariate ? s.gCov.p : nullptr,
                          cov_ss, apply_log, pseudo_count, s.gPart.p, ld, group ? s.gF.p : nullptr, covariate ? s.gR.p : nullptr,
                          s.stream);
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_LOG], s.stream));
        h->launches += 2;
        if (group) RUVIII_CUDA_CHECK(cudaMemcpyAsync(F_out + g0, s.gF.p, (size_t)ng * 8, cudaMemcpyDeviceToHost, s.stream));
        if (covariate) RUVIII_CUDA_CHECK(cudaMemcpyAsync(r_out + g0, s.gR.p, (size_t)ng * 8, cudaMemcpyDeviceToHost, s.stream));
    }
    sync_all(h);
    if (ms_out) {
        RUVIII_CUDA_CHECK(cudaSetDevice(h->shards[0].device));
        cudaEventElapsedTime(ms_out, h->shards[0].ev[E_START], h->shards[0].ev[E_LOG]);
    }
}

// Spearman's rho of every gene with a sample covariate (supervisedFindNCG.R:214-229 / :270-285 with corr.method = "spearman").
// The covariate's ranks (ties averaged) are O(m1 log m1) host bookkeeping; ranking the m1 values of every gene is the kernel.
void do_gene_spearman(ruviii_handle* h, const double* X, int64_t ldX, int m1, int n, int slot, const double* covariate, double* rho_out,
                      float* ms_out) {
    const bool resident = X == nullptr;
    const bool regressed = resident && slot == -2;
    if (regressed) {
        if (!h->reg_valid) throw Error(RUVIII_ERR_STATE, "ruviii_gene_spearman: slot -2 needs a ruviii_regress_out first");
        m1 = h->reg_m1;
        n = h->reg_n;
    } else if (resident) {
        if (!h->planned || !h->uploaded) throw Error(RUVIII_ERR_STATE, "ruviii_gene_spearman on resident data needs a plan and an upload");
        if (slot >= 0 && (!h->ran || slot >= (int)h->ks.size())) throw Error(RUVIII_ERR_STATE, "ruviii_gene_spearman: no such output block");
        m1 = h->m1;
        n = h->n_local_total;
        if (slot < 0) ensure_logged(h);
    }
    if (m1 < 2 || n <= 0 || !covariate || !rho_out) throw Error(RUVIII_ERR_INVALID, "ruviii_gene_spearman: bad arguments");
    if (!resident && ldX < n) throw Error(RUVIII_ERR_INVALID, "ruviii_gene_spearman: row stride < n");
    if (m1 > spearman_max_rows())
        throw Error(RUVIII_ERR_UNSUPPORTED, "ruviii_gene_spearman: more than " + std::to_string(spearman_max_rows()) +
                                                " samples do not fit the shared-memory sort");
    // average ranks of the covariate
    std::vector<int> order(m1);
    for (int i = 0; i < m1; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return covariate[a] < covariate[b]; });
    std::vector<double> cr(m1);
    for (int a = 0; a < m1;) {
        int b = a;
        while (b + 1 < m1 && covariate[order[b + 1]] == covariate[order[a]]) ++b;
        for (int t = a; t <= b; ++t) cr[order[t]] = 0.5 * (a + b) + 1.0;
        a = b + 1;
    }
    const double cmean = 0.5 * (m1 + 1.0);
    long double css = 0.0L;
    for (int i = 0; i < m1; ++i) css += (long double)(cr[i] - cmean) * (cr[i] - cmean);
    const int nsh = (int)h->shards.size();
    h->launches = 0;
    for (int i = 0; i < nsh; ++i) {
        Shard& s = h->shards[i];
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        const int g0 = regressed ? s.reg_g0 : resident ? s.g0 : (int)((int64_t)n * i / nsh);
        const int ng = regressed ? s.reg_n : resident ? s.n : (int)((int64_t)n * (i + 1) / nsh) - g0;
        if (ng <= 0) continue;
        const int64_t ld = regressed ? s.reg_ld : resident ? s.ld : round_up(ng, kRowAlign);
        s.gCov.alloc(m1, false);
        s.gR.alloc(ld, false);
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(s.gCov.p, cr.data(), (size_t)m1 * 8, cudaMemcpyHostToDevice, s.stream));
        const double* src;
        if (!resident) {
            s.gX.alloc((size_t)m1 * ld, false);
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(s.gX.p, ld * 8, X + g0, ldX * 8, (size_t)ng * 8, m1, cudaMemcpyHostToDevice, s.stream));
            src = s.gX.p;
        } else {
            src = regressed ? s.gReg.p : slot >= 0 ? s.out.p + (size_t)slot * m1 * s.ld : s.Y.p;
        }
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_START], s.stream));
        launch_gene_spearman(src, ld, m1, ng, s.gCov.p, cmean, (double)css, s.gR.p, s.stream);
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_LOG], s.stream));
        h->launches += 1;
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(rho_out + g0, s.gR.p, (size_t)ng * 8, cudaMemcpyDeviceToHost, s.stream));
    }
    sync_all(h);
    if (ms_out) {
        RUVIII_CUDA_CHECK(cudaSetDevice(h->shards[0].device));
        cudaEventElapsedTime(ms_out, h->shards[0].ev[E_START], h->shards[0].ev[E_LOG]);
    }
}

}  // namespace ruviii

using namespace ruviii;

#pragma GCC visibility push(default)
extern "C" {

int ruviii_regress_out(ruviii_handle* h, const double* X, int64_t ldX, int32_t m1, int32_t n, int32_t slot, int32_t apply_log,
                       double pseudo_count, const double* design, int32_t p, double* residuals_out, int64_t ld_out, int32_t* rank_out,
                       float* ms_out) {
    if (!h) return RUVIII_ERR_INVALID;
    return guarded(h, [&] {
        do_regress_out(h, X, ldX, m1, n, slot, apply_log, pseudo_count, design, p, residuals_out, ld_out, rank_out, ms_out);
    });
}

int ruviii_gene_spearman(ruviii_handle* h, const double* X, int64_t ldX, int32_t m1, int32_t n, int32_t slot, const double* covariate,
                         double* rho_out, float* ms_out) {
    if (!h) return RUVIII_ERR_INVALID;
    return guarded(h, [&] { do_gene_spearman(h, X, ldX, m1, n, slot, covariate, rho_out, ms_out); });
}

int ruviii_pca(ruviii_handle* h, const double* X, int64_t ldX, int32_t m1, int32_t n, int32_t slot, int32_t apply_log,
               double pseudo_count, int32_t center, int32_t scale, int32_t k, double* u_out, double* d_out, double* v_out,
               float* stage_ms6) {
    if (!h) return RUVIII_ERR_INVALID;
    return guarded(h, [&] { do_pca(h, X, ldX, m1, n, slot, apply_log, pseudo_count, center, scale, k, u_out, d_out, v_out, stage_ms6); });
}

int ruviii_gene_stats(ruviii_handle* h, const double* X, int64_t ldX, int32_t m1, int32_t n, int32_t slot, int32_t apply_log,
                      double pseudo_count, const int32_t* group_of_sample, const double* covariate, double* F_out, double* r_out,
                      float* ms_out) {
    if (!h) return RUVIII_ERR_INVALID;
    return guarded(h, [&] {
        do_gene_stats(h, X, ldX, m1, n, slot, apply_log, pseudo_count, group_of_sample, covariate, F_out, r_out, ms_out);
    });
}

int ruviii_set_prps_sets(ruviii_handle* h, const int32_t* set_start, const int32_t* set_rows, int32_t n_ps) {
    if (!h) return RUVIII_ERR_INVALID;
    return guarded(h, [&] {
        h->planned = false;
        if (!set_start || n_ps <= 0) {
            h->prps_start.clear();
            h->prps_rows.clear();
            h->prps_n = 0;
            return;
        }
        if (set_start[0] != 0) throw Error(RUVIII_ERR_INVALID, "ruviii_set_prps_sets: set_start[0] must be 0");
        for (int j = 0; j < n_ps; ++j)
            if (set_start[j + 1] <= set_start[j]) throw Error(RUVIII_ERR_INVALID, "ruviii_set_prps_sets: empty or unordered set");
        if (!set_rows) throw Error(RUVIII_ERR_INVALID, "ruviii_set_prps_sets: set_rows is NULL");
        for (int t = 0; t < set_start[n_ps]; ++t)
            if (set_rows[t] < 0) throw Error(RUVIII_ERR_INVALID, "ruviii_set_prps_sets: negative sample index");
        h->prps_start.assign(set_start, set_start + n_ps + 1);
        h->prps_rows.assign(set_rows, set_rows + set_start[n_ps]);
        h->prps_n = n_ps;
    });
}

int ruviii_prps(ruviii_handle* h, const double* Y, int64_t ldY, int32_t m1, int32_t n, int32_t apply_log, double pseudo_count,
                const int32_t* set_start, const int32_t* set_rows, int32_t n_ps, double* PS_out, int64_t ldPS) {
    if (!h) return RUVIII_ERR_INVALID;
    return guarded(h, [&] {
        if (!Y || !set_start || !set_rows || !PS_out || m1 <= 0 || n <= 0 || n_ps <= 0 || ldY < n || ldPS < n)
            throw Error(RUVIII_ERR_INVALID, "ruviii_prps: bad arguments");
        // only the member rows are needed on the device: upload each distinct one once, in ascending order
        const int total = set_start[n_ps];
        std::vector<int> uniq(set_rows, set_rows + total);
        for (int r : uniq)
            if (r < 0 || r >= m1) throw Error(RUVIII_ERR_INVALID, "ruviii_prps: sample index out of range");
        for (int j = 0; j < n_ps; ++j)
            if (set_start[j + 1] <= set_start[j]) throw Error(RUVIII_ERR_INVALID, "ruviii_prps: empty or unordered set");
        std::sort(uniq.begin(), uniq.end());
        uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
        std::vector<int> pos(m1, -1), rows(total);
        for (size_t i = 0; i < uniq.size(); ++i) pos[uniq[i]] = (int)i;
        for (int t = 0; t < total; ++t) rows[t] = pos[set_rows[t]];
        const int nu = (int)uniq.size(), nsh = (int)h->shards.size();
        std::vector<DevBuf<double>> dY(nsh), dP(nsh);
        std::vector<DevBuf<int>> dS(nsh), dR(nsh);
        for (int i = 0; i < nsh; ++i) {
            Shard& s = h->shards[i];
            RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
            const int g0 = (int)((int64_t)n * i / nsh), ng = (int)((int64_t)n * (i + 1) / nsh) - g0;
            if (ng <= 0) continue;
            const int64_t ld = round_up(ng, kRowAlign);
            dY[i].alloc((size_t)nu * ld);
            dP[i].alloc((size_t)n_ps * ld, false);
            dS[i].alloc(n_ps + 1);
            dR[i].alloc(total);
            RUVIII_CUDA_CHECK(cudaMemcpyAsync(dS[i].p, set_start, (size_t)(n_ps + 1) * sizeof(int), cudaMemcpyHostToDevice, s.stream));
            RUVIII_CUDA_CHECK(cudaMemcpyAsync(dR[i].p, rows.data(), (size_t)total * sizeof(int), cudaMemcpyHostToDevice, s.stream));
            for (int a = 0; a < nu;) {                 // runs of consecutive samples go up as one 2-D copy
                int b = a + 1;
                while (b < nu && uniq[b] == uniq[b - 1] + 1) ++b;
                RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(dY[i].p + (size_t)a * ld, ld * 8, Y + (size_t)uniq[a] * ldY + g0, ldY * 8,
                                                    (size_t)ng * 8, b - a, cudaMemcpyHostToDevice, s.stream));
                a = b;
            }
            if (apply_log) launch_log2_inplace(dY[i].p, ld, nu, ng, pseudo_count, s.stream);   // prpsForCategoricalUV.R:91
            RowSrc src{dY[i].p, ld, nullptr, 0, nu};
            launch_group_mean(src, dS[i].p, dR[i].p, n_ps, ng, dP[i].p, ld, s.stream);
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(PS_out + g0, ldPS * 8, dP[i].p, ld * 8, (size_t)ng * 8, n_ps, cudaMemcpyDeviceToHost,
                                                s.stream));
        }
        sync_all(h);
        h->launches = nsh * (apply_log ? 2 : 1);
        for (int i = 0; i < nsh; ++i) {
            cudaSetDevice(h->shards[i].device);
            dY[i].release(); dP[i].release(); dS[i].release(); dR[i].release();
        }
    });
}

int ruviii_set_eta(ruviii_handle* h, const double* eta, int32_t q, int32_t n) {
    if (!h) return RUVIII_ERR_INVALID;
    return guarded(h, [&] {
        if (!eta || q <= 0) {
            h->eta_host.clear();
            h->eta_q = h->eta_n = 0;
        } else {
            if (n <= 0) throw Error(RUVIII_ERR_INVALID, "ruviii_set_eta: n <= 0");
            if (q > kMaxK) throw Error(RUVIII_ERR_UNSUPPORTED, "eta with more than 32 rows is not supported");
            h->eta_host.assign(eta, eta + (size_t)q * n);
            h->eta_q = q;
            h->eta_n = n;
        }
        h->planned = false;                 // the plan sizes the RUV1 buffers
    });
}

int ruviii_fast_residop(ruviii_handle* h, const double* A, int32_t rows, int32_t n, const int32_t* group_of_row,
                        double* out) {
    if (!h) return RUVIII_ERR_INVALID;
    return guarded(h, [&] {
        if (!A || !out || !group_of_row || rows <= 0 || n <= 0) throw Error(RUVIII_ERR_INVALID, "fast_residop: bad arguments");
        Shard& s = h->shards[0];
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        std::map<int32_t, std::vector<int>> groups;
        for (int i = 0; i < rows; ++i) groups[group_of_row[i]].push_back(i);
        std::vector<int> gs{0}, gr;
        for (auto& kv : groups) {           // singletons included here: their residual rows are written as zeros
            for (int r : kv.second) gr.push_back(r);
            gs.push_back((int)gr.size());
        }
        const int64_t ld = round_up(n, kRowAlign);
        DevBuf<double> dA, dO;
        DevBuf<int> dgs, dgr;
        dA.alloc((size_t)rows * ld);
        dO.alloc((size_t)rows * ld);
        dgs.alloc(gs.size());
        dgr.alloc(gr.size());
        RUVIII_CUDA_CHECK(cudaMemcpy(dgs.p, gs.data(), gs.size() * 4, cudaMemcpyHostToDevice));
        RUVIII_CUDA_CHECK(cudaMemcpy(dgr.p, gr.data(), gr.size() * 4, cudaMemcpyHostToDevice));
        RUVIII_CUDA_CHECK(cudaMemcpy2D(dA.p, ld * 8, A, (size_t)n * 8, (size_t)n * 8, rows, cudaMemcpyHostToDevice));
        RowSrc src{dA.p, ld, dA.p, ld, rows};
        launch_residop(src, dgs.p, dgr.p, (int)gs.size() - 1, 0, n, dO.p, ld, s.stream);
        RUVIII_CUDA_CHECK(cudaStreamSynchronize(s.stream));
        // compact row t of the kernel output is the residual of input row gr[t]
        std::vector<double> tmp((size_t)rows * n);
        RUVIII_CUDA_CHECK(cudaMemcpy2D(tmp.data(), (size_t)n * 8, dO.p, ld * 8, (size_t)n * 8, rows, cudaMemcpyDeviceToHost));
        for (int t = 0; t < rows; ++t) std::memcpy(out + (size_t)gr[t] * n, tmp.data() + (size_t)t * n, (size_t)n * 8);
        dA.release(); dO.release(); dgs.release(); dgr.release();
    });
}

int ruviii_probe(ruviii_handle* h, int32_t which, int32_t size, double* value_out) {
    if (!h || !value_out) return RUVIII_ERR_INVALID;
    return guarded(h, [&] {
        Shard& s = h->shards[0];
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        if (which == 0) {
            *value_out = probe_copy_gbs((size_t)std::max(size, 64) << 20, s.stream);
        } else if (which == 2) {
            *value_out = probe_dmma_tflops(h->n_sms, s.stream);
        } else if (which == 5) {
            // median microseconds of an in-place FP64 sum-allreduce of `size` KiB over all shards (every rank must call)
            const size_t count = (size_t)std::max(size, 1) * 128;
            std::vector<DevBuf<double>> bufs(h->shards.size());
            std::vector<double*> ptrs;
            for (size_t i = 0; i < h->shards.size(); ++i) {
                RUVIII_CUDA_CHECK(cudaSetDevice(h->shards[i].device));
                bufs[i].alloc(count);
                ptrs.push_back(bufs[i].p);
            }
            RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            std::vector<float> t;
            for (int it = 0; it < 25; ++it) {
                sync_all(h);
                RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
                cudaEventRecord(e0, s.stream);
                allreduce_sum(h, ptrs, count);
                RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
                cudaEventRecord(e1, s.stream);
                cudaEventSynchronize(e1);
                float ms = 0.f;
                cudaEventElapsedTime(&ms, e0, e1);
                if (it >= 5) t.push_back(ms);
            }
            sync_all(h);
            cudaEventDestroy(e0); cudaEventDestroy(e1);
            for (size_t i = 0; i < h->shards.size(); ++i) { cudaSetDevice(h->shards[i].device); bufs[i].release(); }
            std::sort(t.begin(), t.end());
            *value_out = 1e3 * t[t.size() / 2];
        } else if (which == 6) {
            // aggregate GB/s of `size` MiB host->device and `size` MiB device->host copied AT THE SAME TIME (pinned memory, two
            // streams): what the overlapped part of ruviii_normalise can expect from the full-duplex link
            const size_t bytes = (size_t)std::max(size, 16) << 20;
            void *hin = nullptr, *hout = nullptr;
            RUVIII_CUDA_CHECK(cudaHostAlloc(&hin, bytes, cudaHostAllocDefault));
            RUVIII_CUDA_CHECK(cudaHostAlloc(&hout, bytes, cudaHostAllocDefault));
            std::memset(hin, 1, bytes);
            std::memset(hout, 0, bytes);
            DevBuf<double> din, dout;
            din.alloc(bytes / 8, false);
            dout.alloc(bytes / 8);
            double best = 1e30;
            for (int it = 0; it < 3; ++it) {
                sync_all(h);
                const auto t0 = std::chrono::steady_clock::now();
                RUVIII_CUDA_CHECK(cudaMemcpyAsync(din.p, hin, bytes, cudaMemcpyHostToDevice, s.copy_stream));
                RUVIII_CUDA_CHECK(cudaMemcpyAsync(hout, dout.p, bytes, cudaMemcpyDeviceToHost, s.d2h_stream));
                sync_all(h);
                best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
            }
            din.release(); dout.release();
            cudaFreeHost(hin); cudaFreeHost(hout);
            *value_out = 2.0 * (double)bytes / best / 1e9;
        } else if (which == 4) {
            double r[8];
            probe_copy_variants(r, s.stream);
            *value_out = r[std::min(std::max(size, 0), 7)];
        } else if (which == 3) {
            *value_out = probe_host_gather_ms(11000, 20000, std::max(size, 1), s.stream);
        } else if (which == 1) {
            if (!h->blas && cublasCreate(&h->blas) != CUBLAS_STATUS_SUCCESS) throw Error(RUVIII_ERR_CUDA, "cublasCreate failed");
            cublasSetStream(h->blas, s.stream);
            const int N = std::max(size, 256);
            DevBuf<double> a, b, c;
            a.alloc((size_t)N * N); b.alloc((size_t)N * N); c.alloc((size_t)N * N);
            const double one = 1.0, zero = 0.0;
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            float best = 1e30f;
            for (int it = 0; it < 4; ++it) {
                cudaEventRecord(e0, s.stream);
                cublasDgemm(h->blas, CUBLAS_OP_T, CUBLAS_OP_N, N, N, N, &one, a.p, N, b.p, N, &zero, c.p, N);
                cudaEventRecord(e1, s.stream);
                cudaEventSynchronize(e1);
                float ms = 0.f;
                cudaEventElapsedTime(&ms, e0, e1);
                if (it >= 1) best = std::min(best, ms);
            }
            cudaEventDestroy(e0); cudaEventDestroy(e1);
            a.release(); b.release(); c.release();
            *value_out = 2.0 * N * (double)N * N / (best * 1e-3) / 1e12;
        } else {
            throw Error(RUVIII_ERR_INVALID, "unknown probe");
        }
    });
}

}  // extern "C"
#pragma GCC visibility pop

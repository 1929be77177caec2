// The overlapped host-to-host path of ruviii_normalise.
#include "engine_internal.cuh"

namespace ruviii {

// ---- pipelined host-to-host path of ruviii_normalise -----------------------------------------------------------
// The one-shot call is PCIe-bound (C3: 1.95 GB in, 1.76 GB out, ~35 ms each way at ~55 GB/s against ~2 ms of
// kernels), and the link is full duplex, so the only real lever is to start the download before the upload has
// finished.  W needs the control columns of EVERY row of Y, which would serialise the two directions; so those
// columns (m1 x c doubles, 5 % of Y) are gathered straight out of the pinned host buffer by a zero-copy kernel while
// the bulk DMA is still streaming Y in row chunks.  Timeline per shard:
//   copy stream : PS -> Y chunk 0 -> Y chunk 1 -> ...                              (H2D DMA)
//   d2h stream  : zero-copy gather of Y[, ctl]  ....  newY chunk 0 -> newY chunk 1 -> ...   (D2H DMA)
//   main stream : K1 K2 [allreduce] K3 [bcast] K4 | wait gather | K5 [allreduce] solve | per chunk: wait upload, K6
// Taken when the host buffers are pinned, every replicate row is a PRPS row (the PRPS regime: residop/Gram/eigen
// need PS only), the sample-side Gram is used and nothing but newY / W / the leading rows of fullalpha is wanted.
// Device-visible alias of a pinned (cudaHostAlloc'ed / registered) host pointer, or nullptr for pageable memory.
const void* host_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

bool can_pipeline(ruviii_handle* h, const double* Y, const double* PS, const double* newY_out, const double* newY_ps_out) {
    if (h->n_ps > 0 && !PS) return false;      // replicate.data is built on the device from the whole assay
    // selected by RUVIII_PIPELINE=3 (engine.cu normalise_any); eligibility only here
    const bool off = false;
    const bool skip = h->no_replicates || h->kk == 0;
    return !off && !skip && h->eta_q == 0 && !h->gene_side && h->active_all_ps && !h->prm.average_rep && h->n_alpha_rows == h->kk &&
           h->m1 >= 256 && newY_ps_out == nullptr && host_pinned(Y) != nullptr && host_pinned(newY_out) != nullptr;
}

void do_normalise_pipelined(ruviii_handle* h, const double* Y, int64_t ldY, const double* PS, int64_t ldPS, double* newY_out,
                            double* fullalpha_out, double* W_out) {
    const bool fast = h->prm.mode == RUVIII_MODE_FAST;
    const int kp = h->kp, kk = h->kk, R = h->R, nk = (int)h->ks.size(), m1 = h->m1;
    const int64_t n = h->n_local_total;
    const double t_enter = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    h->launches = 0;
    auto for_shards = [&](auto&& fn) {
        for (auto& s : h->shards) {
            RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
            fn(s);
        }
    };
    // row chunks of ~1/12 of the assay (>= 256 rows, multiple of 64)
    const int chunk = (int)round_up(std::max(256, ceil_div(m1, 12)), 64);
    const int n_chunks = ceil_div(m1, chunk);
    // Share of the row chunks (the LAST ones in upload order) whose control columns come by zero-copy.  The scattered reads
    // cost ~1.7 ns each however many GPUs issue them (a request-rate limit on the host side of PCIe: 18-19 ms for the
    // 11 M reads of C3 on one GPU, and the same total on eight), the bulk upload of a shard runs at ~55 GB/s; W is known
    // when both the zero-copy share and the upload of the leading chunks are done, so the share that balances the two is
    // (t_ps + t_upload) / (t_zero_copy_all + t_upload); a factor 1.2 leans towards zero-copy because it starts at t = 0
    // while the leading chunks also wait for the PRPS block (measured optimum 75-100 % at N = 1).  RUVIII_ZC_PCT overrides.
    int zc_pct = env_int("RUVIII_ZC_PCT", -1);
    if (zc_pct < 0) {
        const double t_zc = (double)m1 * (double)h->n_ctl_global * 1.7e-9;
        const double t_up = (double)m1 * (double)h->shards[0].n * 8.0 / 55e9, t_ps = (double)h->n_ps * (double)h->shards[0].n * 8.0 / 55e9;
        zc_pct = (int)(100.0 * std::min(1.0, 1.2 * (t_ps + t_up) / (t_zc + t_up)) + 0.5);
    }
    zc_pct = std::max(0, std::min(100, zc_pct));
    const int n_zc = std::min(n_chunks, (n_chunks * zc_pct + 50) / 100);
    const int lead = n_chunks - n_zc;
    const int row_zc0 = std::min(m1, lead * chunk);
    // RUVIII_PIPE_TRACE=1: timestamps (ms after entry, shard 0) of the milestones of the overlapped pipeline on stderr
    static const bool trace = env_int("RUVIII_PIPE_TRACE", 0) != 0;
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    auto mark = [&](const char* what, Shard& s, cudaStream_t st) {
        if (!trace || &s != &h->shards[0]) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        marks.emplace_back(what, e);
    };
    for_shards([&](Shard& s) {
        while ((int)s.ev_chunk.size() < n_chunks + 2) {
            cudaEvent_t e;
            RUVIII_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            s.ev_chunk.push_back(e);
            RUVIII_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            s.ev_upd.push_back(e);
        }
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_START], s.stream));
        // --- uploads: PS first, then Y in row chunks (copy stream) ---
        if (h->n_ps > 0)
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(s.PS.p, s.ld * 8, PS + s.g0, ldPS * 8, (size_t)s.n * 8, h->n_ps,
                                                cudaMemcpyHostToDevice, s.copy_stream));
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_chunk[n_chunks], s.copy_stream));          // PS resident
        mark("ps_uploaded", s, s.copy_stream);
        // control columns of the PRPS rows (STACKED mode): from the device copy, right behind its upload
        if (h->rows_w > m1 && s.n_ctl > 0) {
            RowSrc ps_src{nullptr, 0, s.PS.p, s.ld, 0};
            launch_ctl_gather(ps_src, h->rows_w - m1, s.ctl_idx.p, s.n_ctl, s.Yc.p + (size_t)m1 * s.ldc, s.ldc, 0, 0.0, s.copy_stream);
        }
        if (lead == 0) RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_fork, s.copy_stream));   // nothing but the PRPS rows is gathered here
        for (int c = 0; c < n_chunks; ++c) {
            const int r0 = c * chunk, rows = std::min(chunk, m1 - r0);
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(s.Y.p + (size_t)r0 * s.ld, s.ld * 8, Y + (size_t)r0 * ldY + s.g0, ldY * 8,
                                                (size_t)s.n * 8, rows, cudaMemcpyHostToDevice, s.copy_stream));
            RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_chunk[c], s.copy_stream));
            if (c < lead && s.n_ctl > 0) {
                // leading chunks: their control columns are gathered on the device as soon as the chunk has landed
                RowSrc dsrc{s.Y.p + (size_t)r0 * s.ld, s.ld, nullptr, 0, rows};
                launch_ctl_gather(dsrc, rows, s.ctl_idx.p, s.n_ctl, s.Yc.p + (size_t)r0 * s.ldc, s.ldc, h->prm.apply_log ? rows : 0,
                                  h->prm.pseudo_count, s.copy_stream);
            }
            if (c == lead - 1) {
                RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_fork, s.copy_stream));   // device-side gathers complete
                mark("lead_chunks_uploaded+gathered", s, s.copy_stream);
            }
        }
        mark("all_uploaded", s, s.copy_stream);
        // --- trailing chunks: zero-copy gather of their control columns out of the pinned host assay (d2h stream, idle
        //     until K6).  It starts at once; the split between the two routes balances its ~19 ms per 11 M scattered reads
        //     (bound by the host side of PCIe) against the upload time of the leading chunks. ---
        if (row_zc0 < m1) {
            RowSrc hsrc{static_cast<const double*>(host_pinned(Y)) + (size_t)row_zc0 * ldY + s.g0, ldY, nullptr, 0, m1 - row_zc0};
            launch_ctl_gather(hsrc, m1 - row_zc0, s.ctl_idx.p, s.n_ctl, s.Yc.p + (size_t)row_zc0 * s.ldc, s.ldc,
                              h->prm.apply_log ? m1 - row_zc0 : 0, h->prm.pseudo_count, s.d2h_stream);
        }
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_gather, s.d2h_stream));
        mark("zero_copy_gather_done", s, s.d2h_stream);
        // --- K1, K2 on the PRPS rows ---
        RUVIII_CUDA_CHECK(cudaStreamWaitEvent(s.stream, s.ev_chunk[n_chunks], 0));
        RowSrc src{s.Y.p, s.ld, s.PS.p, s.ld, m1};
        launch_helmert(src, s.grp_start.p, s.grp_rows.p, h->n_groups_active, h->max_group, s.n, s.Y0.p, s.ld, s.stream);
        launch_gram(s.gp, s.Y0.p, s.ld, s.tile_map.p, s.gram_ws.p, s.G.p, s.tmap, s.stream);
    });
    h->launches += 4;
    {
        std::vector<double*> b;
        for (auto& s : h->shards) b.push_back(s.G.p);
        allreduce_sum(h, b, (size_t)R * R);
    }
    if (h->rank0 == 0) eigensolve(h);
    {
        std::vector<double*> b;
        for (auto& s : h->shards) b.push_back(s.U.p);
        broadcast0(h, b, (size_t)R * kp);
        b.clear();
        for (auto& s : h->shards) b.push_back(s.evals.p);
        broadcast0(h, b, (size_t)std::max(h->n_keep, kp) + 2);
    }
    for_shards([&](Shard& s) {
        launch_project(s.Y0.p, s.ld, R, s.n, s.U.p, kp, s.rsplit, s.part.p, s.ld, s.stream, kk);
        launch_alpha_finalize(s.part.p, s.rsplit, kp, s.n, s.ld, s.alpha.p, s.ctl_idx.p, s.n_ctl, s.acT.p, s.stream);
        RowSrc src{s.Y.p, s.ld, s.PS.p, s.ld, m1};
        RUVIII_CUDA_CHECK(cudaStreamWaitEvent(s.stream, s.ev_gather, 0));
        RUVIII_CUDA_CHECK(cudaStreamWaitEvent(s.stream, s.ev_fork, 0));
        launch_ctl_products(src, h->rows_w, s.ctl_idx.p, s.n_ctl, s.acT.p, kp, s.AG.p, s.AG.p + (size_t)h->rows_w * kp,
                            s.n_ctl > 0 ? s.Yc.p : nullptr, s.ldc, s.stream);
    });
    h->launches += 5 + lead;
    {
        std::vector<double*> b;
        for (auto& s : h->shards) b.push_back(s.AG.p);
        allreduce_sum(h, b, (size_t)h->rows_w * kp + (size_t)kp * kp);
    }
    for_shards([&](Shard& s) {
        launch_solve(h, s);
        mark("W_known", s, s.stream);
        const int64_t stride = (int64_t)m1 * s.ld;
        // --- K6 per chunk as it lands, download right behind it ---
        for (int c = 0; c < n_chunks; ++c) {
            const int r0 = c * chunk, rows = std::min(chunk, m1 - r0);
            RUVIII_CUDA_CHECK(cudaStreamWaitEvent(s.stream, s.ev_chunk[c], 0));
            if (h->prm.apply_log) launch_log2_inplace(s.Y.p + (size_t)r0 * s.ld, s.ld, rows, s.n, h->prm.pseudo_count, s.stream);
            launch_update_smem(s.Y.p + (size_t)r0 * s.ld, s.ld, rows, s.n, s.At.p + (size_t)r0 * kp, kp, s.at.p, s.ld, kk,
                               h->kmask, h->slot_of_k, s.out.p + (size_t)r0 * s.ld, s.ld, stride, s.stream);
            for (auto& d : h->dup_slots)
                RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(s.out.p + (size_t)d.first * stride + (size_t)r0 * s.ld, s.ld * 8,
                                                    s.out.p + (size_t)d.second * stride + (size_t)r0 * s.ld, s.ld * 8,
                                                    (size_t)s.n * 8, rows, cudaMemcpyDeviceToDevice, s.stream));
            RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_upd[c], s.stream));
            RUVIII_CUDA_CHECK(cudaStreamWaitEvent(s.d2h_stream, s.ev_upd[c], 0));
            for (int i = 0; i < nk; ++i)
                RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(newY_out + (size_t)i * m1 * n + (size_t)r0 * n + s.g0, n * 8,
                                                    s.out.p + (size_t)i * stride + (size_t)r0 * s.ld, s.ld * 8, (size_t)s.n * 8,
                                                    rows, cudaMemcpyDeviceToHost, s.d2h_stream));
            if (c == 0) mark("first_chunk_downloaded", s, s.d2h_stream);
            if (c == n_chunks / 2) mark("half_downloaded", s, s.d2h_stream);
        }
        mark("all_downloaded", s, s.d2h_stream);
        if (W_out && &s == &h->shards[0]) launch_w_blocks(h, s);
        if (fullalpha_out)
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(fullalpha_out + s.g0, n * 8, s.alpha.p, s.ld * 8, (size_t)s.n * 8, h->n_alpha_rows,
                                                cudaMemcpyDeviceToHost, s.stream));
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_UPDATE], s.stream));
    });
    h->launches += n_chunks * (h->prm.apply_log ? 2 : 1);
    if (W_out) {
        Shard& s = h->shards[0];
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        size_t w_total = 0;
        for (int i = 0; i < nk; ++i) w_total += (size_t)h->rows_w * h->k_eff[i];
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(W_out, s.W.p, w_total * 8, cudaMemcpyDeviceToHost, s.stream));
    }
    for (size_t i = 0; i < h->shards.size(); ++i) {
        Shard& s = h->shards[i];
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(&h->h_flags[1 + i], s.status.p, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    }
    sync_all(h);
    const double t_done = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    if (trace) {
        fprintf(stderr, "[pipe trace] zc_pct=%d chunks=%d lead=%d:", zc_pct, n_chunks, lead);
        for (auto& mk : marks) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, h->shards[0].ev[E_START], mk.second);
            fprintf(stderr, " %s=%.2f", mk.first, ms);
            cudaEventDestroy(mk.second);
        }
        fprintf(stderr, " host_total=%.2f (ms)\n", (t_done - t_enter) * 1e3);
    }
    if (h->rank0 == 0 && h->h_flags[0] != 0)
        throw Error(RUVIII_ERR_SOLVER, "cuSOLVER syevd devInfo = " + std::to_string(h->h_flags[0]));
    for (size_t i = 0; i < h->shards.size(); ++i)
        if (h->h_flags[1 + i] != 0)
            throw Error(RUVIII_ERR_SINGULAR, "ac %*% t(ac) is not positive definite at pivot " + std::to_string(h->h_flags[1 + i]) +
                                                 " (solve() in ruvIII.R:144 would fail)");
    ruviii_info& I = h->info;
    std::memset(&I, 0, sizeof(I));
    I.struct_size = sizeof(ruviii_info);
    I.m = h->m; I.m1 = m1; I.n_ps = h->n_ps; I.n = (int)n; I.p = h->p; I.active_rows = h->active_rows; I.r = h->r;
    I.gram_order = h->gorder; I.n_ctl = (int)h->n_ctl_global; I.kmax_eff = kk; I.gram_side = 1;
    I.n_shards = (int)h->shards.size(); I.kernel_launches = h->launches;
    I.eig_solver = h->eig_solver_used; I.eig_iterations = h->eig_iterations; I.pipelined = 1;
    I.host_t_enter = t_enter; I.host_t_done = t_done;
    I.ms_total = (float)((t_done - t_enter) * 1e3);          // host wall time of the overlapped copy + compute pipeline
    {
        Shard& s = h->shards[0];
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        std::vector<double> ev(kk + 1, 0.0);
        const int have = std::min(kk + 1, h->n_keep);
        RUVIII_CUDA_CHECK(cudaMemcpy(ev.data(), s.evals.p, have * sizeof(double), cudaMemcpyDeviceToHost));
        I.eig_top = ev[0]; I.eig_kmax = ev[kk - 1]; I.eig_next = have > kk ? ev[kk] : 0.0;
    }
    (void)fast;
    h->uploaded = true;
    h->logged = h->prm.apply_log != 0;
    h->yc_valid = true;                  // the whole control block was gathered on the way in
    h->ran = true;
    h->pipelined_last = 1;
}

}  // namespace ruviii

// Staged host-to-host path of ruviii_normalise: works from PAGEABLE host memory (what R hands over: REALSXP buffers live
// on R's heap and cannot be page-locked per call at a sensible cost) as well as from pinned buffers.
//
// The one-shot call is PCIe-bound (C3: 1.95 GB in, 1.76 GB out around ~1.5 ms of kernels), so the path is organised
// around keeping both DMA directions busy at once:
//   phase A  the control columns Y[, ctl] (5 % of Y; W needs them from EVERY row, ruvIII.R:118-120,144) are copied out of
//            the caller's matrix by a pool of host threads into one pinned block and uploaded with the PRPS rows; the device
//            runs K1..K5 on them, so W is known after a few milliseconds instead of after the whole upload;
//   phase B  the assay streams through in row chunks: [host threads: caller -> pinned slot] -> H2D -> K6 on the chunk ->
//            D2H -> [host threads: pinned slot -> caller's output], three slots per direction, upload of chunk c + 1,
//            update of chunk c and download of chunk c - 1 in flight together.
// With pinned caller buffers the host copies drop out and the DMA engines read / write the caller's memory directly.
// No arithmetic happens on the host: the pool only moves bytes (the log transform of the gathered control block runs on
// the device).  One shard per process (the R session with one GPU, or one rank of a torchrun launch).
#include "engine_internal.cuh"

#include <condition_variable>
#include <functional>
#include <sched.h>
#include <thread>

namespace ruviii {

// ---- a small persistent pool of host threads that executes index ranges ------------------------------------------------
struct HostPool {
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable cv, cv_done;
    std::function<void(int)> job;
    int n_items = 0, next = 0, pending = 0;
    uint64_t gen = 0;
    bool stop = false;

    explicit HostPool(int n) {
        for (int i = 0; i < n; ++i) threads.emplace_back([this] { worker(); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(m);
            stop = true;
        }
        cv.notify_all();
        for (auto& t : threads) t.join();
    }
    void worker() {
        uint64_t seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(m);
            cv.wait(lk, [&] { return stop || (gen != seen && next < n_items); });
            if (stop) return;
            seen = gen;
            while (next < n_items) {
                const int i = next++;
                lk.unlock();
                job(i);
                lk.lock();
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
    // runs fn(0) .. fn(n - 1) on the pool and the calling thread; returns when all are done
    void run(int n, std::function<void(int)> fn) {
        if (n <= 0) return;
        std::unique_lock<std::mutex> lk(m);
        job = std::move(fn);
        n_items = n;
        next = 0;
        pending = n;
        ++gen;
        cv.notify_all();
        while (next < n_items) {
            const int i = next++;
            lk.unlock();
            job(i);
            lk.lock();
            --pending;
        }
        cv_done.wait(lk, [&] { return pending == 0; });
    }
};

struct StageState {
    std::unique_ptr<HostPool> pool;
    int n_threads = 0;
    static constexpr int kSlots = 3;
    double* in_slot[kSlots] = {nullptr, nullptr, nullptr};
    double* out_slot[kSlots] = {nullptr, nullptr, nullptr};
    size_t in_bytes = 0, out_bytes = 0;
    double* yc_host = nullptr;
    size_t yc_bytes = 0;
    ~StageState() {
        for (auto p : in_slot) if (p) cudaFreeHost(p);
        for (auto p : out_slot) if (p) cudaFreeHost(p);
        if (yc_host) cudaFreeHost(yc_host);
    }
};

void stage_state_destroy(StageState* st) { delete st; }

namespace {

int host_threads_default() {
    int avail = (int)std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) avail = std::min(avail > 0 ? avail : 1 << 30, CPU_COUNT(&set));
    if (avail <= 0) avail = 4;
    return std::max(1, std::min(16, avail));
}

void ensure_pinned(double*& p, size_t& have, size_t need) {
    if (need <= have && p) return;
    if (p) cudaFreeHost(p);
    p = nullptr;
    have = 0;
    RUVIII_CUDA_CHECK(cudaHostAlloc((void**)&p, need, cudaHostAllocDefault));
    have = need;
}

// copies `rows` rows of `row_bytes` bytes: src pitch -> dst pitch, rows split over the pool
void pool_copy_rows(HostPool& pool, int n_threads, char* dst, size_t dst_pitch, const char* src, size_t src_pitch, size_t row_bytes,
                    int64_t rows) {
    if (rows <= 0 || row_bytes == 0) return;
    if (dst_pitch == row_bytes && src_pitch == row_bytes) {          // one contiguous block: split by bytes (64-byte aligned cuts)
        const size_t total = (size_t)rows * row_bytes;
        const int parts = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_threads * 2, total >> 20));
        pool.run(parts, [=](int i) {
            const size_t a = (total * (size_t)i / parts) & ~size_t(63), b = i + 1 == parts ? total : (total * (size_t)(i + 1) / parts) & ~size_t(63);
            std::memcpy(dst + a, src + a, b - a);
        });
        return;
    }
    const int parts = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)n_threads * 2, rows));
    pool.run(parts, [=](int i) {
        const int64_t r0 = rows * i / parts, r1 = rows * (i + 1) / parts;
        for (int64_t r = r0; r < r1; ++r) std::memcpy(dst + (size_t)r * dst_pitch, src + (size_t)r * src_pitch, row_bytes);
    });
}

}  // namespace

const void* host_pinned(const void* p);   // engine_pipeline.cu

bool can_stage(ruviii_handle* h, const double* PS, const double* newY_ps_out) {
    if (h->n_ps > 0 && !PS) return false;      // replicate.data is built on the device from the whole assay
    const bool skip = h->no_replicates || h->kk == 0;
    return h->shards.size() == 1 && !skip && h->eta_q == 0 && !h->gene_side && h->active_all_ps && !h->prm.average_rep &&
           h->n_alpha_rows == h->kk && h->m1 >= 256 && newY_ps_out == nullptr;
}

void do_normalise_staged(ruviii_handle* h, const double* Y, int64_t ldY, const double* PS, int64_t ldPS, double* const* out_blocks,
                         double* fullalpha_out, double* W_out) {
    Shard& s = h->shards[0];
    RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
    const int kp = h->kp, kk = h->kk, R = h->R, nk = (int)h->ks.size(), m1 = h->m1, n_ps = h->n_ps;
    const int64_t n = s.n;
    const double t_enter = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    h->launches = 0;
    if (!h->stage) h->stage = new StageState();
    StageState& st = *h->stage;
    const int want_threads = std::max(1, env_int("RUVIII_HOST_THREADS", host_threads_default()));
    if (!st.pool || st.n_threads != want_threads) {
        st.pool.reset(new HostPool(want_threads - 1));     // the calling thread works too
        st.n_threads = want_threads;
    }
    HostPool& pool = *st.pool;
    static const bool trace = env_int("RUVIII_PIPE_TRACE", 0) != 0;
    auto now_ms = [&] { return 1e3 * (std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count() - t_enter); };
    double t_gather = 0, t_ps = 0, t_w = 0;

    const bool in_pinned = host_pinned(Y) != nullptr;
    bool out_pinned = true;
    for (int i = 0; i < nk; ++i) out_pinned = out_pinned && host_pinned(out_blocks[i]) != nullptr;
    const bool ps_pinned = n_ps == 0 || host_pinned(PS) != nullptr;

    // ---- row chunks: in-slot <= ~48 MB, out-slot (nk blocks) <= ~96 MB ----
    const size_t row_bytes = (size_t)n * 8;
    int64_t rows = std::min<int64_t>((int64_t)(48u << 20) / (int64_t)row_bytes, (int64_t)(96u << 20) / (int64_t)(row_bytes * nk));
    rows = std::max<int64_t>(64, rows / 64 * 64);
    const int chunk = (int)std::min<int64_t>(rows, round_up(m1, 64));
    const int n_chunks = ceil_div(m1, chunk);
    constexpr int NS = StageState::kSlots;
    auto ensure_slots = [&](double** slots, size_t& have, size_t need) {
        if (need <= have && slots[0]) return;
        for (int i = 0; i < NS; ++i) {
            if (slots[i]) cudaFreeHost(slots[i]);
            slots[i] = nullptr;
        }
        have = 0;
        for (int i = 0; i < NS; ++i) RUVIII_CUDA_CHECK(cudaHostAlloc((void**)&slots[i], need, cudaHostAllocDefault));
        have = need;
    };
    if (!in_pinned || !ps_pinned) ensure_slots(st.in_slot, st.in_bytes, (size_t)chunk * row_bytes);
    if (!out_pinned) ensure_slots(st.out_slot, st.out_bytes, (size_t)chunk * row_bytes * nk);
    ensure_pinned(st.yc_host, st.yc_bytes, std::max<size_t>(64, (size_t)m1 * std::max(s.n_ctl, 1) * 8));
    while ((int)s.ev_chunk.size() < n_chunks + 2 || (int)s.ev_d2h.size() < n_chunks + 2) {
        cudaEvent_t e;
        RUVIII_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        s.ev_chunk.push_back(e);
        RUVIII_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        s.ev_upd.push_back(e);
        RUVIII_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        s.ev_d2h.push_back(e);
    }
    RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_START], s.stream));

    // ================= phase A: PRPS rows + control columns -> W =================
    // PRPS rows (replicate.data): through the in-slots when pageable
    if (n_ps > 0) {
        if (ps_pinned) {
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(s.PS.p, s.ld * 8, PS + s.g0, ldPS * 8, row_bytes, n_ps, cudaMemcpyHostToDevice, s.copy_stream));
        } else {
            int slot = 0;
            for (int r0 = 0; r0 < n_ps; r0 += chunk, slot = (slot + 1) % NS) {
                const int rr = std::min(chunk, n_ps - r0);
                if (r0 >= NS * chunk) RUVIII_CUDA_CHECK(cudaEventSynchronize(s.ev_chunk[slot]));       // slot still in flight
                pool_copy_rows(pool, st.n_threads, (char*)st.in_slot[slot], row_bytes, (const char*)(PS + (size_t)r0 * ldPS + s.g0),
                               (size_t)ldPS * 8, row_bytes, rr);
                RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(s.PS.p + (size_t)r0 * s.ld, s.ld * 8, st.in_slot[slot], row_bytes, row_bytes, rr,
                                                    cudaMemcpyHostToDevice, s.copy_stream));
                RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_chunk[slot], s.copy_stream));
            }
        }
    }
    RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_chunk[n_chunks], s.copy_stream));          // PS resident
    t_ps = now_ms();
    // K1, K2 on the PRPS rows start as soon as they have landed
    RUVIII_CUDA_CHECK(cudaStreamWaitEvent(s.stream, s.ev_chunk[n_chunks], 0));
    {
        RowSrc src{s.Y.p, s.ld, s.PS.p, s.ld, m1};
        launch_helmert(src, s.grp_start.p, s.grp_rows.p, h->n_groups_active, h->max_group, s.n, s.Y0.p, s.ld, s.stream);
        launch_gram(s.gp, s.Y0.p, s.ld, s.tile_map.p, s.gram_ws.p, s.G.p, s.tmap, s.stream);
    }
    h->launches += 3;
    // control columns of the sample rows: host threads copy Y[r, ctl] into one pinned block (bytes only; the log transform
    // of ruvIII.R:89 is applied to the block on the device)
    if (s.n_ctl > 0) {
        const int nc = s.n_ctl;
        const int* idx = s.ctl_idx_host.data();
        double* dst = st.yc_host;
        const double* src = Y + s.g0;
        const int parts = std::max(1, std::min(st.n_threads * 4, m1 / 64));
        pool.run(parts, [=](int i) {
            const int64_t r0 = (int64_t)m1 * i / parts, r1 = (int64_t)m1 * (i + 1) / parts;
            for (int64_t r = r0; r < r1; ++r) {
                const double* row = src + (size_t)r * ldY;
                double* d = dst + (size_t)r * nc;
                for (int c = 0; c < nc; ++c) d[c] = row[idx[c]];
            }
        });
        RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(s.Yc.p, s.ldc * 8, st.yc_host, (size_t)nc * 8, (size_t)nc * 8, m1, cudaMemcpyHostToDevice,
                                            s.copy_stream));
        if (h->prm.apply_log) launch_log2_inplace(s.Yc.p, s.ldc, m1, nc, h->prm.pseudo_count, s.copy_stream);
        // control columns of the PRPS rows (STACKED mode): from the device copy
        if (h->rows_w > m1) {
            RowSrc ps_src{nullptr, 0, s.PS.p, s.ld, 0};
            launch_ctl_gather(ps_src, h->rows_w - m1, s.ctl_idx.p, s.n_ctl, s.Yc.p + (size_t)m1 * s.ldc, s.ldc, 0, 0.0, s.copy_stream);
        }
        h->launches += 2;
    }
    RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_gather, s.copy_stream));
    t_gather = now_ms();
    {
        std::vector<double*> b{s.G.p};
        allreduce_sum(h, b, (size_t)R * R);
    }
    if (h->rank0 == 0) eigensolve(h);
    {
        std::vector<double*> b{s.U.p};
        broadcast0(h, b, (size_t)R * kp);
        b[0] = s.evals.p;
        broadcast0(h, b, (size_t)std::max(h->n_keep, kp) + 2);
    }
    {
        launch_project(s.Y0.p, s.ld, R, s.n, s.U.p, kp, s.rsplit, s.part.p, s.ld, s.stream);
        launch_alpha_finalize(s.part.p, s.rsplit, kp, s.n, s.ld, s.alpha.p, s.ctl_idx.p, s.n_ctl, s.acT.p, s.stream);
        RowSrc src{s.Y.p, s.ld, s.PS.p, s.ld, m1};
        RUVIII_CUDA_CHECK(cudaStreamWaitEvent(s.stream, s.ev_gather, 0));
        launch_ctl_products(src, h->rows_w, s.ctl_idx.p, s.n_ctl, s.acT.p, kp, s.AG.p, s.AG.p + (size_t)h->rows_w * kp,
                            s.n_ctl > 0 ? s.Yc.p : nullptr, s.ldc, s.stream);
    }
    {
        std::vector<double*> b{s.AG.p};
        allreduce_sum(h, b, (size_t)h->rows_w * kp + (size_t)kp * kp);
    }
    {
        const double* A = s.AG.p;
        const double* Gm = s.AG.p + (size_t)h->rows_w * kp;
        launch_chol_solve(A, h->rows_w, h->mu_rows, Gm, kk, kp, s.Linv.p, s.colmean.p, s.status.p, s.stream);
        launch_atilde(A, h->rows_w, s.colmean.p, s.Linv.p, kk, kp, s.At.p, s.stream);
        launch_alpha_tilde(s.alpha.p, s.Linv.p, kk, kp, s.n, s.ld, s.at.p, s.stream);
    }
    h->launches += 6;
    t_w = now_ms();

    // ================= phase B: stream the assay through =================
    const int64_t stride = (int64_t)m1 * s.ld;
    // the in-slots carried the PRPS rows: their DMA must have left them before the first chunks are staged into them
    if (!ps_pinned && !in_pinned && n_ps > 0) RUVIII_CUDA_CHECK(cudaEventSynchronize(s.ev_chunk[n_chunks]));
    auto drain = [&](int c) {                      // pinned out-slot -> the caller's (pageable) output blocks
        if (out_pinned) return;
        const int r0 = c * chunk, rr = std::min(chunk, m1 - r0);
        RUVIII_CUDA_CHECK(cudaEventSynchronize(s.ev_d2h[c]));
        const double* slot = st.out_slot[c % NS];
        for (int i = 0; i < nk; ++i)
            pool_copy_rows(pool, st.n_threads, (char*)(out_blocks[i] + (size_t)r0 * n), row_bytes, (const char*)(slot + (size_t)i * rr * n),
                           row_bytes, row_bytes, rr);
    };
    for (int c = 0; c < n_chunks; ++c) {
        const int r0 = c * chunk, rr = std::min(chunk, m1 - r0);
        // ---- upload chunk c ----
        if (in_pinned) {
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(s.Y.p + (size_t)r0 * s.ld, s.ld * 8, Y + (size_t)r0 * ldY + s.g0, ldY * 8, row_bytes, rr,
                                                cudaMemcpyHostToDevice, s.copy_stream));
        } else {
            const int slot = c % NS;
            if (c >= NS) RUVIII_CUDA_CHECK(cudaEventSynchronize(s.ev_chunk[c - NS]));     // the slot's previous upload has left it
            pool_copy_rows(pool, st.n_threads, (char*)st.in_slot[slot], row_bytes, (const char*)(Y + (size_t)r0 * ldY + s.g0), (size_t)ldY * 8,
                           row_bytes, rr);
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(s.Y.p + (size_t)r0 * s.ld, s.ld * 8, st.in_slot[slot], row_bytes, row_bytes, rr,
                                                cudaMemcpyHostToDevice, s.copy_stream));
        }
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_chunk[c], s.copy_stream));
        // ---- K6 on the chunk ----
        RUVIII_CUDA_CHECK(cudaStreamWaitEvent(s.stream, s.ev_chunk[c], 0));
        if (h->prm.apply_log) launch_log2_inplace(s.Y.p + (size_t)r0 * s.ld, s.ld, rr, s.n, h->prm.pseudo_count, s.stream);
        launch_update_smem(s.Y.p + (size_t)r0 * s.ld, s.ld, rr, s.n, s.At.p + (size_t)r0 * kp, kp, s.at.p, s.ld, kk, h->kmask, h->slot_of_k,
                           s.out.p + (size_t)r0 * s.ld, s.ld, stride, s.stream);
        for (auto& d : h->dup_slots)
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(s.out.p + (size_t)d.first * stride + (size_t)r0 * s.ld, s.ld * 8,
                                                s.out.p + (size_t)d.second * stride + (size_t)r0 * s.ld, s.ld * 8, row_bytes, rr,
                                                cudaMemcpyDeviceToDevice, s.stream));
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_upd[c], s.stream));
        // ---- download chunk c ----
        RUVIII_CUDA_CHECK(cudaStreamWaitEvent(s.d2h_stream, s.ev_upd[c], 0));
        if (!out_pinned && c >= NS) drain(c - NS);                                         // frees the slot this download lands in
        for (int i = 0; i < nk; ++i) {
            double* dst = out_pinned ? out_blocks[i] + (size_t)r0 * n : st.out_slot[c % NS] + (size_t)i * rr * n;
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(dst, row_bytes, s.out.p + (size_t)i * stride + (size_t)r0 * s.ld, s.ld * 8, row_bytes, rr,
                                                cudaMemcpyDeviceToHost, s.d2h_stream));
        }
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_d2h[c], s.d2h_stream));
    }
    if (!out_pinned)
        for (int c = std::max(0, n_chunks - NS); c < n_chunks; ++c) drain(c);
    h->launches += n_chunks * (h->prm.apply_log ? 2 : 1);
    {
        size_t off = 0;
        for (int i = 0; i < nk; ++i) {
            launch_w_from_atilde(s.At.p, h->rows_w, s.Linv.p, h->k_eff[i], kp, s.W.p + off, s.stream);
            off += (size_t)h->rows_w * h->k_eff[i];
        }
        h->launches += nk;
        if (fullalpha_out)
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(fullalpha_out + s.g0, h->n_local_total * 8, s.alpha.p, s.ld * 8, row_bytes, h->n_alpha_rows,
                                                cudaMemcpyDeviceToHost, s.stream));
        if (W_out) RUVIII_CUDA_CHECK(cudaMemcpyAsync(W_out, s.W.p, off * 8, cudaMemcpyDeviceToHost, s.stream));
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_UPDATE], s.stream));
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(&h->h_flags[1], s.status.p, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    }
    sync_all(h);
    const double t_done = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    if (trace)
        fprintf(stderr, "[stage trace] threads=%d chunks=%d x %d rows in_pinned=%d out_pinned=%d: ps_issued=%.2f gather_issued=%.2f W_queued=%.2f "
                        "total=%.2f (ms)\n", st.n_threads, n_chunks, chunk, (int)in_pinned, (int)out_pinned, t_ps, t_gather, t_w, (t_done - t_enter) * 1e3);
    if (h->rank0 == 0 && h->h_flags[0] != 0)
        throw Error(RUVIII_ERR_SOLVER, "cuSOLVER syevd devInfo = " + std::to_string(h->h_flags[0]));
    if (h->h_flags[1] != 0)
        throw Error(RUVIII_ERR_SINGULAR, "ac %*% t(ac) is not positive definite at pivot " + std::to_string(h->h_flags[1]) +
                                             " (solve() in ruvIII.R:144 would fail)");
    ruviii_info& I = h->info;
    std::memset(&I, 0, sizeof(I));
    I.struct_size = sizeof(ruviii_info);
    I.m = h->m; I.m1 = m1; I.n_ps = n_ps; I.n = (int)h->n_local_total; I.p = h->p; I.active_rows = h->active_rows; I.r = h->r;
    I.gram_order = h->gorder; I.n_ctl = (int)h->n_ctl_global; I.kmax_eff = kk; I.gram_side = 1;
    I.n_shards = 1; I.kernel_launches = h->launches;
    I.eig_solver = h->eig_solver_used; I.eig_iterations = h->eig_iterations;
    I.pipelined = in_pinned && out_pinned ? 2 : 3;          // 2 = staged path on pinned buffers, 3 = through the bounce buffers
    I.host_t_enter = t_enter; I.host_t_done = t_done;
    I.ms_total = (float)((t_done - t_enter) * 1e3);
    {
        std::vector<double> ev(kk + 1, 0.0);
        const int have = std::min(kk + 1, h->n_keep);
        RUVIII_CUDA_CHECK(cudaMemcpy(ev.data(), s.evals.p, have * sizeof(double), cudaMemcpyDeviceToHost));
        I.eig_top = ev[0]; I.eig_kmax = ev[kk - 1]; I.eig_next = have > kk ? ev[kk] : 0.0;
    }
    h->uploaded = true;
    h->logged = h->prm.apply_log != 0;
    h->ran = true;
    h->pipelined_last = I.pipelined;
}

}  // namespace ruviii

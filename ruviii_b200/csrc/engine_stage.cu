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

#include <atomic>
#include <condition_variable>
#include <functional>
#include <immintrin.h>
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
    // While a one-shot call is in flight the workers poll for the next batch instead of sleeping on the condition variable:
    // a batch is ~1 ms of copying, a futex wake-up of 15 threads costs a good part of that.
    std::atomic<bool> hot{false};
    std::atomic<uint64_t> gen_hint{0};

    explicit HostPool(int n) {
        for (int i = 0; i < n; ++i) threads.emplace_back([this] { worker(); });
    }
    ~HostPool() {
        hot.store(false, std::memory_order_release);
        {
            std::lock_guard<std::mutex> lk(m);
            stop = true;
        }
        cv.notify_all();
        for (auto& t : threads) t.join();
    }
    void set_hot(bool on) {
        hot.store(on, std::memory_order_release);
        if (!on) return;
        std::lock_guard<std::mutex> lk(m);
        cv.notify_all();
    }
    void worker() {
        uint64_t seen = 0;
        for (;;) {
            bool spun_out = false;
            if (hot.load(std::memory_order_acquire)) {
                // poll (with pauses) until a new batch is published or the call ends -- but only for ~100 us: when the cores
                // are oversubscribed (several ranks of one host, each with its own pool) a spinning thread steals the core a
                // working one needs (r2m2: the pageable path took 162 ms on two ranks against 80 ms on one)
                int spins = 0;
                while (hot.load(std::memory_order_acquire) && gen_hint.load(std::memory_order_acquire) == seen && spins < 40000) {
                    _mm_pause();
                    ++spins;
                }
                spun_out = spins >= 40000;
            }
            std::unique_lock<std::mutex> lk(m);
            if (!hot.load(std::memory_order_acquire) || spun_out)
                cv.wait(lk, [&] { return stop || (gen != seen && next < n_items) || (!spun_out && hot.load(std::memory_order_acquire)); });
            if (stop) return;
            if (gen == seen) continue;
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
        gen_hint.store(gen, std::memory_order_release);
        cv.notify_all();
        while (next < n_items) {
            const int i = next++;
            lk.unlock();
            job(i);
            lk.lock();
            --pending;
        }
        if (hot.load(std::memory_order_acquire)) {
            while (pending != 0) {          // the stragglers are at most one item each: poll
                lk.unlock();
                _mm_pause();
                lk.lock();
            }
        } else {
            cv_done.wait(lk, [&] { return pending == 0; });
        }
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
    return std::max(1, std::min(32, avail));
}

// Bulk copy with non-temporal stores: the destination (a pinned slot the DMA engine reads next, or the caller's output
// that nobody reads soon) should not be pulled into the cache first -- a plain store costs a read-for-ownership of every
// destination line, i.e. a third more memory traffic on a path that is bound by host memory bandwidth.
__attribute__((target("avx2"))) void copy_nt_avx2(char* dst, const char* src, size_t bytes) {
    size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
    if (head > bytes) head = bytes;
    std::memcpy(dst, src, head);
    dst += head; src += head; bytes -= head;
    const size_t blocks = bytes / 128;
    for (size_t i = 0; i < blocks; ++i) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + 64));
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + 96), d);
        src += 128; dst += 128;
    }
    std::memcpy(dst, src, bytes - blocks * 128);
    _mm_sfence();
}

void bulk_copy(char* dst, const char* src, size_t bytes) {
    static const bool nt = __builtin_cpu_supports("avx2") && env_int("RUVIII_NT_COPY", 1) != 0;
    if (nt && bytes >= 16384) copy_nt_avx2(dst, src, bytes);
    else std::memcpy(dst, src, bytes);
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
            bulk_copy(dst + a, src + a, b - a);
        });
        return;
    }
    const int parts = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)n_threads * 2, rows));
    pool.run(parts, [=](int i) {
        const int64_t r0 = rows * i / parts, r1 = rows * (i + 1) / parts;
        for (int64_t r = r0; r < r1; ++r) bulk_copy(dst + (size_t)r * dst_pitch, src + (size_t)r * src_pitch, row_bytes);
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
    struct HotGuard { HostPool& p; explicit HotGuard(HostPool& q) : p(q) { p.set_hot(true); } ~HotGuard() { p.set_hot(false); } } hot_guard(pool);
    static const bool trace = env_int("RUVIII_PIPE_TRACE", 0) != 0;
    auto now_ms = [&] { return 1e3 * (std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count() - t_enter); };
    double t_gather = 0, t_ps = 0, t_w = 0;

    const bool in_pinned = host_pinned(Y) != nullptr;
    bool out_pinned = true;
    for (int i = 0; i < nk; ++i) out_pinned = out_pinned && host_pinned(out_blocks[i]) != nullptr;
    const bool ps_pinned = n_ps == 0 || host_pinned(PS) != nullptr;

    // ---- row chunks: in-slot <= ~96 MB, out-slot (nk blocks) <= ~192 MB ----
    const size_t row_bytes = (size_t)n * 8;
    int64_t rows = std::min<int64_t>((int64_t)(96u << 20) / (int64_t)row_bytes, (int64_t)(192u << 20) / (int64_t)(row_bytes * nk));
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
    // Two schedules.  DUPLEX (default): one pass of the host threads over Y copies out the control columns first, W is known
    // early, and upload / update / download of the chunks overlap.  TWO-PHASE (RUVIII_STAGE_TWO_PHASE=1, pageable buffers
    // only): the control columns are picked up WHILE each row is staged in (no separate pass over Y), W follows the last
    // chunk, then the chunks are updated and drained.  Measured at C3 on the 16-core host (r2h): duplex 77-81 ms, two-phase
    // 90-94 ms -- with bounce buffers every byte crosses host memory three times per direction (caller -> slot by the CPU,
    // slot -> device by the DMA engine), the path is bound by host memory bandwidth (~160 GB/s of traffic), and the duplex
    // schedule simply keeps both directions' traffic going at once.
    const bool two_phase = !in_pinned && !out_pinned && env_int("RUVIII_STAGE_TWO_PHASE", 0) != 0;
    // the in-slots carried the PRPS rows: their DMA must have left them before the first chunks are staged into them
    if (!ps_pinned && !in_pinned && n_ps > 0) RUVIII_CUDA_CHECK(cudaEventSynchronize(s.ev_chunk[n_chunks]));
    auto stage_in_chunk = [&](int c, bool with_gather) {          // caller's rows -> pinned slot (+ control columns) -> device
        const int r0 = c * chunk, rr = std::min(chunk, m1 - r0);
        const int slot = c % NS;
        if (c >= NS) RUVIII_CUDA_CHECK(cudaEventSynchronize(s.ev_chunk[c - NS]));     // the slot's previous upload has left it
        if (!with_gather || s.n_ctl == 0) {
            pool_copy_rows(pool, st.n_threads, (char*)st.in_slot[slot], row_bytes, (const char*)(Y + (size_t)r0 * ldY + s.g0), (size_t)ldY * 8,
                           row_bytes, rr);
        } else {
            const int nc = s.n_ctl;
            const int* idx = s.ctl_idx_host.data();
            double* ycd = st.yc_host + (size_t)r0 * nc;
            char* dslot = (char*)st.in_slot[slot];
            const double* src = Y + (size_t)r0 * ldY + s.g0;
            const int parts = std::max(1, std::min(st.n_threads * 2, rr));
            pool.run(parts, [=](int i) {
                const int64_t a = (int64_t)rr * i / parts, b = (int64_t)rr * (i + 1) / parts;
                for (int64_t r = a; r < b; ++r) {
                    const double* row = src + (size_t)r * ldY;
                    bulk_copy(dslot + (size_t)r * row_bytes, (const char*)row, row_bytes);
                    double* d = ycd + (size_t)r * nc;
                    for (int q = 0; q < nc; ++q) d[q] = row[idx[q]];          // the row has just been read: cache hits
                }
            });
        }
        RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(s.Y.p + (size_t)r0 * s.ld, s.ld * 8, st.in_slot[slot], row_bytes, row_bytes, rr,
                                            cudaMemcpyHostToDevice, s.copy_stream));
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_chunk[c], s.copy_stream));
    };
    if (two_phase)
        for (int c = 0; c < n_chunks; ++c) stage_in_chunk(c, true);
    // Pinned input: nothing on the host side stands between the caller's memory and the DMA engine, so the first chunk
    // uploads are queued right behind the PRPS rows and stream while the host threads pick out the control columns.  Only
    // ~256 MB of them: the copy engine serves its queue in order ACROSS streams, so the control block queued later waits
    // for everything queued here (with every chunk queued up front W came after the whole upload: 75 ms instead of 51, r2i).
    int n_early = 0;
    if (in_pinned) {
        auto upload_pinned = [&](int c) {
            const int r0 = c * chunk, rr = std::min(chunk, m1 - r0);
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(s.Y.p + (size_t)r0 * s.ld, s.ld * 8, Y + (size_t)r0 * ldY + s.g0, ldY * 8, row_bytes, rr,
                                                cudaMemcpyHostToDevice, s.copy_stream));
            RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_chunk[c], s.copy_stream));
        };
        const size_t early_bytes = (size_t)env_int("RUVIII_STAGE_EARLY_MB", 256) << 20;
        while (n_early < n_chunks && (size_t)n_early * chunk * row_bytes < early_bytes) upload_pinned(n_early++);
    }
    // control columns of the sample rows: host threads copy Y[r, ctl] into one pinned block (bytes only; the log transform
    // of ruvIII.R:89 is applied to the block on the device)
    if (s.n_ctl > 0) {
        const int nc = s.n_ctl;
        if (!two_phase) {
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
        }
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
        launch_project(s.Y0.p, s.ld, R, s.n, s.U.p, kp, s.rsplit, s.part.p, s.ld, s.stream, kk);
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
    launch_solve(h, s);
    h->launches += 3;
    t_w = now_ms();

    // ================= phase B: stream the assay through =================
    const int64_t stride = (int64_t)m1 * s.ld;
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
            if (c >= n_early) {
                RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(s.Y.p + (size_t)r0 * s.ld, s.ld * 8, Y + (size_t)r0 * ldY + s.g0, ldY * 8, row_bytes, rr,
                                                    cudaMemcpyHostToDevice, s.copy_stream));
                RUVIII_CUDA_CHECK(cudaEventRecord(s.ev_chunk[c], s.copy_stream));
            }
        } else if (!two_phase) {
            stage_in_chunk(c, false);
        }
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
        for (int i = 0; i < nk; ++i) off += (size_t)h->rows_w * h->k_eff[i];
        if (W_out) launch_w_blocks(h, s);
        if (fullalpha_out)
            RUVIII_CUDA_CHECK(cudaMemcpy2DAsync(fullalpha_out + s.g0, h->n_local_total * 8, s.alpha.p, s.ld * 8, row_bytes, h->n_alpha_rows,
                                                cudaMemcpyDeviceToHost, s.stream));
        if (W_out) RUVIII_CUDA_CHECK(cudaMemcpyAsync(W_out, s.W.p, off * 8, cudaMemcpyDeviceToHost, s.stream));
        RUVIII_CUDA_CHECK(cudaEventRecord(s.ev[E_UPDATE], s.stream));
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(&h->h_flags[1], s.status.p, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    }
    sync_all(h);
    const double t_done = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    if (h->multi_process) p2p_check(h);
    if (trace)
        fprintf(stderr, "[stage trace] %s threads=%d chunks=%d x %d rows in_pinned=%d out_pinned=%d: ps_issued=%.2f gather_issued=%.2f W_queued=%.2f "
                        "total=%.2f (ms)\n", two_phase ? "two-phase" : "duplex", st.n_threads, n_chunks, chunk, (int)in_pinned, (int)out_pinned, t_ps, t_gather, t_w, (t_done - t_enter) * 1e3);
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
    h->yc_valid = true;                  // the whole control block was gathered on the way in
    h->ran = true;
    h->pipelined_last = I.pipelined;
}

}  // namespace ruviii

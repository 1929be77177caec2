// Sum-allreduce of the small matrices of the pass over NVLink peer memory, written as ONE kernel per GPU instead of an NCCL
// call (SURVEY.md 8e collectives 1 and 3: the partial Gram, R^2 doubles, and [A | ac ac' | flag], m kp + kp^2 + 1 doubles).
//
// At these sizes (5 MB and 1.6 MB at C3) the exchange is latency-bound: ncclAllReduce costs 77 us and 45 us on 8 GPUs, 19 %
// of the 0.66 ms pass.  With one process per GPU every rank maps the others' exchange buffer (cudaIpc handles, swapped once
// through NCCL) and runs a two-shot allreduce by posted P2P STORES only -- no remote loads on the critical path:
//   phase 0  rank r pushes its contribution to slice q (the q-th 1/W of the elements) into rank q's inbox[r]    (W - 1 stores)
//   signal / wait A   (system-scope counters in the owner's memory, one increment per CTA)
//   phase 1  rank q adds the W contributions of its slice in rank order 0..W-1 (its own one straight from the matrix)
//            and pushes the sum straight into every rank's copy of the matrix                                     (W - 1 stores)
//   signal / wait B: the matrix (which lives in the exchange buffer, so that peers can store into it) is complete.
// Every element is summed by exactly ONE rank in a FIXED order and then distributed, so all ranks hold the same bits -- which
// is what lets every rank run the eigensolver redundantly instead of waiting for a broadcast (engine.cu).
// Buffers alternate between two inbox parities, so a fast rank can start the next exchange while a slow one still reads.
// A wait that is not satisfied within ~4 s (a peer died before its launch) sets an error flag instead of spinning for ever.
#include "engine_internal.cuh"

namespace ruviii {

namespace {

constexpr int kP2PThreads = 256;
constexpr int kMaxPeers = 8;

struct P2PArgs {
    unsigned char* base[kMaxPeers];  // mapped exchange buffer of every rank (base[me] is local)
    size_t inbox_off;                // byte offset of the inbox area of the current parity: W slots of `slice` doubles
    size_t buf_off;                  // byte offset of the matrix being reduced (the same in every rank's buffer)
    int W, me;
    int64_t count, slice;            // elements, elements per owner (multiple of 2)
    unsigned calls;                  // exchanges so far incl. this one: every rank's counter reaches it once per exchange
    int* error;                      // local: set when a wait timed out
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_relaxed_sys(unsigned* p) {
    asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// every CTA waits until all W ranks have reached `target` in the local counter array
__device__ __forceinline__ bool wait_all(const unsigned* mine, int W, unsigned target, int* error) {
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    if (threadIdx.x < W) {
        // relaxed polls (an acquire at system scope on every iteration costs a fence each time), one acquire fence at the end
        const unsigned long long t0 = globaltimer_ns();
        unsigned spins = 0;
        while ((int)(ld_relaxed_sys(mine + threadIdx.x) - target) < 0) {
            if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > 4000000000ull) { ok = 0; *error = 1; break; }
        }
        asm volatile("fence.acq_rel.sys;" ::: "memory");
    }
    __syncthreads();
    return ok != 0;
}

// One arrival per RANK crosses NVLink, not one per CTA: the CTAs of a rank meet at a local counter (each after a system-scope
// fence of its own peer stores) and the last one to arrive signals every peer.  (With every CTA signalling every peer, 148
// remote atomics per peer piled up on one address: the 1.6 MB exchange took 38 us on two GPUs against NCCL's 22, r2p2b.)
__device__ __forceinline__ void signal_peers(const P2PArgs& a, size_t flag_off, size_t local_off, unsigned local_target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();                                      // this CTA's peer stores are performed
        unsigned* local = reinterpret_cast<unsigned*>(a.base[a.me] + local_off);
        const unsigned old = atomicAdd(local, 1u);
        if (old + 1 == local_target) {                               // last CTA of this rank
            // ONE system-scope fence, then the W increments as posted relaxed atomics (fence + relaxed write = a release
            // pattern).  A release on each of them made every increment wait for the previous remote one to be performed:
            // 8 x ~5 us per signal, 92 us for the 1.6 MB exchange on 8 GPUs against NCCL's 39 (r2p8).
            __threadfence_system();
            for (int q = 0; q < a.W; ++q) red_relaxed_sys(reinterpret_cast<unsigned*>(a.base[q] + flag_off) + a.me);
        }
    }
}

// flags (first 4 KB of every exchange buffer): [0, 1024) counters A, one unsigned per source rank, written by the peers;
// [1024, 2048) local arrival counter A; [2048, 3072) counters B; [3072, 4096) local arrival counter B
__global__ void __launch_bounds__(kP2PThreads) p2p_allreduce_kernel(P2PArgs a) {
    const int tid = threadIdx.x, cta = blockIdx.x, nct = gridDim.x, W = a.W, me = a.me;
    const int64_t slice2 = a.slice / 2;                          // double2 elements per owner
    const int64_t per = (slice2 + nct - 1) / nct;                // per CTA within a slice
    const int64_t v0 = (int64_t)cta * per, v1 = min(slice2, v0 + per);
    const int64_t count2 = (a.count + 1) / 2;                    // the buffers extend to an even element count
    double2* mine = reinterpret_cast<double2*>(a.base[me] + a.buf_off);
    // ---- phase 0: push my contribution to every other owner's inbox (posted stores over NVLink) ----
    for (int dq = 1; dq < W; ++dq) {
        const int q = (me + dq) % W;                             // staggered so that the ranks do not all hit the same peer first
        const double2* src = mine + (int64_t)q * slice2;
        double2* dst = reinterpret_cast<double2*>(a.base[q] + a.inbox_off) + (int64_t)me * slice2;
        const int64_t lim = min(v1, count2 - (int64_t)q * slice2);
        int64_t i = v0 + tid;
        for (; i + 3 * kP2PThreads < lim; i += 4 * kP2PThreads) {
            const double2 x0 = src[i], x1 = src[i + kP2PThreads], x2 = src[i + 2 * kP2PThreads], x3 = src[i + 3 * kP2PThreads];
            dst[i] = x0; dst[i + kP2PThreads] = x1; dst[i + 2 * kP2PThreads] = x2; dst[i + 3 * kP2PThreads] = x3;
        }
        for (; i < lim; i += kP2PThreads) dst[i] = src[i];
    }
    signal_peers(a, 0, 1024, a.calls * (unsigned)nct);
    if (!wait_all(reinterpret_cast<unsigned*>(a.base[me]), W, a.calls, a.error)) return;
    // ---- phase 1: sum my slice in rank order 0..W-1 and push the sum straight into every rank's copy of the matrix ----
    {
        double2* own = mine + (int64_t)me * slice2;
        const double2* in = reinterpret_cast<const double2*>(a.base[me] + a.inbox_off);
        const int64_t lim = min(v1, count2 - (int64_t)me * slice2);
        for (int64_t i = v0 + tid; i < lim; i += kP2PThreads) {
            double2 s = make_double2(0.0, 0.0);
            for (int p = 0; p < W; ++p) {
                const double2 v = p == me ? own[i] : __ldcg(in + (int64_t)p * slice2 + i);
                s.x += v.x;
                s.y += v.y;
            }
            own[i] = s;
            for (int dq = 1; dq < W; ++dq) {
                const int q = (me + dq) % W;
                (reinterpret_cast<double2*>(a.base[q] + a.buf_off) + (int64_t)me * slice2)[i] = s;
            }
        }
    }
    signal_peers(a, 2048, 3072, a.calls * (unsigned)nct);
    wait_all(reinterpret_cast<unsigned*>(a.base[me] + 2048), W, a.calls, a.error);
}

}  // namespace

struct P2PState {
    void* xbuf = nullptr;            // [flags 4 KB | inbox parity 0 | inbox parity 1 | matrix areas ...]
    size_t cap_inbox = 0;            // doubles per inbox area
    size_t cap_mat = 0;              // doubles of the matrix areas together
    size_t mat_used = 0;             // doubles handed out by p2p_carve since the last setup
    void* peer[kMaxPeers] = {};      // mapped exchange buffers of all ranks (peer[me] = xbuf)
    DevBuf<unsigned char> handles;   // W x 64 bytes, gathered
    int* h_error = nullptr;          // pinned, mapped: the kernel writes it only when a wait times out; the host reads it directly
    int* d_error = nullptr;
    unsigned calls = 0;              // exchanges done so far (the same on every rank)
    int n_ctas = 64;
    bool ready = false;
    ~P2PState() {
        for (int i = 0; i < kMaxPeers; ++i)
            if (peer[i] && peer[i] != xbuf) cudaIpcCloseMemHandle(peer[i]);
        if (xbuf) cudaFree(xbuf);
        if (h_error) cudaFreeHost(h_error);
    }
};
void p2p_state_destroy(P2PState* p) { delete p; }

// (Re)creates the exchange buffer -- inboxes of `inbox_doubles` each and `mat_doubles` for the matrices that are reduced in
// place (they LIVE in the exchange buffer, p2p_carve, so that a peer can store into them) -- and swaps the IPC handles.
// COLLECTIVE: every rank calls it from do_plan with the same arguments.  Returns false (NCCL stays in charge) when peer
// mapping is not possible.
bool p2p_setup(ruviii_handle* h, size_t inbox_doubles, size_t mat_doubles) {
    static const bool off = env_int("RUVIII_P2P", 1) == 0;
    if (off || !h->multi_process || h->world < 2 || h->world > kMaxPeers) return false;
    Shard& s = h->shards[0];
    RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
    if (!h->p2p) h->p2p = new P2PState();
    P2PState& st = *h->p2p;
    inbox_doubles = (size_t)round_up((int64_t)inbox_doubles + 2 * h->world, 512);
    mat_doubles = (size_t)round_up((int64_t)mat_doubles + 64, 512);
    st.mat_used = 0;
    if (st.ready && inbox_doubles <= st.cap_inbox && mat_doubles <= st.cap_mat) return true;
    // every rank reaches this point with the same capacities (they derive from the plan), so all of them reallocate together
    sync_all(h);
    for (int i = 0; i < kMaxPeers; ++i) {
        if (st.peer[i] && st.peer[i] != st.xbuf) cudaIpcCloseMemHandle(st.peer[i]);
        st.peer[i] = nullptr;
    }
    if (st.xbuf) { cudaFree(st.xbuf); st.xbuf = nullptr; }
    st.ready = false;
    st.calls = 0;
    st.n_ctas = std::max(8, std::min(h->n_sms, env_int("RUVIII_P2P_CTAS", h->n_sms)));
    const size_t bytes = 4096 + (2 * inbox_doubles + mat_doubles) * sizeof(double);
    int ok_local = 1;
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof(mine));
    if (cudaMalloc(&st.xbuf, bytes) != cudaSuccess || cudaMemset(st.xbuf, 0, bytes) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine, st.xbuf) != cudaSuccess) {
        cudaGetLastError();
        ok_local = 0;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    st.handles.alloc((size_t)h->world * 64 + 64);
    if (!st.h_error) {
        RUVIII_CUDA_CHECK(cudaHostAlloc((void**)&st.h_error, 64, cudaHostAllocMapped));
        std::memset(st.h_error, 0, 64);
        RUVIII_CUDA_CHECK(cudaHostGetDevicePointer((void**)&st.d_error, st.h_error, 0));
    }
    RUVIII_CUDA_CHECK(cudaMemcpy(st.handles.p + (size_t)h->rank0 * 64, &mine, 64, cudaMemcpyHostToDevice));
    RUVIII_NCCL_CHECK(ncclAllGather(st.handles.p + (size_t)h->rank0 * 64, st.handles.p, 64, ncclUint8, s.comm, s.stream));
    RUVIII_CUDA_CHECK(cudaStreamSynchronize(s.stream));
    std::vector<unsigned char> all((size_t)h->world * 64);
    RUVIII_CUDA_CHECK(cudaMemcpy(all.data(), st.handles.p, all.size(), cudaMemcpyDeviceToHost));
    for (int r = 0; r < h->world && ok_local; ++r) {
        if (r == h->rank0) { st.peer[r] = st.xbuf; continue; }
        cudaIpcMemHandle_t hd;
        std::memcpy(&hd, all.data() + (size_t)r * 64, 64);
        bool zero = true;
        for (int b = 0; b < 64; ++b) zero = zero && all[(size_t)r * 64 + b] == 0;
        if (zero || cudaIpcOpenMemHandle(&st.peer[r], hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            st.peer[r] = nullptr;
            ok_local = 0;
        }
    }
    // all ranks must take the same route: agree through one more (tiny) NCCL exchange
    double flag = ok_local ? 0.0 : 1.0;
    DevBuf<double> agree;
    agree.alloc(2);
    RUVIII_CUDA_CHECK(cudaMemcpy(agree.p, &flag, 8, cudaMemcpyHostToDevice));
    RUVIII_NCCL_CHECK(ncclAllReduce(agree.p, agree.p, 1, ncclDouble, ncclSum, s.comm, s.stream));
    RUVIII_CUDA_CHECK(cudaStreamSynchronize(s.stream));
    RUVIII_CUDA_CHECK(cudaMemcpy(&flag, agree.p, 8, cudaMemcpyDeviceToHost));
    agree.release();
    st.cap_inbox = inbox_doubles;
    st.cap_mat = mat_doubles;
    st.ready = flag == 0.0;
    return st.ready;
}

// `count` doubles of the exchange buffer's matrix area (zeroed), or nullptr when the route is not available / the area is full
double* p2p_carve(ruviii_handle* h, size_t count) {
    P2PState* st = h->p2p;
    if (!st || !st->ready) return nullptr;
    count = (size_t)round_up((int64_t)count + 2, 64);
    if (st->mat_used + count > st->cap_mat) return nullptr;
    double* p = reinterpret_cast<double*>(static_cast<unsigned char*>(st->xbuf) + 4096) + 2 * st->cap_inbox + st->mat_used;
    st->mat_used += count;
    RUVIII_CUDA_CHECK(cudaMemset(p, 0, count * sizeof(double)));
    RUVIII_CUDA_CHECK(cudaStreamSynchronize(cudaStreamLegacy));
    return p;
}

// In-place sum over all ranks of `count` doubles at `buf`, which must lie in the exchange buffer's matrix area (p2p_carve).
// Returns false when it does not, or the route is not set up: the caller then uses NCCL.
bool p2p_allreduce(ruviii_handle* h, double* buf, size_t count) {
    P2PState* st = h->p2p;
    if (!st || !st->ready || !h->multi_process) return false;
    // the same decision on every rank (it depends on the payload only): the inbox parity below counts calls
    static const int64_t min_bytes = (int64_t)env_int("RUVIII_P2P_MIN_KB", 3072) * 1024;
    if ((int64_t)count * 8 < min_bytes) return false;        // small payloads: NCCL's low-latency protocol wins (r2p2: 22 vs 38 us)
    const int W = h->world;
    const int64_t slice = round_up(ceil_div((int64_t)count, W), 2);
    if ((size_t)slice * W > st->cap_inbox) return false;
    unsigned char* base = static_cast<unsigned char*>(st->xbuf);
    unsigned char* mat0 = base + 4096 + 2 * st->cap_inbox * sizeof(double);
    unsigned char* b = reinterpret_cast<unsigned char*>(buf);
    if (b < mat0 || b + (count + 1) * sizeof(double) > mat0 + st->cap_mat * sizeof(double)) return false;
    Shard& s = h->shards[0];
    P2PArgs a{};
    a.W = W;
    a.me = h->rank0;
    a.count = (int64_t)count;
    a.slice = slice;
    a.error = st->d_error;
    st->calls += 1;
    a.calls = st->calls;
    a.inbox_off = 4096 + (size_t)(st->calls & 1u) * st->cap_inbox * sizeof(double);
    a.buf_off = (size_t)(b - base);
    for (int r = 0; r < W; ++r) a.base[r] = static_cast<unsigned char*>(st->peer[r]);
    p2p_allreduce_kernel<<<st->n_ctas, kP2PThreads, 0, s.stream>>>(a);
    RUVIII_LAUNCH_CHECK();
    return true;
}

// after a synchronisation: did an exchange time out?
void p2p_check(ruviii_handle* h) {
    P2PState* st = h->p2p;
    if (!st || !st->ready) return;
    if (*static_cast<volatile int*>(st->h_error)) {
        *st->h_error = 0;
        st->ready = false;
        throw Error(RUVIII_ERR_NCCL, "peer-memory allreduce: a rank did not arrive within 4 s (did a peer fail before its launch?)");
    }
}

}  // namespace ruviii

// Sum-allreduce of the small matrices of the pass over NVLink peer memory, written as ONE kernel per GPU instead of an NCCL
// call (SURVEY.md 8e collectives 1 and 3: the partial Gram, R^2 doubles, and [A | ac ac' | flag], m kp + kp^2 + 1 doubles).
//
// At these sizes (5 MB and 1.6 MB at C3) the exchange is latency-bound: ncclAllReduce costs 77 us and 45 us on 8 GPUs, 19 %
// of the 0.66 ms pass.  With one process per GPU every rank maps the others' exchange buffer (cudaIpc handles, swapped once
// through NCCL) and runs a two-shot allreduce by posted P2P STORES only -- no remote loads on the critical path:
//   phase 0  rank r pushes its contribution to slice q (the q-th 1/W of the elements) into rank q's inbox[r]    (W - 1 stores)
//   signal / wait A   (system-scope counters in the owner's memory, one increment per CTA)
//   phase 1  rank q adds the W contributions of its slice in rank order 0..W-1 (its own one straight from the source
//            buffer) and pushes the sum into every rank's result area                                             (W stores)
//   signal / wait B, then every rank copies the complete result over its source buffer (in place, like ncclAllReduce).
// Every element is summed by exactly ONE rank in a FIXED order and then distributed, so all ranks hold the same bits -- which
// is what lets every rank run the eigensolver redundantly instead of waiting for a broadcast (engine.cu).
// Buffers alternate between two inbox parities, so a fast rank can start the next exchange while a slow one still reads.
// A wait that is not satisfied within ~4 s (a peer died before its launch) sets an error flag instead of spinning for ever.
#include "engine_internal.cuh"

namespace ruviii {

namespace {

constexpr int kP2PCtas = 32;
constexpr int kP2PThreads = 256;
constexpr int kMaxPeers = 8;

struct P2PArgs {
    double* inbox[kMaxPeers];        // rank q's inbox area of the current parity: W slots of `slice` doubles
    double* result[kMaxPeers];       // rank q's result area
    unsigned* flag_a[kMaxPeers];     // rank q's counters: [me] incremented by every CTA of mine
    unsigned* flag_b[kMaxPeers];
    int W, me;
    int64_t count, slice;            // elements, elements per owner (multiple of 2)
    unsigned target;                 // counter value every peer's slot reaches when all its CTAs of this call have signalled
    double* buf;                     // local source, overwritten with the result
    int* error;                      // local: set when a wait timed out
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_sys(unsigned* p) {
    asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
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
        const unsigned long long t0 = globaltimer_ns();
        while ((int)(ld_acquire_sys(mine + threadIdx.x) - target) < 0) {
            if (globaltimer_ns() - t0 > 4000000000ull) { ok = 0; *error = 1; break; }
        }
    }
    __syncthreads();
    return ok != 0;
}

__global__ void __launch_bounds__(kP2PThreads) p2p_allreduce_kernel(P2PArgs a) {
    const int tid = threadIdx.x, cta = blockIdx.x, W = a.W, me = a.me;
    const int64_t slice2 = a.slice / 2;                          // double2 elements per owner
    const int64_t per = (slice2 + kP2PCtas - 1) / kP2PCtas;     // per CTA within a slice
    const int64_t v0 = (int64_t)cta * per, v1 = min(slice2, v0 + per);
    const int64_t count2 = (a.count + 1) / 2;                    // buf is padded to an even element count by the caller
    // ---- phase 0: push my contribution to every other owner's inbox ----
    for (int dq = 1; dq < W; ++dq) {
        const int q = (me + dq) % W;                             // staggered so that the ranks do not all hit the same peer first
        const double2* src = reinterpret_cast<const double2*>(a.buf) + (int64_t)q * slice2;
        double2* dst = reinterpret_cast<double2*>(a.inbox[q]) + (int64_t)me * slice2;
        for (int64_t i = v0 + tid; i < v1; i += kP2PThreads)
            if ((int64_t)q * slice2 + i < count2) dst[i] = src[i];
    }
    __syncthreads();
    if (tid < W && tid != me) { __threadfence_system(); red_release_sys(a.flag_a[tid] + me); }
    if (tid == me) red_release_sys(a.flag_a[me] + me);
    if (!wait_all(a.flag_a[me], W, a.target, a.error)) return;
    // ---- phase 1: sum my slice in rank order, push the result everywhere ----
    {
        const double2* own = reinterpret_cast<const double2*>(a.buf) + (int64_t)me * slice2;
        const double2* in = reinterpret_cast<const double2*>(a.inbox[me]);
        for (int64_t i = v0 + tid; i < v1; i += kP2PThreads) {
            if ((int64_t)me * slice2 + i >= count2) break;
            double2 s = make_double2(0.0, 0.0);
            for (int p = 0; p < W; ++p) {
                const double2 v = p == me ? own[i] : __ldcg(in + (int64_t)p * slice2 + i);
                s.x += v.x;
                s.y += v.y;
            }
            for (int q = 0; q < W; ++q) (reinterpret_cast<double2*>(a.result[q]) + (int64_t)me * slice2)[i] = s;
        }
    }
    __syncthreads();
    if (tid < W && tid != me) { __threadfence_system(); red_release_sys(a.flag_b[tid] + me); }
    if (tid == me) { __threadfence(); red_release_sys(a.flag_b[me] + me); }
    if (!wait_all(a.flag_b[me], W, a.target, a.error)) return;
    // ---- copy the complete result over the source (this CTA's share of the whole array) ----
    {
        const int64_t per_all = (count2 + kP2PCtas - 1) / kP2PCtas;
        const int64_t c0 = (int64_t)cta * per_all, c1 = min(count2, c0 + per_all);
        const double2* res = reinterpret_cast<const double2*>(a.result[me]);
        double2* out = reinterpret_cast<double2*>(a.buf);
        for (int64_t i = c0 + tid; i < c1; i += kP2PThreads) out[i] = __ldcg(res + i);
    }
}

}  // namespace

struct P2PState {
    void* xbuf = nullptr;            // [flags 4 KB | inbox parity 0 | inbox parity 1 | result], each area `cap` doubles
    size_t cap = 0;                  // doubles per area
    void* peer[kMaxPeers] = {};      // mapped exchange buffers of all ranks (peer[me] = xbuf)
    DevBuf<unsigned char> handles;   // W x 64 bytes, gathered
    int* h_error = nullptr;          // pinned, mapped: the kernel writes it only when a wait times out; the host reads it directly
    int* d_error = nullptr;
    unsigned calls = 0;              // exchanges done so far (the same on every rank)
    bool ready = false;
    int device = 0;
    ~P2PState() {
        for (int i = 0; i < kMaxPeers; ++i)
            if (peer[i] && peer[i] != xbuf) cudaIpcCloseMemHandle(peer[i]);
        if (xbuf) cudaFree(xbuf);
        if (h_error) cudaFreeHost(h_error);
    }
};
void p2p_state_destroy(P2PState* p) { delete p; }

// (Re)creates the exchange buffers for `cap_doubles` per area and swaps the IPC handles.  COLLECTIVE: every rank calls it from
// do_plan with the same argument.  Returns false (and leaves NCCL in charge) when peer mapping is not possible.
bool p2p_setup(ruviii_handle* h, size_t cap_doubles) {
    static const bool off = env_int("RUVIII_P2P", 1) == 0;
    if (off || !h->multi_process || h->world < 2 || h->world > kMaxPeers) return false;
    Shard& s = h->shards[0];
    RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
    if (!h->p2p) h->p2p = new P2PState();
    P2PState& st = *h->p2p;
    cap_doubles = (size_t)round_up((int64_t)cap_doubles + 2 * h->world, 512);
    if (st.ready && cap_doubles <= st.cap) return true;
    // every rank reaches this point with the same capacities (they derive from the plan), so all of them reallocate together
    sync_all(h);
    for (int i = 0; i < kMaxPeers; ++i) {
        if (st.peer[i] && st.peer[i] != st.xbuf) cudaIpcCloseMemHandle(st.peer[i]);
        st.peer[i] = nullptr;
    }
    if (st.xbuf) { cudaFree(st.xbuf); st.xbuf = nullptr; }
    st.ready = false;
    st.calls = 0;
    st.device = s.device;
    const size_t bytes = 4096 + 3 * cap_doubles * sizeof(double);
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
    st.cap = cap_doubles;
    st.ready = flag == 0.0;
    return st.ready;
}

// In-place sum over all ranks of `count` doubles at `buf` (device memory of this rank; the allocation must extend to an even
// element count).  Returns false when the P2P route is not set up or too small: the caller then uses NCCL.
bool p2p_allreduce(ruviii_handle* h, double* buf, size_t count) {
    P2PState* st = h->p2p;
    if (!st || !st->ready || !h->multi_process) return false;
    const int W = h->world;
    const int64_t slice = round_up(ceil_div((int64_t)count, W), 2);
    if ((size_t)slice * W > st->cap) return false;
    Shard& s = h->shards[0];
    P2PArgs a{};
    a.W = W;
    a.me = h->rank0;
    a.count = (int64_t)count;
    a.slice = slice;
    a.buf = buf;
    a.error = st->d_error;
    st->calls += 1;
    a.target = st->calls * (unsigned)kP2PCtas;
    const int parity = (int)(st->calls & 1u);
    for (int r = 0; r < W; ++r) {
        unsigned char* base = static_cast<unsigned char*>(st->peer[r]);
        a.flag_a[r] = reinterpret_cast<unsigned*>(base);
        a.flag_b[r] = reinterpret_cast<unsigned*>(base + 2048);
        double* areas = reinterpret_cast<double*>(base + 4096);
        a.inbox[r] = areas + (size_t)parity * st->cap;
        a.result[r] = areas + 2 * st->cap;
    }
    p2p_allreduce_kernel<<<kP2PCtas, kP2PThreads, 0, s.stream>>>(a);
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

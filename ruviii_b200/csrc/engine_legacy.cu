// The four generic matrix primitives RcppExports.R:12-26 declares beside fastResidop: matrixMult(A, B), matSubtraction(A, B),
// matTranspose(A), matInverse(A).  The reference never implemented or called them (no src/, no caller in R/); they are
// exported so that the package's declared native boundary is complete.  None of them is on the RUV-III hot path: the
// product is a plain library GEMM (cuBLAS Dgemm), the inverse LAPACK's dgesv route (cuSOLVER getrf + getrs against the
// identity, what R's solve(A) does), subtraction and transposition are trivially bandwidth-bound kernels.
// All matrices are column-major (R) host buffers; everything runs on the first GPU of the handle.
#include "engine_internal.cuh"

using namespace ruviii;

namespace {

__global__ void __launch_bounds__(256) subtract_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                       double* __restrict__ out, int64_t count) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < count; i += (int64_t)gridDim.x * 256) out[i] = a[i] - b[i];
}

__global__ void __launch_bounds__(256) identity_kernel(double* __restrict__ a, int n) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < (int64_t)n * n; i += (int64_t)gridDim.x * 256)
        a[i] = (i / n == i % n) ? 1.0 : 0.0;
}

struct Tmp {
    std::vector<DevBuf<double>*> bufs;
    ~Tmp() { for (auto b : bufs) b->release(); }
};

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int ruviii_mat_mult(ruviii_handle* h, const double* A, int32_t m, int32_t k, const double* B, int32_t n, double* out) {
    if (!h) return RUVIII_ERR_INVALID;
    return guarded(h, [&] {
        if (!A || !B || !out || m <= 0 || k <= 0 || n <= 0) throw Error(RUVIII_ERR_INVALID, "matrixMult: bad arguments");
        Shard& s = h->shards[0];
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        if (!h->blas && cublasCreate(&h->blas) != CUBLAS_STATUS_SUCCESS) throw Error(RUVIII_ERR_CUDA, "cublasCreate failed");
        DevBuf<double> dA, dB, dC;
        Tmp tmp{{&dA, &dB, &dC}};
        dA.alloc((size_t)m * k, false);
        dB.alloc((size_t)k * n, false);
        dC.alloc((size_t)m * n, false);
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(dA.p, A, (size_t)m * k * 8, cudaMemcpyHostToDevice, s.stream));
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(dB.p, B, (size_t)k * n * 8, cudaMemcpyHostToDevice, s.stream));
        const double one = 1.0, zero = 0.0;
        cublasSetStream(h->blas, s.stream);
        if (cublasDgemm(h->blas, CUBLAS_OP_N, CUBLAS_OP_N, m, n, k, &one, dA.p, m, dB.p, k, &zero, dC.p, m) != CUBLAS_STATUS_SUCCESS)
            throw Error(RUVIII_ERR_CUDA, "cublasDgemm (matrixMult) failed");
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(out, dC.p, (size_t)m * n * 8, cudaMemcpyDeviceToHost, s.stream));
        RUVIII_CUDA_CHECK(cudaStreamSynchronize(s.stream));
    });
}

int ruviii_mat_subtract(ruviii_handle* h, const double* A, const double* B, int64_t count, double* out) {
    if (!h) return RUVIII_ERR_INVALID;
    return guarded(h, [&] {
        if (!A || !B || !out || count <= 0) throw Error(RUVIII_ERR_INVALID, "matSubtraction: bad arguments");
        Shard& s = h->shards[0];
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        DevBuf<double> dA, dB;
        Tmp tmp{{&dA, &dB}};
        dA.alloc((size_t)count, false);
        dB.alloc((size_t)count, false);
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(dA.p, A, (size_t)count * 8, cudaMemcpyHostToDevice, s.stream));
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(dB.p, B, (size_t)count * 8, cudaMemcpyHostToDevice, s.stream));
        subtract_kernel<<<(int)std::min<int64_t>((count + 255) / 256, 8 * h->n_sms), 256, 0, s.stream>>>(dA.p, dB.p, dA.p, count);
        RUVIII_LAUNCH_CHECK();
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(out, dA.p, (size_t)count * 8, cudaMemcpyDeviceToHost, s.stream));
        RUVIII_CUDA_CHECK(cudaStreamSynchronize(s.stream));
    });
}

int ruviii_mat_transpose(ruviii_handle* h, const double* A, int32_t m, int32_t n, double* out) {
    if (!h) return RUVIII_ERR_INVALID;
    return guarded(h, [&] {
        if (!A || !out || m <= 0 || n <= 0) throw Error(RUVIII_ERR_INVALID, "matTranspose: bad arguments");
        Shard& s = h->shards[0];
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        DevBuf<double> dA, dT;
        Tmp tmp{{&dA, &dT}};
        dA.alloc((size_t)m * n, false);
        dT.alloc((size_t)m * n, false);
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(dA.p, A, (size_t)m * n * 8, cudaMemcpyHostToDevice, s.stream));
        // column-major m x n  ==  row-major n x m (pitch m); its transpose, row-major m x n (pitch n)  ==  column-major n x m
        launch_transpose(dA.p, m, n, m, dT.p, n, s.stream);
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(out, dT.p, (size_t)m * n * 8, cudaMemcpyDeviceToHost, s.stream));
        RUVIII_CUDA_CHECK(cudaStreamSynchronize(s.stream));
    });
}

int ruviii_mat_inverse(ruviii_handle* h, const double* A, int32_t n, double* out) {
    if (!h) return RUVIII_ERR_INVALID;
    return guarded(h, [&] {
        if (!A || !out || n <= 0) throw Error(RUVIII_ERR_INVALID, "matInverse: bad arguments");
        Shard& s = h->shards[0];
        RUVIII_CUDA_CHECK(cudaSetDevice(s.device));
        cusolverDnHandle_t solver = nullptr;
        RUVIII_SOLVER_CHECK(cusolverDnCreate(&solver));
        struct Guard { cusolverDnHandle_t h; ~Guard() { cusolverDnDestroy(h); } } guard{solver};
        RUVIII_SOLVER_CHECK(cusolverDnSetStream(solver, s.stream));
        DevBuf<double> dA, dI, dW;
        DevBuf<int> piv;
        Tmp tmp{{&dA, &dI, &dW}};
        dA.alloc((size_t)n * n, false);
        dI.alloc((size_t)n * n, false);
        piv.alloc((size_t)n + 4);
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(dA.p, A, (size_t)n * n * 8, cudaMemcpyHostToDevice, s.stream));
        identity_kernel<<<(int)std::min<int64_t>(((int64_t)n * n + 255) / 256, 8 * h->n_sms), 256, 0, s.stream>>>(dI.p, n);
        RUVIII_LAUNCH_CHECK();
        int lwork = 0;
        RUVIII_SOLVER_CHECK(cusolverDnDgetrf_bufferSize(solver, n, n, dA.p, n, &lwork));
        dW.alloc((size_t)std::max(lwork, 1), false);
        int* info = piv.p + n;
        RUVIII_SOLVER_CHECK(cusolverDnDgetrf(solver, n, n, dA.p, n, dW.p, piv.p, info));
        int h_info = 0;
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(&h_info, info, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
        RUVIII_CUDA_CHECK(cudaStreamSynchronize(s.stream));
        if (h_info != 0) {
            piv.release();
            throw Error(RUVIII_ERR_SINGULAR, "matInverse: the matrix is exactly singular: U[" + std::to_string(h_info) + "," +
                                                 std::to_string(h_info) + "] = 0 (R's solve() would fail)");
        }
        RUVIII_SOLVER_CHECK(cusolverDnDgetrs(solver, CUBLAS_OP_N, n, n, dA.p, n, piv.p, dI.p, n, info));
        RUVIII_CUDA_CHECK(cudaMemcpyAsync(out, dI.p, (size_t)n * n * 8, cudaMemcpyDeviceToHost, s.stream));
        RUVIII_CUDA_CHECK(cudaStreamSynchronize(s.stream));
        piv.release();
    });
}

}  // extern "C"
#pragma GCC visibility pop

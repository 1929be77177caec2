// Times the sections of one Jacobi round of the persistent eigensolver (the real jacobi_phase, compiled from the product
// source with clock stamps).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ruviii_b200/csrc -o
// scripts/jacobi_probe scripts/jacobi_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
__device__ long long g_jacc[8];
#define RUVIII_JSTAMP_DECL long long jacc_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long jlast_ = clock64();
#define RUVIII_JSTAMP(k) do { const long long t_ = clock64(); jacc_[k] += t_ - jlast_; jlast_ = t_; } while (0)
#define RUVIII_JSTAMP_FLUSH if (threadIdx.x == 0) { for (int i_ = 0; i_ < 8; ++i_) g_jacc[i_] = jacc_[i_]; }
#include "../ruviii_b200/csrc/eigen_persist.cu"

namespace ruviii { namespace {
__global__ void __launch_bounds__(PT) jacobi_probe_kernel(const double* H, double* th, double* V, long long* out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    for (int e = threadIdx.x; e < PB * PB; e += PT) sm.A[e / PB][e % PB] = H[e];
    __syncthreads();
    const long long t0 = clock64();
    const int sweeps = jacobi_phase(sm);
    const long long t1 = clock64();
    for (int e = threadIdx.x; e < PB * PB; e += PT) V[e] = sm.V[e / PB][e % PB];
    if (threadIdx.x < PB) th[threadIdx.x] = sm.th[threadIdx.x];
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = sweeps; for (int i = 0; i < 8; ++i) out[2 + i] = g_jacc[i]; }
}
}}

int main() {
    using namespace ruviii;
    std::vector<double> H(PB * PB);
    srand(7);
    for (int i = 0; i < PB; ++i)
        for (int j = 0; j <= i; ++j) {
            double v = (rand() / (double)RAND_MAX - 0.5) * 2.0;
            if (i == j) v = i < 10 ? 1000.0 / (1 << i) + v : 1.0 + 0.2 * v;
            H[i * PB + j] = H[j * PB + i] = v;
        }
    double *dH, *dth, *dV; long long* dout;
    cudaMalloc(&dH, PB * PB * 8); cudaMalloc(&dth, PB * 8); cudaMalloc(&dV, PB * PB * 8); cudaMallocManaged(&dout, 16 * 8);
    cudaMemcpy(dH, H.data(), PB * PB * 8, cudaMemcpyHostToDevice);
    const size_t smem = sizeof(Smem) + 64;
    cudaFuncSetAttribute(jacobi_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 3; ++rep) {
        jacobi_probe_kernel<<<1, PT, smem>>>(dH, dth, dV, dout);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    }
    std::vector<double> th(PB), V(PB * PB);
    cudaMemcpy(th.data(), dth, PB * 8, cudaMemcpyDeviceToHost); cudaMemcpy(V.data(), dV, PB * PB * 8, cudaMemcpyDeviceToHost);
    // residual || H V - V diag(th) ||_max
    double worst = 0.0;
    for (int j = 0; j < PB; ++j)
        for (int i = 0; i < PB; ++i) {
            double s = 0.0;
            for (int q = 0; q < PB; ++q) s += H[i * PB + q] * V[q * PB + j];
            worst = fmax(worst, fabs(s - th[j] * V[i * PB + j]));
        }
    const long long rounds = dout[1] * (PB - 1);
    printf("sweeps %lld rounds %lld total %lld cycles (%.0f per round), residual %.2e\n", dout[1], rounds, dout[0], (double)dout[0] / rounds, worst);
    const char* nm[] = {"barrier->top (index math)", "loads + n2", "two rsqrt + select", "shuffles", "rotate + stores", "__syncthreads"};
    for (int i = 0; i < 6; ++i) printf("  %-28s %8.1f cycles per round\n", nm[i], (double)dout[2 + i] / rounds);
    return 0;
}

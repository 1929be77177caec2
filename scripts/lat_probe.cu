// Dependent-chain latencies (cycles per op) of the instructions the latency-bound phases of the persistent eigensolver are
// made of (Cholesky, Jacobi: one CTA of 256 threads, FP64).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o
// scripts/lat_probe scripts/lat_probe.cu ; run on the GPU box: ./scripts/lat_probe
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 2048;

__global__ void probe(double seed, float fseed, long long* out, double* sink) {
    __shared__ double sh[512];
    __shared__ int chase[256];
    const int tid = threadIdx.x;
    sh[tid] = seed + tid; sh[256 + tid] = seed - tid;
    chase[tid] = (tid + 32) & 255;
    __syncthreads();
    long long t0, t1;
    double x = seed, y = seed * 0.5;
    float f = fseed;
    int k = 0;
    // DFMA
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, 1.0000001, y);
    t1 = clock64(); if (tid == 0) out[k] = t1 - t0; ++k;
    // DMUL
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x * 1.0000001;
    t1 = clock64(); if (tid == 0) out[k] = t1 - t0; ++k;
    // DADD
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x + y;
    t1 = clock64(); if (tid == 0) out[k] = t1 - t0; ++k;
    // rsqrt(double)
    x = fabs(x) + 1.0;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) x = rsqrt(x) + 1.5;
    t1 = clock64(); if (tid == 0) out[k] = t1 - t0; ++k;
    // sqrt(double)
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) x = sqrt(x) + 1.5;
    t1 = clock64(); if (tid == 0) out[k] = t1 - t0; ++k;
    // 1/x (double)
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) x = 1.0 / x + 1.5;
    t1 = clock64(); if (tid == 0) out[k] = t1 - t0; ++k;
    // FFMA
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) f = fmaf(f, 1.0001f, fseed);
    t1 = clock64(); if (tid == 0) out[k] = t1 - t0; ++k;
    // rsqrtf
    f = fabsf(f) + 1.0f;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) f = rsqrtf(f) + 1.5f;
    t1 = clock64(); if (tid == 0) out[k] = t1 - t0; ++k;
    // double -> float -> double round trip
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) x = (double)((float)x) + 1.5;
    t1 = clock64(); if (tid == 0) out[k] = t1 - t0; ++k;
    // LDS pointer chase
    int p = tid;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) p = chase[p];
    t1 = clock64(); if (tid == 0) out[k] = t1 - t0; ++k;
    // generic load of shared memory, pointer chase
    {
        const int* g = chase;
        const int* volatile gv = g;                // hide the address space from the compiler
        const int* gp = gv;
        t0 = clock64();
#pragma unroll 8
        for (int i = 0; i < N; ++i) p = gp[p];
        t1 = clock64(); if (tid == 0) out[k] = t1 - t0; ++k;
    }
    // LDS.64 -> DFMA -> STS.64 -> LDS.64 (same address) round trip
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) { double v = sh[tid]; v = fma(v, 1.0000001, y); sh[tid] = v; }
    t1 = clock64(); if (tid == 0) out[k] = t1 - t0; ++k;
    // SHFL of a double (two 32-bit shuffles) chain
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (tid + 1) & 31) + 1.0;
    t1 = clock64(); if (tid == 0) out[k] = t1 - t0; ++k;
    // __syncthreads
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) __syncthreads();
    t1 = clock64(); if (tid == 0) out[k] = t1 - t0; ++k;
    // __syncwarp
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) __syncwarp();
    t1 = clock64(); if (tid == 0) out[k] = t1 - t0; ++k;
    // local-memory load (dynamically indexed register array)
    {
        double arr[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) arr[i] = seed + i;
        int q = tid & 7;
        t0 = clock64();
#pragma unroll 8
        for (int i = 0; i < N; ++i) { arr[q] += 1.0; q = (int)arr[(q + 1) & 7] & 7; }
        t1 = clock64(); if (tid == 0) out[k] = t1 - t0; ++k;
        x += arr[q];
    }
    sink[tid] = x + f + p;
}

int main() {
    const char* names[] = {"DFMA", "DMUL", "DADD", "rsqrt(double)+DADD", "sqrt(double)+DADD", "1/x(double)+DADD", "FFMA", "rsqrtf+FADD",
                           "F2F f64->f32->f64 + DADD", "LDS chase", "generic LD (shared) chase", "LDS->DFMA->STS round trip",
                           "SHFL double + DADD", "__syncthreads", "__syncwarp", "local array RMW + dependent read"};
    long long* out; double* sink;
    cudaMallocManaged(&out, 64 * sizeof(long long));
    cudaMalloc(&sink, 1024 * sizeof(double));
    for (int threads : {32, 256}) {
        for (int rep = 0; rep < 2; ++rep) {
            probe<<<1, threads>>>(1.25, 0.75f, out, sink);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 1; }
        }
        printf("threads per CTA = %d (cycles per dependent op, thread 0)\n", threads);
        for (int i = 0; i < 16; ++i) printf("  %-36s %8.1f\n", names[i], (double)out[i] / N);
    }
    return 0;
}

// Shared helpers for the RUV-III engine kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace ruviii {

constexpr int kMaxK = 32;          // largest k (= rows of alpha used by the update) the fused kernels carry in registers
constexpr int kRowAlign = 16;      // device row pitch of every genes-contiguous matrix, in doubles (128 B)

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define RUVIII_CUDA_CHECK(expr)                                                                   \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            throw ::ruviii::Error(2, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +  \
                                         __FILE__ + ":" + std::to_string(__LINE__) + ")");        \
    } while (0)

#define RUVIII_LAUNCH_CHECK() RUVIII_CUDA_CHECK(cudaGetLastError())

inline int64_t round_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// padded k (register-array extent) for a given kmax
inline int pad_k(int k) { return k <= 4 ? 4 : k <= 8 ? 8 : k <= 16 ? 16 : 32; }

#ifdef __CUDACC__
// streaming 16-byte accesses: data touched exactly once must not displace the reusable working set in L1
__device__ __forceinline__ double2 ld_stream(const double* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(double* p, double2 v) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

}  // namespace ruviii

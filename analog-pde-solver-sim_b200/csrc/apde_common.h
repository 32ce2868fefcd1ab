// Shared host/device helpers for libapde.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "apde.h"

namespace apde {

// ---- thread-local error message ------------------------------------------------------------
void set_error(const char *fmt, ...);
const char *get_error();

#define APDE_CUDA(call)                                                                       \
    do {                                                                                      \
        cudaError_t err__ = (call);                                                           \
        if (err__ != cudaSuccess) {                                                           \
            ::apde::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                   \
                              cudaGetErrorString(err__));                                     \
            return (err__ == cudaErrorMemoryAllocation) ? APDE_ENOMEM : APDE_ECUDA;           \
        }                                                                                     \
    } while (0)

#define APDE_TRY(call)                                                                        \
    do {                                                                                      \
        int rc__ = (call);                                                                    \
        if (rc__ != APDE_OK) return rc__;                                                     \
    } while (0)

// Fails (no fallback) unless the current device is compute capability 10.x.
int require_sm100(cudaDeviceProp *prop_out = nullptr);

// ---- RAII device buffer --------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    int alloc(size_t n) {
        release();
        if (n == 0) n = 16;
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMalloc(%zu bytes) failed: %s", n, cudaGetErrorString(e));
            return APDE_ENOMEM;
        }
        bytes = n;
        return APDE_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};

struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    ~EventPair() {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
    int init() {
        APDE_CUDA(cudaEventCreate(&a));
        APDE_CUDA(cudaEventCreate(&b));
        return APDE_OK;
    }
};

// ---- device-side bit tricks for max-reductions of non-negative reals -------------------------
#ifdef __CUDACC__
// fmax-style: NaN never wins (the reference's strict '>' ignores NaN changes, solver.py:277)
__device__ __forceinline__ double max_ignore_nan(double a, double b) { return (b > a) ? b : a; }
__device__ __forceinline__ float max_ignore_nan(float a, float b) { return (b > a) ? b : a; }

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max_ignore_nan(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max_ignore_nan(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Non-negative IEEE values order like their unsigned bit patterns.
__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v) {
    atomicMax(reinterpret_cast<unsigned long long *>(addr),
              static_cast<unsigned long long>(__double_as_longlong(v)));
}
__device__ __forceinline__ void atomic_max_nonneg(float *addr, float v) {
    atomicMax(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}
#endif

}  // namespace apde

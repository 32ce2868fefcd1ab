// Inline-PTX wrappers used by the sm_100a kernels: mbarrier, bulk-async copies (TMA engine, SASS UBLKCP),
// bulk L2 prefetch, shared-memory vector access, named barriers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace apde {
namespace ptx {

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// ---- mbarrier (shared::cta) ------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// ---- bulk-async copy global -> shared (1-D, TMA engine), completion on an mbarrier ---------------
__device__ __forceinline__ void bulk_load(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// pull `bytes` (multiple of 16, 16-byte aligned) from HBM into L2 with one instruction
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// ---- 16-byte shared-memory vectors typed as the element type --------------------------------------
__device__ __forceinline__ void lds16(double *w, unsigned addr) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(w[0]), "=d"(w[1]) : "r"(addr));
}
__device__ __forceinline__ void lds16(float *w, unsigned addr) {
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(w[0]), "=f"(w[1]), "=f"(w[2]), "=f"(w[3]) : "r"(addr));
}
__device__ __forceinline__ void sts16(unsigned addr, const double *w) {
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(addr), "d"(w[0]), "d"(w[1]) : "memory");
}
__device__ __forceinline__ void sts16(unsigned addr, const float *w) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(w[0]), "f"(w[1]), "f"(w[2]), "f"(w[3]) : "memory");
}

// ---- 16-byte global vectors ----------------------------------------------------------------------
__device__ __forceinline__ void ldg16(double *w, const double *p) {
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(w[0]), "=d"(w[1]) : "l"(p));
}
__device__ __forceinline__ void ldg16(float *w, const float *p) {
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(w[0]), "=f"(w[1]), "=f"(w[2]), "=f"(w[3]) : "l"(p));
}
__device__ __forceinline__ void stg16(double *p, const double *w) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(w[0]), "d"(w[1]) : "memory");
}
__device__ __forceinline__ void stg16(float *p, const float *w) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(w[0]), "f"(w[1]), "f"(w[2]), "f"(w[3])
                 : "memory");
}

}  // namespace ptx
}  // namespace apde

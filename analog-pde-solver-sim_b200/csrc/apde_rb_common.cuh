// Shared device code of the red-black kernels: arithmetic traits, node update, strip geometry.
#pragma once
#include "apde_plan.h"

namespace apde {

// ------------------------------------------------------------------------------------------
// arithmetic traits
// ------------------------------------------------------------------------------------------
template <typename T> struct Ar;
template <> struct Ar<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
    static __device__ __forceinline__ double abs(double a) { return fabs(a); }
};
template <> struct Ar<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
};

template <typename T, bool NOISE, bool SRC>
__device__ __forceinline__ T rb_relax(T un, T us, T uw, T ue, T gn, T gs, T gw, T ge, T sv) {
    T acc;
    if (NOISE) {
        acc = Ar<T>::mul(gn, un);
        acc = Ar<T>::fma(gs, us, acc);
        acc = Ar<T>::fma(gw, uw, acc);
        acc = Ar<T>::fma(ge, ue, acc);
    } else {
        acc = Ar<T>::add(un, us);
        acc = Ar<T>::add(acc, uw);
        acc = Ar<T>::add(acc, ue);
    }
    if (SRC) acc = Ar<T>::sub(acc, sv);
    return Ar<T>::mul(acc, (T)0.25);
}

struct Geo {
    long long pitch, padl;
    long long lrows, cols, grows;
    long long row0;        // global row of local row 0
    long long own_lo, own_hi;  // owned local rows [own_lo, own_hi)
};


inline Geo make_geo(const apde_plan *p) {
    Geo g;
    g.pitch = p->pitch;
    g.padl = p->padl;
    g.lrows = p->lrows;
    g.cols = p->d.cols;
    g.grows = p->d.grows;
    g.row0 = p->row0;
    g.own_lo = p->ghost_top;
    g.own_hi = p->lrows - p->ghost_bot;
    return g;
}

// implemented in apde_rb_stream_f64.cu / apde_rb_stream_f32.cu
int run_stream_f64(apde_plan *p, int64_t n_sweeps, int64_t sweep_base, int phase, cudaStream_t st);
int run_stream_f32(apde_plan *p, int64_t n_sweeps, int64_t sweep_base, int phase, cudaStream_t st);

}  // namespace apde

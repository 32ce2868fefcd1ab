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
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};
template <> struct Ar<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};

// The red-black node update, split so that the operand that arrives last in the streaming kernel's
// stage pipeline (the south neighbour, produced by the previous stage of the same iteration) enters
// through ONE fused multiply-add:
//     mismatch   : t = gN*uN; t = fma(gW,uW,t); t = fma(gE,uE,t); t = t - src; tq = t*0.25;
//                  new = fma(gS*0.25, uS, tq)
//     no mismatch: tq = (((uN + uW) + uE) - src) * 0.25;   new = fma(uS, 0.25, tq)
// Mathematically the reference's (n+s+w+e-src)/4 (solver.py:275 / :350-356); the association is the
// throughput path's own (its iterates are compared with the red-black oracle, which uses the same tree).
template <typename T, bool NOISE, bool SRC>
__device__ __forceinline__ T rb_partial(T un, T uw, T ue, T gn, T gw, T ge, T sv) {
    T t;
    if (NOISE) {
        t = Ar<T>::mul(gn, un);
        t = Ar<T>::fma(gw, uw, t);
        t = Ar<T>::fma(ge, ue, t);
    } else {
        t = Ar<T>::add(un, uw);
        t = Ar<T>::add(t, ue);
    }
    if (SRC) t = Ar<T>::sub(t, sv);
    return Ar<T>::mul(t, (T)0.25);
}
template <typename T, bool NOISE>
__device__ __forceinline__ T rb_south_weight(T gs) {
    return NOISE ? Ar<T>::mul(gs, (T)0.25) : (T)0.25;
}
template <typename T>
__device__ __forceinline__ T rb_finish(T tq, T gsq, T us) {
    return Ar<T>::fma(gsq, us, tq);
}
template <typename T, bool NOISE, bool SRC>
__device__ __forceinline__ T rb_relax(T un, T us, T uw, T ue, T gn, T gs, T gw, T ge, T sv) {
    return rb_finish<T>(rb_partial<T, NOISE, SRC>(un, uw, ue, gn, gw, ge, sv), rb_south_weight<T, NOISE>(gs), us);
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
// one pass over a band of local rows with explicit buffers (the streamed one-shot solve)
int run_stream_band_f64(apde_plan *p, int td, int64_t resid_slot, int src_buf, int64_t row_lo, int64_t row_hi, int64_t rbh, cudaStream_t st);
int run_stream_band_f32(apde_plan *p, int td, int64_t resid_slot, int src_buf, int64_t row_lo, int64_t row_hi, int64_t rbh, cudaStream_t st);
// implemented in apde_rb_pipe_f64.cu / apde_rb_pipe_f32.cu
int run_pipe_f64(apde_plan *p, int64_t n_sweeps, int64_t sweep_base, int phase, cudaStream_t st);
int run_pipe_f32(apde_plan *p, int64_t n_sweeps, int64_t sweep_base, int phase, cudaStream_t st);

}  // namespace apde

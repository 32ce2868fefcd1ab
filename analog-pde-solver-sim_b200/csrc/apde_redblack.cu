// apde_redblack.cu -- red-black Gauss-Seidel throughput path on sm_100a (fp64 / fp32).
//
// Same fixed point as the reference sweeps (analog_pde_solver/solver.py:263-287, :344-368) but
// relaxed colour by colour ((i+j) even first), so every node of a colour is independent.
// Per node, with conductances g = 1/noise precomputed once (recip kernel):
//     acc = gN*uN; acc = fma(gS,uS,acc); acc = fma(gW,uW,acc); acc = fma(gE,uE,acc)     (mismatch)
//     acc = ((uN+uS)+uW)+uE                                                           (no mismatch; solver.py:350-356 tree)
//     acc = acc - src ; new = acc*0.25 ; change = |new-old|
// All arithmetic uses explicit round-to-nearest intrinsics so that every kernel variant and the
// CPU checker used by the tests (its red-black functions) produce identical bits.
//
// Strip layout (see apde_plan.h): lrows x pitch elements, `padl` zero columns left of column 0,
// ghost rows above/below when the strip has neighbours.  Local row 0 and lrows-1 are never
// updated (they are the global Dirichlet ring, or the outermost ghost row).
#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "apde_plan.h"
#include "apde_rb_common.cuh"

namespace apde {

// ==========================================================================================
// v0: one launch per colour, in place.  Baseline / cross-check variant (40 B/LUP at fp64).
// ==========================================================================================
template <typename T, bool NOISE, bool SRC>
__global__ void __launch_bounds__(256) rb_colour_kernel(T *__restrict__ u, const T *__restrict__ src,
                                                        const T *__restrict__ gh,
                                                        const T *__restrict__ gv, Geo g, int colour,
                                                        T *resid_slot, const int *stop) {
    if (stop && *stop) return;
    const long long hw = (g.cols - 1) / 2;
    const long long jj = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    T mx = (T)0;
    for (long long li = 1 + blockIdx.y; li <= g.lrows - 2; li += gridDim.y) {
        const long long gi = g.row0 + li;
        if (gi <= 0 || gi >= g.grows - 1) continue;  // ring rows never move
        const long long j = 1 + 2 * jj + ((gi + 1 + colour) & 1);
        if (jj < hw && j <= g.cols - 2) {
            const long long c = li * g.pitch + g.padl + j;
            T gn = (T)1, gs = (T)1, gw = (T)1, ge = (T)1, sv = (T)0;
            if (NOISE) {
                gn = gv[c - g.pitch];
                gs = gv[c];
                gw = gh[c - 1];
                ge = gh[c];
            }
            if (SRC) sv = src[c];
            const T old = u[c];
            const T nw = rb_relax<T, NOISE, SRC>(u[c - g.pitch], u[c + g.pitch], u[c - 1], u[c + 1],
                                                 gn, gs, gw, ge, sv);
            u[c] = nw;
            if (li >= g.own_lo && li < g.own_hi) mx = max_ignore_nan(mx, Ar<T>::abs(Ar<T>::sub(nw, old)));
        }
    }
    __shared__ T red[8];
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
        T v = threadIdx.x < 8 ? red[threadIdx.x] : (T)0;
        v = warp_max(v);
        if (threadIdx.x == 0 && v > (T)0) atomic_max_nonneg(resid_slot, v);
    }
}

// ==========================================================================================
// General per-colour kernel: over-relaxation and the normalised ("KCL") update (not in the
// reference, whose update divides by the constant 4 -- solver.py:269-275 -- and therefore diverges with
// mismatch once the grid is large, SURVEY finding 2).  The CPU checker of the tests states the same tree (RBX_BODY).
//   kcl   : scale = 1 / (((gN + gS) + gW) + gE) instead of 0.25
//   omega : new = fma(omega, gs - old, old)
// In place, one launch per colour; memory-bound on u + operands, no temporal blocking: these variants buy
// orders of magnitude in sweep COUNT, not in sweep rate.
// ==========================================================================================
template <typename T, bool NOISE, bool SRC>
__global__ void __launch_bounds__(256) rb_general_kernel(T *__restrict__ u, const T *__restrict__ src,
                                                         const T *__restrict__ gh, const T *__restrict__ gv, Geo g,
                                                         int colour, T *resid_slot, const int *stop, T omega, int kcl) {
    if (stop && *stop) return;
    const long long hw = (g.cols - 1) / 2;
    const long long jj = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const bool sor = omega != (T)1;
    T mx = (T)0;
    for (long long li = 1 + blockIdx.y; li <= g.lrows - 2; li += gridDim.y) {
        const long long gi = g.row0 + li;
        if (gi <= 0 || gi >= g.grows - 1) continue;  // ring rows never move
        const long long j = 1 + 2 * jj + ((gi + 1 + colour) & 1);
        if (jj < hw && j <= g.cols - 2) {
            const long long c = li * g.pitch + g.padl + j;
            const T un = u[c - g.pitch], us = u[c + g.pitch], uw = u[c - 1], ue = u[c + 1], old = u[c];
            T tq, gsq, scale = (T)0.25;
            if (NOISE) {
                const T gn = gv[c - g.pitch], gs = gv[c], gw = gh[c - 1], ge = gh[c];
                if (kcl) scale = Ar<T>::div((T)1, Ar<T>::add(Ar<T>::add(Ar<T>::add(gn, gs), gw), ge));
                tq = Ar<T>::mul(gn, un);
                tq = Ar<T>::fma(gw, uw, tq);
                tq = Ar<T>::fma(ge, ue, tq);
                gsq = Ar<T>::mul(gs, scale);
            } else {
                tq = Ar<T>::add(un, uw);
                tq = Ar<T>::add(tq, ue);
                gsq = scale;
            }
            if (SRC) tq = Ar<T>::sub(tq, src[c]);
            tq = Ar<T>::mul(tq, scale);
            T nw = Ar<T>::fma(gsq, us, tq);
            if (sor) nw = Ar<T>::fma(omega, Ar<T>::sub(nw, old), old);
            u[c] = nw;
            if (li >= g.own_lo && li < g.own_hi) mx = max_ignore_nan(mx, Ar<T>::abs(Ar<T>::sub(nw, old)));
        }
    }
    __shared__ T red[8];
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
        T v = threadIdx.x < 8 ? red[threadIdx.x] : (T)0;
        v = warp_max(v);
        if (threadIdx.x == 0 && v > (T)0) atomic_max_nonneg(resid_slot, v);
    }
}

// ------------------------------------------------------------------------------------------
// helpers: conversion / reciprocal / fill / checksum
// ------------------------------------------------------------------------------------------
// host float64 row-major (global) -> strip layout, plan dtype; recip => store 1/x (K5)
template <typename T, bool RECIP>
__global__ void pack_kernel(const double *__restrict__ in, long long in_rows, long long in_cols,
                            long long in_row0 /*global row of in[0]*/, T *__restrict__ out, Geo g) {
    const long long n = g.lrows * in_cols;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x) {
        const long long li = k / in_cols, j = k - li * in_cols;
        const long long r = g.row0 + li - in_row0;
        if (r < 0 || r >= in_rows) continue;
        double v = in[r * in_cols + j];
        if (RECIP) v = 1.0 / v;
        out[li * g.pitch + g.padl + j] = (T)v;
    }
}
template <typename T>
__global__ void unpack_kernel(const T *__restrict__ in, Geo g, double *__restrict__ out,
                              long long out_row0, long long out_rows) {
    const long long n = out_rows * g.cols;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x) {
        const long long orow = k / g.cols, j = k - orow * g.cols;
        const long long li = out_row0 + orow - g.row0;
        if (li < g.own_lo || li >= g.own_hi) continue;
        out[k] = (double)in[li * g.pitch + g.padl + j];
    }
}

// Philox4x32-10 (Salmon et al.), counter = link id, key = seed
__device__ __forceinline__ void philox4x32_10(unsigned int c[4], unsigned int k0, unsigned int k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const unsigned int n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ double philox_normal(unsigned long long id, unsigned long long seed) {
    unsigned int c[4] = {(unsigned int)id, (unsigned int)(id >> 32), 0x41504445u /*'APDE'*/, 0u};
    philox4x32_10(c, (unsigned int)seed, (unsigned int)(seed >> 32));
    const double u1 = ((double)c[0] * 4294967296.0 + (double)c[1] + 0.5) * (1.0 / 18446744073709551616.0);
    const double u2 = ((double)c[2] * 4294967296.0 + (double)c[3] + 0.5) * (1.0 / 18446744073709551616.0);
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

template <typename T>
__global__ void fill_kernel(T *__restrict__ u, T *__restrict__ src, T *__restrict__ gh,
                            T *__restrict__ gv, Geo g, apde_fill_desc f) {
    const long long n = g.lrows * g.cols;
    const double PI = 3.14159265358979323846;
    const long long m = (g.grows > g.cols ? g.grows : g.cols) - 1;
    const double h = 1.0 / (double)m;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x) {
        const long long li = k / g.cols, j = k - li * g.cols;
        const long long gi = g.row0 + li;
        const long long c = li * g.pitch + g.padl + j;
        u[c] = (f.top_hot && gi == 0) ? (T)1 : (T)0;
        if (src) {
            double s = 0.0;
            if (f.kind == 1) {
                const double x = (double)gi / (double)(g.grows - 1), y = (double)j / (double)(g.cols - 1);
                s = (h * h) * (-2.0 * PI * PI * sin(PI * x) * sin(PI * y));
            }
            src[c] = (T)s;
        }
        if (gh) {
            // gh[gi][j]: link (gi,j)-(gi,j+1), j < cols-1 ; gv[gi][j]: link (gi,j)-(gi+1,j), gi < grows-1
            const unsigned long long idh = (unsigned long long)gi * (unsigned long long)g.cols + (unsigned long long)j;
            const unsigned long long idv = idh + (unsigned long long)g.grows * (unsigned long long)g.cols;
            double vh = 1.0, vv = 1.0;
            if (f.noise_level > 0.0) {
                vh = 1.0 / (1.0 + f.noise_level * philox_normal(idh, f.seed));
                vv = 1.0 / (1.0 + f.noise_level * philox_normal(idv, f.seed));
            }
            gh[c] = (j < g.cols - 1) ? (T)vh : (T)0;
            gv[c] = (gi < g.grows - 1) ? (T)vv : (T)0;
        }
    }
}

template <typename T>
__global__ void checksum_kernel(const T *__restrict__ u, Geo g, double *out /*[0]=sum,[1]=absmax,[2]=bit sum*/) {
    const long long n = (g.own_hi - g.own_lo) * g.cols;
    double s = 0.0, m = 0.0;
    unsigned long long bits = 0;  // wrap-around sum of the bit patterns: order independent, hence bitwise reproducible
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x) {
        const long long r = k / g.cols, j = k - r * g.cols;
        const T e = u[(g.own_lo + r) * g.pitch + g.padl + j];
        const double v = (double)e;
        s += v;
        m = max_ignore_nan(m, fabs(v));
        const unsigned long long b = sizeof(T) == 8 ? (unsigned long long)__double_as_longlong((double)e)
                                                    : (unsigned long long)__float_as_uint((float)e);
        bits += b * (unsigned long long)(2 * ((g.row0 + g.own_lo + r) * g.cols + j) + 1);  // position-weighted
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        bits += __shfl_xor_sync(0xffffffffu, bits, o);
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out, s);
        atomic_max_nonneg(out + 1, m);
        atomicAdd(reinterpret_cast<unsigned long long *>(out + 2), bits);
    }
}

// sum of squared voltage drops over the links this strip owns: horizontal links of owned rows and
// vertical links from an owned row to the row below it (owned or ghost) -- the two sums of
// EnergyModel.analog_energy_joules (solver.py:85-92 upstream), accumulated in fp64
template <typename T>
__global__ void energy_kernel(const T *__restrict__ u, Geo g, double *out /*[0]=sum dh^2,[1]=sum dv^2*/) {
    const long long n = (g.own_hi - g.own_lo) * g.cols;
    double sh = 0.0, sv = 0.0;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x) {
        const long long r = k / g.cols, j = k - r * g.cols;
        const long long li = g.own_lo + r, gi = g.row0 + li;
        const long long c = li * g.pitch + g.padl + j;
        const double v = (double)u[c];
        if (j + 1 < g.cols) {
            const double d = (double)u[c + 1] - v;
            sh += d * d;
        }
        if (gi + 1 < g.grows) {
            const double d = (double)u[c + g.pitch] - v;
            sv += d * d;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sh += __shfl_xor_sync(0xffffffffu, sh, o);
        sv += __shfl_xor_sync(0xffffffffu, sv, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out, sh);
        atomicAdd(out + 1, sv);
    }
}

// K6: the convergence decision, on the device.  After the sweeps of one check interval, one thread block looks at
// the interval's per-sweep maxima (solver.py:283: max_change < tol, strict) and, the first time one is below tol,
// publishes {stop = 1, intervals = this interval's ordinal} to a status block the host can read without a copy
// (mapped pinned memory).  Every sweep kernel launched afterwards sees the flag and returns at once, so the host may
// queue check intervals ahead without ever reading a residual back.
struct SolveStatus {
    int stop;        // 1 once an interval has converged
    int intervals;   // ordinal (1-based) of the interval that converged
};
template <typename T>
__global__ void stop_check_kernel(const T *__restrict__ resid, int n, double tol, int ordinal, int *dev_flag,
                                  SolveStatus *status) {
    if (*dev_flag) return;
    __shared__ int hit;
    if (threadIdx.x == 0) hit = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x)
        if ((double)resid[k] < tol) hit = 1;
    __syncthreads();
    if (threadIdx.x == 0 && hit) {
        *dev_flag = 1;  // what the sweep kernels poll (device memory: an L2 hit, not a PCIe read)
        status->intervals = ordinal;
        __threadfence_system();
        status->stop = 1;  // what the host polls
    }
}

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------
// Tuning / cross-check knobs from the environment (read at plan creation and whenever a cached plan is reused).
static void apply_env_knobs(apde_plan *p) {
    p->variant = default_rb_variant();
    // 16-byte vectors per lane per row, measured per variant at 16384^2 (profiles/r2_rb_sweep_async.txt): with conductances
    // one (the six operand segments per row of two would not fit the shared-memory rings); without, two for fp64
    // (Poisson 660 vs 554 GLUP/s) and for fp32 Laplace (1894 vs 1556), one for fp32 Poisson (1284 vs 1187)
    p->lane_vectors = p->d.has_noise ? 1 : (p->d.dtype_bytes == 8 || !p->d.has_source) ? 2 : 1;
    if (const char *v = getenv("APDE_LANE_VECTORS")) p->lane_vectors = atoi(v) == 2 ? 2 : 1;
    p->row_block = 256;
    p->row_block_forced = false;
    if (const char *v = getenv("APDE_ROW_BLOCK")) {
        p->row_block = std::max(8, atoi(v));
        p->row_block_forced = true;
    }
    p->edge_rows = 128;
    if (const char *v = getenv("APDE_EDGE_ROWS")) p->edge_rows = std::max(1, atoi(v));
}

int plan_create(apde_plan **out, const apde_plan_desc *d) {
    if (!out || !d) {
        set_error("plan_create: NULL argument");
        return APDE_EINVAL;
    }
    *out = nullptr;
    if ((d->dtype_bytes != 4 && d->dtype_bytes != 8) || d->grows < 1 || d->cols < 1 ||
        d->row_begin < 0 || d->row_end > d->grows || d->row_begin >= d->row_end ||
        d->temporal_depth < 1 || d->temporal_depth > 6) {
        set_error("plan_create: bad descriptor (dtype=%d grows=%lld cols=%lld rows=[%lld,%lld) depth=%d)",
                  d->dtype_bytes, (long long)d->grows, (long long)d->cols, (long long)d->row_begin,
                  (long long)d->row_end, d->temporal_depth);
        return APDE_EINVAL;
    }
    cudaDeviceProp prop;
    APDE_TRY(require_sm100(&prop));
    apde_plan *p = new (std::nothrow) apde_plan();
    if (!p) {
        set_error("plan_create: out of host memory");
        return APDE_ENOMEM;
    }
    p->d = *d;
    APDE_CUDA(cudaGetDevice(&p->device));
    p->sm_count = prop.multiProcessorCount;
    p->elem = d->dtype_bytes;
    p->halo = 2 * (int64_t)d->temporal_depth;
    p->ghost_top = d->row_begin > 0 ? p->halo : 0;
    p->ghost_bot = d->row_end < d->grows ? p->halo : 0;
    const int64_t owned = d->row_end - d->row_begin;
    if ((p->ghost_top || p->ghost_bot) && owned < p->halo) {
        set_error("plan_create: strip of %lld rows is thinner than its halo (%lld)", (long long)owned,
                  (long long)p->halo);
        delete p;
        return APDE_EINVAL;
    }
    if (d->row_begin - p->ghost_top < 0 || d->row_end + p->ghost_bot > d->grows) {
        set_error("plan_create: halo reaches outside the global grid");
        delete p;
        return APDE_EINVAL;
    }
    p->lrows = p->ghost_top + owned + p->ghost_bot;
    p->row0 = d->row_begin - p->ghost_top;
    p->padl = 64;  // 512 B (fp64) of zero columns: vector alignment + left tile overhang
    p->pitch = ((p->padl + d->cols + 576 + 63) / 64) * 64;  // right overhang: two of the widest warp strips
    apply_env_knobs(p);
    const size_t bytes = p->plane_elems() * (size_t)p->elem;
    int rc = APDE_OK;
    if ((rc = p->u[0].alloc(bytes)) || (rc = p->u[1].alloc(bytes)) ||
        (d->has_source && (rc = p->src.alloc(bytes))) ||
        (d->has_noise && ((rc = p->gh.alloc(bytes)) || (rc = p->gv.alloc(bytes)))) ||
        (rc = p->scratch.alloc(64))) {
        delete p;
        return rc;
    }
    cudaError_t e = cudaMemset(p->u[0].p, 0, bytes);
    if (e == cudaSuccess) e = cudaMemset(p->u[1].p, 0, bytes);
    if (e == cudaSuccess && d->has_source) e = cudaMemset(p->src.p, 0, bytes);
    if (e == cudaSuccess && d->has_noise) e = cudaMemset(p->gh.p, 0, bytes);
    if (e == cudaSuccess && d->has_noise) e = cudaMemset(p->gv.p, 0, bytes);
    if (e != cudaSuccess) {
        set_error("plan_create: cudaMemset failed: %s", cudaGetErrorString(e));
        delete p;
        return APDE_ECUDA;
    }
    if ((rc = plan_ensure_resid(p, 4096))) {
        delete p;
        return rc;
    }
    *out = p;
    return APDE_OK;
}

// Colour-split copy of one operand plane (apde_plan.h): dst[row][pc/2 + (pc&1) * pitch/2] = src[row][pc].
template <typename T>
__global__ void __launch_bounds__(256) split_colours_kernel(const T *__restrict__ src, T *__restrict__ dst, long long rows, long long pitch) {
    const long long half = pitch >> 1, n = rows * half;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const long long row = k / half, c = k - row * half;
        const T e = src[row * pitch + 2 * c], o = src[row * pitch + 2 * c + 1];
        dst[row * pitch + c] = e;
        dst[row * pitch + half + c] = o;
    }
}

int plan_refresh_split(apde_plan *p, cudaStream_t st) {
    if (!p->split_dirty) return APDE_OK;
    const size_t bytes = p->plane_elems() * (size_t)p->elem;
    const long long rows = p->lrows + 2 * p->rowpad;
    auto one = [&](DevBuf &from, DevBuf &to) -> int {
        if (!from.p) return APDE_OK;
        if (!to.p) APDE_TRY(to.alloc(bytes));
        const int grid = p->sm_count * 8;
        if (p->elem == 8) split_colours_kernel<double><<<grid, 256, 0, st>>>(from.as<double>(), to.as<double>(), rows, p->pitch);
        else split_colours_kernel<float><<<grid, 256, 0, st>>>(from.as<float>(), to.as<float>(), rows, p->pitch);
        return APDE_OK;
    };
    APDE_TRY(one(p->src, p->src_cs));
    APDE_TRY(one(p->gh, p->gh_cs));
    APDE_TRY(one(p->gv, p->gv_cs));
    APDE_CUDA(cudaGetLastError());
    p->split_dirty = false;
    return APDE_OK;
}

int plan_ensure_resid(apde_plan *p, int64_t cap) {
    if (cap <= p->resid_cap) return APDE_OK;
    DevBuf nb;
    APDE_TRY(nb.alloc((size_t)cap * p->elem));
    APDE_CUDA(cudaMemset(nb.p, 0, (size_t)cap * p->elem));
    if (p->resid.p && p->resid_cap > 0)
        APDE_CUDA(cudaMemcpy(nb.p, p->resid.p, (size_t)p->resid_cap * p->elem, cudaMemcpyDeviceToDevice));
    std::swap(p->resid.p, nb.p);
    std::swap(p->resid.bytes, nb.bytes);
    p->resid_cap = cap;
    return APDE_OK;
}

// Host arrays cover global rows [win_lo, win_hi) (noise_v: up to min(win_hi, grows-1)); the strip
// (owned + ghost rows) must lie inside the window.
template <typename T>
static int upload_t(apde_plan *p, int64_t win_lo, int64_t win_hi, const double *grid, const double *nh,
                    const double *nv, const double *src, bool wait) {
    const Geo g = make_geo(p);
    const int64_t grows = p->d.grows, cols = p->d.cols;
    const int64_t r0 = p->row0, r1 = p->row0 + p->lrows;  // global rows [r0, r1) the strip holds
    if (win_lo > r0 || win_hi < r1) {
        set_error("plan_upload: host window [%lld,%lld) does not cover the strip rows [%lld,%lld)",
                  (long long)win_lo, (long long)win_hi, (long long)r0, (long long)r1);
        return APDE_EINVAL;
    }
    if (p->d.has_source && !src) {
        set_error("plan_upload: plan has_source but source is NULL");
        return APDE_EINVAL;
    }
    if (p->d.has_noise && (!nh || !nv)) {
        set_error("plan_upload: plan has_noise but noise_h/noise_v is NULL");
        return APDE_EINVAL;
    }
    const size_t nrow = (size_t)(r1 - r0), skip = (size_t)(r0 - win_lo);
    if (p->stage.bytes < nrow * cols * 8) APDE_TRY(p->stage.alloc(nrow * cols * 8));
    double *stage = p->stage.as<double>();
    const int tb = 256, tg = p->sm_count * 8;
    p->split_dirty = true;
    APDE_CUDA(cudaMemcpyAsync(stage, grid + skip * cols, nrow * cols * 8, cudaMemcpyHostToDevice, 0));
    pack_kernel<T, false><<<tg, tb>>>(stage, r1 - r0, cols, r0, p->U<T>(0), g);
    if (p->d.has_source) {
        APDE_CUDA(cudaMemcpyAsync(stage, src + skip * cols, nrow * cols * 8, cudaMemcpyHostToDevice, 0));
        pack_kernel<T, false><<<tg, tb>>>(stage, r1 - r0, cols, r0, p->S<T>(), g);
    }
    if (p->d.has_noise) {
        if (cols > 1) {
            APDE_CUDA(cudaMemcpyAsync(stage, nh + skip * (cols - 1), nrow * (cols - 1) * 8, cudaMemcpyHostToDevice, 0));
            pack_kernel<T, true><<<tg, tb>>>(stage, r1 - r0, cols - 1, r0, p->GH<T>(), g);
        }
        const int64_t v1 = std::min<int64_t>(r1, grows - 1);  // noise_v has grows-1 rows
        if (v1 > r0) {
            APDE_CUDA(cudaMemcpyAsync(stage, nv + skip * cols, (size_t)(v1 - r0) * cols * 8, cudaMemcpyHostToDevice, 0));
            pack_kernel<T, true><<<tg, tb>>>(stage, v1 - r0, cols, r0, p->GV<T>(), g);
        }
    }
    APDE_CUDA(cudaGetLastError());
    // The colour-split operand planes are rebuilt HERE, behind the packs on the same stream -- not lazily by the first
    // plan_run: in a multi-GPU pass that first call is phase 1 on the side stream, and the phase-2 launch on the main
    // stream would read planes the split kernel is still writing (found in round 2's last session: the halo-exchange
    // end-to-end path returned a few hundred rows with the previous source term after an upload; tests on small grids
    // and the steady-state bench loop never saw it).
    if (p->variant == 1) APDE_TRY(plan_refresh_split(p, 0));
    APDE_CUDA(cudaMemcpyAsync(p->u[1].p, p->u[0].p, p->plane_elems() * sizeof(T), cudaMemcpyDeviceToDevice, 0));
    // from pinned host memory the copies above are asynchronous; callers driving several devices
    // queue all uploads first and wait afterwards (every GPU has its own PCIe link)
    if (wait) APDE_CUDA(cudaStreamSynchronize(0));
    p->cur = 0;
    p->pass_open = false;
    p->filled = true;
    p->launches += 1 + (p->d.has_source ? 1 : 0) + (p->d.has_noise ? 2 : 0);
    return APDE_OK;
}

static int plan_upload_rows_impl(apde_plan *p, int64_t win_lo, int64_t win_hi, const double *grid,
                                 const double *nh, const double *nv, const double *src, bool wait) {
    if (!p || !grid) {
        set_error("plan_upload: NULL argument");
        return APDE_EINVAL;
    }
    APDE_CUDA(cudaSetDevice(p->device));
    return p->elem == 8 ? upload_t<double>(p, win_lo, win_hi, grid, nh, nv, src, wait)
                        : upload_t<float>(p, win_lo, win_hi, grid, nh, nv, src, wait);
}

int plan_upload_rows(apde_plan *p, int64_t win_lo, int64_t win_hi, const double *grid, const double *nh,
                     const double *nv, const double *src) {
    return plan_upload_rows_impl(p, win_lo, win_hi, grid, nh, nv, src, true);
}

int plan_upload(apde_plan *p, const double *grid, const double *nh, const double *nv, const double *src) {
    return plan_upload_rows(p, 0, p ? p->d.grows : 0, grid, nh, nv, src);
}

template <typename T> static int fill_t(apde_plan *p, const apde_fill_desc *f) {
    const Geo g = make_geo(p);
    fill_kernel<T><<<p->sm_count * 8, 256>>>(p->U<T>(0), p->S<T>(),
                                             p->GH<T>(),
                                             p->GV<T>(), g, *f);
    APDE_CUDA(cudaGetLastError());
    APDE_CUDA(cudaMemcpy(p->u[1].p, p->u[0].p, p->plane_elems() * sizeof(T), cudaMemcpyDeviceToDevice));
    p->split_dirty = true;
    if (p->variant == 1) APDE_TRY(plan_refresh_split(p, 0));  // eagerly, like upload_t (see there)
    APDE_CUDA(cudaDeviceSynchronize());
    p->cur = 0;
    p->filled = true;
    p->launches += 1;
    return APDE_OK;
}

int plan_fill(apde_plan *p, const apde_fill_desc *f) {
    if (!p || !f) {
        set_error("plan_fill: NULL argument");
        return APDE_EINVAL;
    }
    APDE_CUDA(cudaSetDevice(p->device));
    return p->elem == 8 ? fill_t<double>(p, f) : fill_t<float>(p, f);
}

// `out` covers global rows [win_lo, win_hi); the owned rows inside the window are written.
// which: 0 = current grid, 1 = source, 2 = gh (conductance of link (i,j)-(i,j+1)), 3 = gv ((i,j)-(i+1,j)).
static int plan_download_rows_impl(apde_plan *p, int which, int64_t win_lo, int64_t win_hi, double *out, bool wait);
int plan_download_rows(apde_plan *p, int64_t win_lo, int64_t win_hi, double *out) {
    return plan_download_rows_impl(p, 0, win_lo, win_hi, out, true);
}
int plan_download_array(apde_plan *p, int which, int64_t win_lo, int64_t win_hi, double *out) {
    return plan_download_rows_impl(p, which, win_lo, win_hi, out, true);
}
static int plan_download_rows_impl(apde_plan *p, int which, int64_t win_lo, int64_t win_hi, double *out, bool wait) {
    if (!p || !out || win_lo < 0 || win_hi > p->d.grows || win_lo >= win_hi || which < 0 || which > 3) {
        set_error("plan_download: bad argument");
        return APDE_EINVAL;
    }
    if ((which == 1 && !p->d.has_source) || (which >= 2 && !p->d.has_noise)) {
        set_error("plan_download: the plan holds no %s array", which == 1 ? "source" : "conductance");
        return APDE_ESTATE;
    }
    APDE_CUDA(cudaSetDevice(p->device));
    const Geo g = make_geo(p);
    const int64_t lo = std::max<int64_t>(win_lo, p->d.row_begin), hi = std::min<int64_t>(win_hi, p->d.row_end);
    if (lo >= hi) return APDE_OK;
    const size_t nrow = (size_t)(hi - lo);
    if (p->stage.bytes < nrow * g.cols * 8) APDE_TRY(p->stage.alloc(nrow * g.cols * 8));
    double *stage = p->stage.as<double>();
    auto pick = [&](auto tag) {
        using E = decltype(tag);
        return which == 0 ? p->U<E>(p->cur) : which == 1 ? p->S<E>() : which == 2 ? p->GH<E>() : p->GV<E>();
    };
    if (p->elem == 8)
        unpack_kernel<double><<<p->sm_count * 8, 256>>>(pick(double{}), g, stage, lo, hi - lo);
    else
        unpack_kernel<float><<<p->sm_count * 8, 256>>>(pick(float{}), g, stage, lo, hi - lo);
    APDE_CUDA(cudaGetLastError());
    APDE_CUDA(cudaMemcpyAsync(out + (size_t)(lo - win_lo) * g.cols, stage, nrow * g.cols * 8, cudaMemcpyDeviceToHost, 0));
    if (wait) APDE_CUDA(cudaStreamSynchronize(0));
    p->launches += 1;
    return APDE_OK;
}

int plan_download(apde_plan *p, double *grid) {
    return plan_download_rows(p, 0, p ? p->d.grows : 1, grid);
}

// ---- v0 driver ---------------------------------------------------------------------------
template <typename T, bool NOISE, bool SRC>
static int run_v0(apde_plan *p, int64_t n_sweeps, int64_t sweep_base, cudaStream_t st) {
    const Geo g = make_geo(p);
    const long long hw = (g.cols - 1) / 2;
    if (hw < 1 || g.lrows < 3) return APDE_OK;
    dim3 block(256);
    dim3 grid((unsigned)((hw + 255) / 256), (unsigned)std::min<long long>(g.lrows - 2, 32768));
    T *u = p->U<T>(p->cur);
    for (int64_t s = 0; s < n_sweeps; ++s) {
        T *slot = p->resid.as<T>() + sweep_base + s;
        for (int colour = 0; colour < 2; ++colour) {
            rb_colour_kernel<T, NOISE, SRC><<<grid, block, 0, st>>>(u, p->S<T>(), p->GH<T>(),
                                                                  p->GV<T>(), g, colour, slot, p->stop_flag);
            p->launches++;
        }
    }
    APDE_CUDA(cudaGetLastError());
    return APDE_OK;
}

template <typename T, bool NOISE, bool SRC>
static int run_general(apde_plan *p, int64_t n_sweeps, int64_t sweep_base, cudaStream_t st) {
    const Geo g = make_geo(p);
    const long long hw = (g.cols - 1) / 2;
    if (hw < 1 || g.lrows < 3) return APDE_OK;
    dim3 block(256);
    dim3 grid((unsigned)((hw + 255) / 256), (unsigned)std::min<long long>(g.lrows - 2, 32768));
    T *u = p->U<T>(p->cur);
    for (int64_t s = 0; s < n_sweeps; ++s) {
        T *slot = p->resid.as<T>() + sweep_base + s;
        for (int colour = 0; colour < 2; ++colour) {
            rb_general_kernel<T, NOISE, SRC><<<grid, block, 0, st>>>(u, p->S<T>(), p->GH<T>(), p->GV<T>(), g, colour, slot,
                                                                   p->stop_flag, (T)p->omega, p->kcl ? 1 : 0);
            p->launches++;
        }
    }
    APDE_CUDA(cudaGetLastError());
    return APDE_OK;
}

template <typename T>
static int run_dispatch(apde_plan *p, int64_t n, int64_t base, int phase, cudaStream_t st) {
    if (p->omega != 1.0 || p->kcl) {
        if (phase == 2) return APDE_OK;  // in place: everything happens in phase 0/1
        const bool N = p->d.has_noise, S = p->d.has_source;
        if (N && S) return run_general<T, true, true>(p, n, base, st);
        if (N) return run_general<T, true, false>(p, n, base, st);
        if (S) return run_general<T, false, true>(p, n, base, st);
        return run_general<T, false, false>(p, n, base, st);
    }
    if (p->variant == 2)
        return sizeof(T) == 8 ? run_pipe_f64(p, n, base, phase, st) : run_pipe_f32(p, n, base, phase, st);
    if (p->variant == 1) {
        APDE_TRY(plan_refresh_split(p, st));
        return sizeof(T) == 8 ? run_stream_f64(p, n, base, phase, st) : run_stream_f32(p, n, base, phase, st);
    }
    if (phase == 2) return APDE_OK;  // v0 does everything in phase 0/1
    const bool N = p->d.has_noise, S = p->d.has_source;
    if (N && S) return run_v0<T, true, true>(p, n, base, st);
    if (N) return run_v0<T, true, false>(p, n, base, st);
    if (S) return run_v0<T, false, true>(p, n, base, st);
    return run_v0<T, false, false>(p, n, base, st);
}

int plan_run(apde_plan *p, int64_t n_sweeps, int64_t sweep_base, int phase, cudaStream_t st) {
    if (!p || n_sweeps < 0 || sweep_base < 0 || phase < 0 || phase > 2) {
        set_error("plan_run: bad argument");
        return APDE_EINVAL;
    }
    if (!p->filled) {
        set_error("plan_run: plan has no data (call apde_plan_upload or apde_plan_fill first)");
        return APDE_ESTATE;
    }
    if ((p->ghost_top || p->ghost_bot) && n_sweeps > p->d.temporal_depth) {
        set_error("plan_run: %lld sweeps between halo exchanges exceeds temporal_depth %d",
                  (long long)n_sweeps, p->d.temporal_depth);
        return APDE_EINVAL;
    }
    APDE_CUDA(cudaSetDevice(p->device));
    if (sweep_base + n_sweeps > p->resid_cap) {
        // growing reallocates: make sure nothing in flight still writes the old array
        APDE_CUDA(cudaDeviceSynchronize());
        APDE_TRY(plan_ensure_resid(p, std::max<int64_t>(2 * p->resid_cap, sweep_base + n_sweeps)));
    }
    // phase 0 zeroes its residual slots itself; phased runs put the two launches on different
    // streams, so the caller clears the slots (apde_plan_clear_residuals) before forking
    if (phase == 0 && n_sweeps > 0)
        APDE_CUDA(cudaMemsetAsync((char *)p->resid.p + (size_t)sweep_base * p->elem, 0, (size_t)n_sweeps * p->elem, st));
    return p->elem == 8 ? run_dispatch<double>(p, n_sweeps, sweep_base, phase, st)
                        : run_dispatch<float>(p, n_sweeps, sweep_base, phase, st);
}

int plan_set_relaxation(apde_plan *p, double omega, int normalised) {
    if (!p || !(omega > 0.0 && omega < 2.0)) {
        set_error("plan_set_relaxation: omega must lie in (0, 2), got %g", omega);
        return APDE_EINVAL;
    }
    p->omega = omega;
    p->kcl = normalised != 0;
    return APDE_OK;
}

int plan_clear_residuals(apde_plan *p, int64_t first, int64_t count, cudaStream_t st) {
    if (!p || first < 0 || count < 0) {
        set_error("plan_clear_residuals: bad argument");
        return APDE_EINVAL;
    }
    APDE_CUDA(cudaSetDevice(p->device));
    if (first + count > p->resid_cap) {
        APDE_CUDA(cudaDeviceSynchronize());
        APDE_TRY(plan_ensure_resid(p, std::max<int64_t>(2 * p->resid_cap, first + count)));
    }
    if (count > 0)
        APDE_CUDA(cudaMemsetAsync((char *)p->resid.p + (size_t)first * p->elem, 0, (size_t)count * p->elem, st));
    return APDE_OK;
}

int plan_halo(apde_plan *p, int side, void **send_ptr, void **recv_ptr, int64_t *nbytes) {
    if (!p || side < 0 || side > 1 || !send_ptr || !recv_ptr || !nbytes) {
        set_error("plan_halo: bad argument");
        return APDE_EINVAL;
    }
    // after a phase-1 launch the fresh rows live in the buffer the open pass is writing
    const size_t rowb = (size_t)p->pitch * p->elem;
    char *base = (char *)p->u[p->pass_open ? 1 - p->cur : p->cur].p + (size_t)p->rowpad * rowb;
    *nbytes = (int64_t)(p->halo * rowb);
    *send_ptr = *recv_ptr = nullptr;
    if (side == 0 && p->ghost_top) {
        *recv_ptr = base;                                  // ghost rows [0, halo)
        *send_ptr = base + (size_t)p->ghost_top * rowb;    // first `halo` owned rows
    } else if (side == 1 && p->ghost_bot) {
        const int64_t own_hi = p->lrows - p->ghost_bot;
        *recv_ptr = base + (size_t)own_hi * rowb;          // ghost rows below
        *send_ptr = base + (size_t)(own_hi - p->halo) * rowb;  // last `halo` owned rows
    }
    return APDE_OK;
}

int plan_read_residuals(apde_plan *p, int64_t first, int64_t count, double *out, cudaStream_t st) {
    if (!p || first < 0 || count < 0 || first + count > p->resid_cap || (!out && count)) {
        set_error("plan_read_residuals: bad range");
        return APDE_EINVAL;
    }
    if (count == 0) return APDE_OK;
    APDE_CUDA(cudaSetDevice(p->device));
    if (p->elem == 8) {
        APDE_CUDA(cudaMemcpyAsync(out, p->resid.as<double>() + first, (size_t)count * 8, cudaMemcpyDeviceToHost, st));
        APDE_CUDA(cudaStreamSynchronize(st));
    } else {
        std::vector<float> tmp((size_t)count);
        APDE_CUDA(cudaMemcpyAsync(tmp.data(), p->resid.as<float>() + first, (size_t)count * 4, cudaMemcpyDeviceToHost, st));
        APDE_CUDA(cudaStreamSynchronize(st));
        for (int64_t k = 0; k < count; ++k) out[k] = (double)tmp[(size_t)k];
    }
    return APDE_OK;
}

int plan_checksum(apde_plan *p, double *sum_out, double *absmax_out, uint64_t *bits_out, cudaStream_t st) {
    if (!p) {
        set_error("plan_checksum: NULL plan");
        return APDE_EINVAL;
    }
    APDE_CUDA(cudaSetDevice(p->device));
    const Geo g = make_geo(p);
    APDE_CUDA(cudaMemsetAsync(p->scratch.p, 0, 24, st));
    if (p->elem == 8)
        checksum_kernel<double><<<p->sm_count * 4, 256, 0, st>>>(p->U<double>(p->cur), g, p->scratch.as<double>());
    else
        checksum_kernel<float><<<p->sm_count * 4, 256, 0, st>>>(p->U<float>(p->cur), g, p->scratch.as<double>());
    APDE_CUDA(cudaGetLastError());
    p->launches++;
    double h[3];
    APDE_CUDA(cudaMemcpyAsync(h, p->scratch.p, 24, cudaMemcpyDeviceToHost, st));
    APDE_CUDA(cudaStreamSynchronize(st));
    if (sum_out) *sum_out = h[0];
    if (absmax_out) *absmax_out = h[1];
    if (bits_out) memcpy(bits_out, &h[2], 8);
    return APDE_OK;
}

int plan_energy_sums(apde_plan *p, double *sum_h2, double *sum_v2, cudaStream_t st) {
    if (!p || !sum_h2 || !sum_v2) {
        set_error("plan_energy_sums: NULL argument");
        return APDE_EINVAL;
    }
    APDE_CUDA(cudaSetDevice(p->device));
    const Geo g = make_geo(p);
    APDE_CUDA(cudaMemsetAsync(p->scratch.p, 0, 16, st));
    if (p->elem == 8)
        energy_kernel<double><<<p->sm_count * 4, 256, 0, st>>>(p->U<double>(p->cur), g, p->scratch.as<double>());
    else
        energy_kernel<float><<<p->sm_count * 4, 256, 0, st>>>(p->U<float>(p->cur), g, p->scratch.as<double>());
    APDE_CUDA(cudaGetLastError());
    p->launches++;
    double h[2];
    APDE_CUDA(cudaMemcpyAsync(h, p->scratch.p, 16, cudaMemcpyDeviceToHost, st));
    APDE_CUDA(cudaStreamSynchronize(st));
    *sum_h2 = h[0];
    *sum_v2 = h[1];
    return APDE_OK;
}

// ------------------------------------------------------------------------------------------
// one-shot host-buffer solve
// ------------------------------------------------------------------------------------------
// Kernel family: 2 = sweep-pipelined shared-memory kernel (apde_rb_pipe.inl), 1 = register-window streaming
// kernel (apde_rb_stream.inl), 0 = per-colour in-place kernels.  APDE_RB_VARIANT=v0|v1|v2 overrides.
int default_rb_variant() {
    if (const char *v = getenv("APDE_RB_VARIANT")) return std::min(2, std::max(0, v[0] == 'v' ? atoi(v + 1) : atoi(v)));
    return 1;
}
int max_temporal_depth() {
    if (const char *v = getenv("APDE_MAX_DEPTH")) return std::min(6, std::max(1, atoi(v)));  // tuning builds
    return default_rb_variant() == 2 ? 6 : 4;
}

// Fused depth the one-shot solves use when the caller passes temporal_depth <= 0: the measured best per
// variant on B200 (profiles/).
int default_temporal_depth(int elem, bool noise) {
    // 16384^2 on B200 (profiles/r2_rb_sweep_async.txt): with the asynchronous shared-memory feed every variant is
    // fastest at the deepest instantiated depth
    (void)elem;
    (void)noise;
    return 4;
}

static apde_plan *g_cached_plan = nullptr;  // never destroyed at exit (the context may be gone by then)

void release_cached_plan() {
    if (g_cached_plan) {
        cudaSetDevice(g_cached_plan->device);
        delete g_cached_plan;
        g_cached_plan = nullptr;
    }
}


// ------------------------------------------------------------------------------------------
// Streamed one-shot solve (fixed sweep count): upload, sweeps and download overlap.
//
// A one-shot call moves the whole problem over PCIe twice (16384^2 fp64 Poisson: 4.3 GB up, 2.1 GB down = 123 ms at
// ~52 GB/s) around 107 ms of sweeps; run one after the other that is 230 ms.  Information travels only 2*TD rows per
// fused pass, so the passes can follow the upload front and the download can follow the last pass -- provided the unit
// that is scheduled does not widen the dependence cone.  Cutting every pass into the SAME row bands does (pass p of band
// b needs pass p-1 of band b+1: 64 passes = 64 bands behind).  Here the bands of pass p are shifted UP by (p-1)*2*TD rows
// (parallelogram tiling): pass p of band b = rows [b*B - s_p, (b+1)*B - s_p), s_p = (p-1)*h, h = 2*TD.  Then
//   * it reads pass p-1 on rows [b*B - s_p - h, (b+1)*B - s_p + h) = up to the END of pass p-1's band b: every pass of
//     band b can run as soon as band b-1 is finished and rows < (b+1)*B + h are on the device;
//   * it writes buffer p%2 on rows whose pass p-2 values nobody needs any more (the next band's pass p-1 reads from
//     (b+1)*B - s_{p-1} - h = (b+1)*B - s_p upwards) -- the two ping-pong buffers suffice;
//   * after its last pass a band is final and goes back to the host while the next bands are still being relaxed.
// Order on the device: band-major (all passes of band 0, then band 1, ...) on one compute stream; uploads (copy + pack +
// colour split per chunk of B rows) run ahead on a second stream, downloads (unpack + copy per band) behind on a third.
// Every node sees exactly the operands of the unstreamed schedule, so the result and the residual history are bitwise
// the same (tests/test_redblack_gpu.py::test_streamed_solve_*).  Only for runs that cannot stop early (tol <= 0): a
// convergence decision needs a sweep's residual over the WHOLE grid before later sweeps may count.
// ------------------------------------------------------------------------------------------
template <typename T, bool RECIP>
__global__ void pack_rows_kernel(const double *__restrict__ in, long long nrows, long long in_cols, T *__restrict__ out0,
                                 T *__restrict__ out1, long long pitch, long long padl) {
    const long long n = nrows * in_cols;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const long long r = k / in_cols, j = k - r * in_cols;
        double v = in[k];
        if (RECIP) v = 1.0 / v;
        out0[r * pitch + padl + j] = (T)v;
        if (out1) out1[r * pitch + padl + j] = (T)v;
    }
}

struct StreamSet {
    cudaStream_t up = nullptr, run = nullptr, run2 = nullptr, down = nullptr;
    std::vector<cudaEvent_t> ev;
    ~StreamSet() {
        for (auto e : ev)
            if (e) cudaEventDestroy(e);
        for (auto s : {up, run, run2, down})
            if (s) cudaStreamDestroy(s);
    }
    int event(cudaEvent_t *out) {
        cudaEvent_t e = nullptr;
        APDE_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ev.push_back(e);
        *out = e;
        return APDE_OK;
    }
};

template <typename T>
static int solve_streamed_t(apde_plan *p, double *grid, const double *nh, const double *nv, const double *src,
                            int64_t n_sweeps, int64_t B, int64_t rbh, float *kernel_ms, int64_t win_lo = 0, int64_t win_hi = -1) {
    const Geo g = make_geo(p);
    const int64_t rows = p->d.grows, cols = p->d.cols, pitch = p->pitch;
    const int TD = std::min(p->d.temporal_depth, 4);
    const int64_t h = 2 * TD, P = (n_sweeps + TD - 1) / TD;
    const int64_t skew_max = (P - 1) * h;
    // result window (apde_solve_redblack_window): residuals are taken over, and the grid is written back for, the rows
    // [win_lo, win_hi) only; launches are cut at the window edges and the pieces outside report into a dump area
    if (win_hi < 0) win_hi = rows;
    const bool windowed = win_lo > 0 || win_hi < rows;
    // Unshifted band edges.  Bands are B rows, except that the first two (B/4, B/2) let the sweeps start after a small
    // upload and the last two (B/2, B/4) leave a small download behind the last pass; together they cover rows + the
    // largest shift (the last bands exist only for the late, far shifted passes).
    std::vector<int64_t> edge0{0};
    {
        const int64_t q = std::max<int64_t>(4 * h, (B / 4) & ~(int64_t)7), hq = std::max<int64_t>(4 * h, (B / 2) & ~(int64_t)7);
        const int64_t need = rows + skew_max;
        int64_t n_full = (need - 2 * (q + hq) + B - 1) / B;
        if (n_full < 0) n_full = 0;
        std::vector<int64_t> sizes{q, hq};
        for (int64_t k = 0; k < n_full; ++k) sizes.push_back(B);
        sizes.push_back(hq);
        sizes.push_back(q);
        for (int64_t sz : sizes) {
            if (edge0.back() >= need) break;
            edge0.push_back(edge0.back() + sz);
        }
    }
    const int64_t nb = (int64_t)edge0.size() - 1, nchunk = nb;
    // chunk c = the rows band c needs beyond chunk c-1 (band b reads rows < edge0[b+1] + h in its first pass)
    auto chunk_edge = [&](int64_t c) { return c <= 0 ? (int64_t)0 : std::min<int64_t>(rows, edge0[(size_t)std::min<int64_t>(c, nb)] + h); };
    auto band_edge = [&](int64_t b, int64_t pass /*1-based*/) {
        return std::min<int64_t>(rows, std::max<int64_t>(0, edge0[(size_t)std::min<int64_t>(b, nb)] - (pass - 1) * h));
    };

    // staging: one chunk going up, one band coming down
    const size_t up_rows = (size_t)(B + h), stage_elems = (up_rows + (size_t)B) * (size_t)cols;
    if (p->stage.bytes < stage_elems * 8) APDE_TRY(p->stage.alloc(stage_elems * 8));
    double *stage_up = p->stage.as<double>(), *stage_dn = stage_up + up_rows * (size_t)cols;
    // colour-split planes: allocated (and their padding rows zeroed) once per plan
    const size_t plane_bytes = p->plane_elems() * sizeof(T);
    auto ensure_plane = [&](DevBuf &from, DevBuf &to) -> int {
        if (!from.p || to.p) return APDE_OK;
        APDE_TRY(to.alloc(plane_bytes));
        APDE_CUDA(cudaMemset(to.p, 0, plane_bytes));
        return APDE_OK;
    };
    APDE_TRY(ensure_plane(p->src, p->src_cs));
    APDE_TRY(ensure_plane(p->gh, p->gh_cs));
    APDE_TRY(ensure_plane(p->gv, p->gv_cs));
    APDE_TRY(plan_ensure_resid(p, windowed ? 2 * n_sweeps : n_sweeps));  // second half: the dump area
    APDE_CUDA(cudaDeviceSynchronize());  // the legacy stream may still hold work on these buffers

    StreamSet ss;
    // the sweeps are the critical path: their stream gets the highest priority, and the conversion kernels of the copy
    // streams (which only have to keep up with PCIe) run on one small CTA per SM beside the two resident sweep CTAs
    int prio_lo = 0, prio_hi = 0;
    APDE_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    APDE_CUDA(cudaStreamCreateWithPriority(&ss.up, cudaStreamNonBlocking, prio_lo));
    APDE_CUDA(cudaStreamCreateWithPriority(&ss.run, cudaStreamNonBlocking, prio_hi));
    APDE_CUDA(cudaStreamCreateWithPriority(&ss.run2, cudaStreamNonBlocking, prio_hi));
    APDE_CUDA(cudaStreamCreateWithPriority(&ss.down, cudaStreamNonBlocking, prio_lo));
    int tg = p->sm_count * 8;
    if (const char *v = getenv("APDE_STREAM_COPY_CTAS")) tg = std::max(1, atoi(v));
    // Two compute lanes: band b runs on lane b % 2, pass p of it behind pass p-1 of band b-1 on the other lane (an event
    // per pass) -- the only dependence between neighbouring bands -- so the last warps of one launch overlap the first
    // of an independent one instead of an idle tail (every launch is a single wave of warps).
    int lanes = 2;
    if (const char *v = getenv("APDE_STREAM_LANES")) lanes = atoi(v) >= 2 ? 2 : 1;
    const int tb = 256;
    APDE_CUDA(cudaMemsetAsync(p->resid.as<T>(), 0, (size_t)n_sweeps * sizeof(T), ss.run));
    cudaEvent_t resid_cleared;
    APDE_TRY(ss.event(&resid_cleared));
    APDE_CUDA(cudaEventRecord(resid_cleared, ss.run));
    APDE_CUDA(cudaStreamWaitEvent(ss.run2, resid_cleared, 0));

    // ---- one loop queues chunk i (upload stream), band i (compute stream) and the download of band i-1 ------------
    // (all asynchronous from pinned host memory; with pageable memory the copies block the host, and this order keeps
    // the compute stream fed while they do)
    std::vector<cudaEvent_t> up_done((size_t)nchunk, nullptr), band_done((size_t)nb, nullptr);
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    APDE_CUDA(cudaEventCreate(&t0));
    APDE_CUDA(cudaEventCreate(&t1));
    struct Timers {
        cudaEvent_t a, b;
        ~Timers() {
            cudaEventDestroy(a);
            cudaEventDestroy(b);
        }
    } timers{t0, t1};
    const int final_buf = (int)(P % 2);
    const T *final_u = p->U<T>(final_buf);
    auto upload_chunk = [&](int64_t c) -> int {
        const int64_t r0 = chunk_edge(c), r1 = chunk_edge(c + 1);
        if (r1 > r0) {
            const int64_t n = r1 - r0;
            auto one = [&](const double *host, int64_t in_cols, int64_t n_rows, bool recip, T *dst0, T *dst1, T *plane, T *plane_cs) -> int {
                if (n_rows <= 0 || in_cols <= 0) return APDE_OK;
                APDE_CUDA(cudaMemcpyAsync(stage_up, host + (size_t)r0 * (size_t)in_cols, (size_t)n_rows * (size_t)in_cols * 8,
                                          cudaMemcpyHostToDevice, ss.up));
                if (recip) pack_rows_kernel<T, true><<<tg, tb, 0, ss.up>>>(stage_up, n_rows, in_cols, dst0 + r0 * pitch, dst1 ? dst1 + r0 * pitch : nullptr, pitch, p->padl);
                else pack_rows_kernel<T, false><<<tg, tb, 0, ss.up>>>(stage_up, n_rows, in_cols, dst0 + r0 * pitch, dst1 ? dst1 + r0 * pitch : nullptr, pitch, p->padl);
                if (plane_cs) split_colours_kernel<T><<<tg, tb, 0, ss.up>>>(plane + r0 * pitch, plane_cs + r0 * pitch, n_rows, pitch);
                p->launches += plane_cs ? 2 : 1;
                return APDE_OK;
            };
            APDE_TRY(one(grid, cols, n, false, p->U<T>(0), p->U<T>(1), nullptr, nullptr));
            if (p->d.has_source) APDE_TRY(one(src, cols, n, false, p->S<T>(), nullptr, p->S<T>(), p->SC<T>()));
            if (p->d.has_noise) {
                APDE_TRY(one(nh, cols - 1, n, true, p->GH<T>(), nullptr, p->GH<T>(), p->GHC<T>()));
                APDE_TRY(one(nv, cols, std::min<int64_t>(r1, rows - 1) - r0, true, p->GV<T>(), nullptr, p->GV<T>(), p->GVC<T>()));
            }
            APDE_CUDA(cudaGetLastError());
        }
        APDE_TRY(ss.event(&up_done[(size_t)c]));
        APDE_CUDA(cudaEventRecord(up_done[(size_t)c], ss.up));
        return APDE_OK;
    };
    std::vector<cudaEvent_t> pass_done[2];  // per pass: the previous band's / this band's
    pass_done[0].assign((size_t)P + 1, nullptr);
    pass_done[1].assign((size_t)P + 1, nullptr);
    auto run_band = [&](int64_t b) -> int {
        cudaStream_t st = (lanes == 2 && (b & 1)) ? ss.run2 : ss.run;
        std::vector<cudaEvent_t> &prev = pass_done[(b + 1) & 1], &mine = pass_done[b & 1];
        APDE_CUDA(cudaStreamWaitEvent(st, up_done[(size_t)std::min<int64_t>(b, nchunk - 1)], 0));
        if (b == 0) APDE_CUDA(cudaEventRecord(t0, st));
        for (int64_t pass = 1; pass <= P; ++pass) {
            const int64_t lo = band_edge(b, pass), hi = (b == nb - 1) ? rows : band_edge(b + 1, pass);
            if (lanes == 2 && b > 0 && pass > 1) APDE_CUDA(cudaStreamWaitEvent(st, prev[(size_t)pass - 1], 0));
            if (hi > lo) {
                const int td = (int)std::min<int64_t>(TD, n_sweeps - (pass - 1) * TD);
                const int src_buf = (int)((pass - 1) % 2);
                // a clipped band (the top band shrinks with every pass, the bottom ones grow) is still cut into the same
                // NUMBER of row blocks: the launch stays one full wave of warps and simply ends sooner
                const int64_t nblk = std::max<int64_t>(1, B / rbh);
                const int64_t rbh_l = std::min<int64_t>(rbh, std::max<int64_t>(32, (hi - lo + nblk - 1) / nblk));
                const int64_t cuts[4] = {lo, std::min(std::max(win_lo, lo), hi), std::max(std::min(win_hi, hi), lo), hi};
                for (int k = 0; k < 3; ++k) {  // above the window / inside / below (one piece unless windowed)
                    const int64_t plo = cuts[k], phi = cuts[k + 1];
                    if (phi <= plo) continue;
                    const int64_t slot = (pass - 1) * TD + (k == 1 ? 0 : n_sweeps);
                    APDE_TRY(sizeof(T) == 8 ? run_stream_band_f64(p, td, slot, src_buf, plo, phi, rbh_l, st)
                                            : run_stream_band_f32(p, td, slot, src_buf, plo, phi, rbh_l, st));
                }
            }
            if (lanes == 2) {
                if (!mine[(size_t)pass]) APDE_TRY(ss.event(&mine[(size_t)pass]));  // slots are reused two bands on: the band in
                APDE_CUDA(cudaEventRecord(mine[(size_t)pass], st));                // between has consumed them (stream order)
            }
        }
        if (b == nb - 1) APDE_CUDA(cudaEventRecord(t1, st));
        APDE_TRY(ss.event(&band_done[(size_t)b]));
        APDE_CUDA(cudaEventRecord(band_done[(size_t)b], st));
        return APDE_OK;
    };
    auto download_band = [&](int64_t b) -> int {
        const int64_t lo = std::max(win_lo, band_edge(b, P)), hi = std::min(win_hi, (b == nb - 1) ? rows : band_edge(b + 1, P));
        if (hi <= lo) return APDE_OK;
        APDE_CUDA(cudaStreamWaitEvent(ss.down, band_done[(size_t)b], 0));
        unpack_kernel<T><<<tg, tb, 0, ss.down>>>(final_u, g, stage_dn, lo, hi - lo);
        APDE_CUDA(cudaGetLastError());
        APDE_CUDA(cudaMemcpyAsync(grid + (size_t)lo * (size_t)cols, stage_dn, (size_t)(hi - lo) * (size_t)cols * 8,
                                  cudaMemcpyDeviceToHost, ss.down));
        p->launches += 1;
        return APDE_OK;
    };
    for (int64_t i = 0; i <= std::max(nchunk, nb); ++i) {
        if (i < nchunk) APDE_TRY(upload_chunk(i));
        if (i < nb) APDE_TRY(run_band(i));
        if (i >= 1 && i - 1 < nb) APDE_TRY(download_band(i - 1));
    }
    APDE_CUDA(cudaStreamSynchronize(ss.up));
    APDE_CUDA(cudaStreamSynchronize(ss.run));
    APDE_CUDA(cudaStreamSynchronize(ss.run2));
    APDE_CUDA(cudaStreamSynchronize(ss.down));
    APDE_CUDA(cudaGetLastError());
    APDE_CUDA(cudaEventElapsedTime(kernel_ms, t0, t1));
    p->cur = final_buf;
    p->pass_open = false;
    p->filled = true;
    p->split_dirty = false;
    return APDE_OK;
}

// Fixed sweep count, result window: see include/apde.h (apde_solve_redblack_window).
int solve_redblack_window(int elem, double *grid, const double *nh, const double *nv, const double *src, int64_t rows,
                          int64_t cols, int64_t n_sweeps, int temporal_depth, int64_t win_lo, int64_t win_hi, double *hist,
                          double *ksec) {
    if (ksec) *ksec = 0.0;
    if (temporal_depth <= 0) temporal_depth = default_temporal_depth(elem, nh != nullptr);
    temporal_depth = std::min(std::min(temporal_depth, max_temporal_depth()), 4);
    int64_t band = 1536, rbh = 192;
    if (const char *v = getenv("APDE_STREAM_BAND")) band = std::max<int64_t>(64, atoll(v));
    if (const char *v = getenv("APDE_STREAM_RBH")) rbh = std::max<int64_t>(8, atoll(v));
    if (rows < 2 * band || band < 8 * (int64_t)temporal_depth || cols < 3 || n_sweeps <= 0 || win_lo < 0 || win_hi > rows || win_lo >= win_hi) {
        set_error("solve_redblack_window: needs rows >= %lld (two bands), n_sweeps > 0 and 0 <= win_lo < win_hi <= rows", (long long)(2 * band));
        return APDE_EINVAL;
    }
    apde_plan_desc d{};
    d.dtype_bytes = elem;
    d.has_noise = (nh != nullptr);
    d.has_source = (src != nullptr);
    d.temporal_depth = temporal_depth;
    d.grows = rows;
    d.cols = cols;
    d.row_begin = 0;
    d.row_end = rows;
    apde_plan *p = nullptr;
    int dev = 0;
    APDE_CUDA(cudaGetDevice(&dev));
    const bool use_cache = getenv("APDE_NO_PLAN_CACHE") == nullptr;
    if (use_cache && g_cached_plan && g_cached_plan->device == dev && memcmp(&g_cached_plan->d, &d, sizeof(d)) == 0) {
        p = g_cached_plan;
        g_cached_plan = nullptr;
        apply_env_knobs(p);
    } else {
        release_cached_plan();
        APDE_TRY(plan_create(&p, &d));
    }
    struct Guard {
        apde_plan *p;
        bool keep;
        ~Guard() {
            if (keep) g_cached_plan = p;
            else delete p;
        }
    } guard{p, use_cache};
    p->omega = 1.0;
    p->kcl = false;
    p->stop_flag = nullptr;
    if (p->variant != 1) {
        set_error("solve_redblack_window: needs the streaming kernel family (APDE_RB_VARIANT unset or v1)");
        return APDE_EINVAL;
    }
    float ms = 0.f;
    APDE_TRY(elem == 8 ? solve_streamed_t<double>(p, grid, nh, nv, src, n_sweeps, band, rbh, &ms, win_lo, win_hi)
                       : solve_streamed_t<float>(p, grid, nh, nv, src, n_sweeps, band, rbh, &ms, win_lo, win_hi));
    APDE_TRY(plan_read_residuals(p, 0, n_sweeps, hist, 0));
    if (ksec) *ksec = ms * 1e-3;
    return APDE_OK;
}

int solve_redblack(int elem, double *grid, const double *nh, const double *nv, const double *src,
                   int64_t rows, int64_t cols, double tol, int64_t max_iter, int check_every,
                   int temporal_depth, double *hist, int64_t *iters, double *ksec, double omega, bool kcl) {
    if (ksec) *ksec = 0.0;
    if (!(omega > 0.0 && omega < 2.0)) {
        set_error("solve_redblack: omega must lie in (0, 2), got %g", omega);
        return APDE_EINVAL;
    }
    if (max_iter <= 0) {
        *iters = max_iter;
        return APDE_OK;
    }
    if (check_every <= 0) check_every = 32;
    if (temporal_depth <= 0) temporal_depth = default_temporal_depth(elem, nh != nullptr);
    temporal_depth = std::min(temporal_depth, max_temporal_depth());
    if (rows < 3 || cols < 3) {
        const int64_t n = (0.0 < tol) ? std::min<int64_t>(check_every, max_iter) : max_iter;
        for (int64_t k = 0; k < n; ++k) hist[k] = 0.0;
        *iters = n;
        return APDE_OK;
    }
    apde_plan_desc d{};
    d.dtype_bytes = elem;
    d.has_noise = (nh != nullptr);
    d.has_source = (src != nullptr);
    d.temporal_depth = temporal_depth;
    d.grows = rows;
    d.cols = cols;
    d.row_begin = 0;
    d.row_end = rows;
    const bool trace = getenv("APDE_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    // Workspace cache: the device arrays of the last one-shot solve are kept and reused when the next
    // call has the same shape / type / operands (cudaMalloc + cudaFree of several GB cost 10-100 ms and
    // vary a lot).  Callers are serialised by the solve mutex.  APDE_NO_PLAN_CACHE=1 disables it,
    // apde_release_cache() frees it.
    apde_plan *p = nullptr;
    const bool use_cache = getenv("APDE_NO_PLAN_CACHE") == nullptr;
    int dev = 0;
    APDE_CUDA(cudaGetDevice(&dev));
    if (use_cache && g_cached_plan && g_cached_plan->device == dev &&
        memcmp(&g_cached_plan->d, &d, sizeof(d)) == 0) {
        p = g_cached_plan;
        g_cached_plan = nullptr;
        apply_env_knobs(p);
    } else {
        release_cached_plan();
        APDE_TRY(plan_create(&p, &d));
    }
    const double t_created = now();
    struct Guard {
        apde_plan *p;
        bool keep;
        ~Guard() {
            if (!p) return;
            if (keep) g_cached_plan = p;
            else delete p;
        }
    } guard{p, use_cache};
    p->omega = omega;
    p->kcl = kcl;
    // fixed-sweep runs of the fused streaming kernel on grids worth it: upload, sweeps and download overlap
    {
        int64_t band = 1536, rbh = 192;  // 8 row blocks x column strips = one wave of the resident warps at 16384 columns (tools/e2e_stream_sweep.py)
        if (const char *v = getenv("APDE_STREAM_BAND")) band = std::max<int64_t>(64, atoll(v));
        if (const char *v = getenv("APDE_STREAM_RBH")) rbh = std::max<int64_t>(8, atoll(v));
        const char *sw = getenv("APDE_STREAMED");  // "0": never, "1": whenever valid (tests), unset: large grids
        const bool valid = !(tol > 0.0) && omega == 1.0 && !kcl && p->variant == 1 && rows >= 2 * band && band >= 4 * 2 * (int64_t)temporal_depth;
        // by default only where it was measured to pay: a band launch is (rows per band / row block) x (column strips) work
        // items, one full wave of the resident warps at >= ~12k columns; narrower grids would run their band launches
        // half empty and keep the plain schedule
        const bool want = sw ? atoi(sw) != 0 : (rows * cols >= (int64_t)1 << 24 && cols >= 12288);
        if (valid && want) {
            float ms = 0.f;
            APDE_TRY(elem == 8 ? solve_streamed_t<double>(p, grid, nh, nv, src, max_iter, band, rbh, &ms)
                               : solve_streamed_t<float>(p, grid, nh, nv, src, max_iter, band, rbh, &ms));
            APDE_TRY(plan_read_residuals(p, 0, max_iter, hist, 0));
            *iters = max_iter;
            if (ksec) *ksec = ms * 1e-3;
            if (trace)
                fprintf(stderr, "[apde] solve_redblack %lldx%lld streamed: create %.1f ms, upload + %lld sweeps + download %.1f ms (first to last pass %.1f ms)\n",
                        (long long)rows, (long long)cols, (t_created - t_begin) * 1e3, (long long)max_iter, (now() - t_created) * 1e3, ms);
            return APDE_OK;
        }
    }
    APDE_TRY(plan_upload(p, grid, nh, nv, src));
    const double t_uploaded = now();
    // ---- sweep loop without host round trips -------------------------------------------------------------
    // Check intervals are queued up to QUEUE_AHEAD deep; after each one a one-block kernel takes the stop decision
    // on the device (stop_check_kernel).  The host only watches a status word in mapped pinned memory -- behind an
    // event of the interval QUEUE_AHEAD back, never behind the newest work -- and stops queueing once it is set; the
    // intervals already queued past the converged one are no-ops on the device.
    struct Pinned {
        SolveStatus *h = nullptr;
        ~Pinned() {
            if (h) cudaFreeHost(h);
        }
    } pin;
    APDE_CUDA(cudaHostAlloc((void **)&pin.h, sizeof(SolveStatus), cudaHostAllocMapped));
    pin.h->stop = 0;
    pin.h->intervals = 0;
    SolveStatus *d_status = nullptr;
    APDE_CUDA(cudaHostGetDevicePointer((void **)&d_status, pin.h, 0));
    const int64_t n_intervals = (max_iter + check_every - 1) / check_every;
    APDE_TRY(plan_ensure_resid(p, max_iter));  // never reallocated while kernels are in flight
    constexpr int QUEUE_AHEAD = 3;
    struct Ev {
        cudaEvent_t e[QUEUE_AHEAD] = {};
        ~Ev() {
            for (auto x : e)
                if (x) cudaEventDestroy(x);
        }
    } evq;
    for (auto &x : evq.e) APDE_CUDA(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
    EventPair ev;
    APDE_TRY(ev.init());
    const bool may_stop = tol > 0.0;  // max_change >= 0 is never below a non-positive (or NaN) tol
    int *dev_flag = reinterpret_cast<int *>(static_cast<char *>(p->scratch.p) + 32);  // bytes 0..23 belong to the checksums
    APDE_CUDA(cudaMemsetAsync(dev_flag, 0, sizeof(int), 0));
    p->stop_flag = may_stop ? dev_flag : nullptr;
    struct ClearFlag {
        apde_plan *p;
        ~ClearFlag() { p->stop_flag = nullptr; }
    } clear_flag{p};
    std::vector<int> cur_after((size_t)n_intervals + 1, p->cur);
    APDE_CUDA(cudaEventRecord(ev.a));
    int64_t queued = 0;
    bool stop = false;
    while (queued < n_intervals && !stop) {
        const int64_t base = queued * check_every, n = std::min<int64_t>(check_every, max_iter - base);
        APDE_TRY(plan_run(p, n, base, 0, 0));
        if (may_stop) {
            if (p->elem == 8)
                stop_check_kernel<double><<<1, 128>>>(p->resid.as<double>() + base, (int)n, tol, (int)(queued + 1), dev_flag, d_status);
            else
                stop_check_kernel<float><<<1, 128>>>(p->resid.as<float>() + base, (int)n, tol, (int)(queued + 1), dev_flag, d_status);
            p->launches++;
            APDE_CUDA(cudaEventRecord(evq.e[queued % QUEUE_AHEAD]));
        }
        ++queued;
        cur_after[(size_t)queued] = p->cur;
        if (may_stop && queued >= QUEUE_AHEAD) {
            // the interval QUEUE_AHEAD back has been decided once its event has passed (its event slot is the one
            // the next interval records into)
            APDE_CUDA(cudaEventSynchronize(evq.e[(queued - QUEUE_AHEAD) % QUEUE_AHEAD]));
            stop = *(volatile int *)&pin.h->stop != 0;
        }
    }
    APDE_CUDA(cudaEventRecord(ev.b));
    APDE_CUDA(cudaEventSynchronize(ev.b));
    APDE_CUDA(cudaGetLastError());
    float total_ms = 0.f;
    APDE_CUDA(cudaEventElapsedTime(&total_ms, ev.a, ev.b));
    int64_t done_intervals = queued;
    if (may_stop && *(volatile int *)&pin.h->stop) done_intervals = pin.h->intervals;
    p->cur = cur_after[(size_t)done_intervals];  // launches queued past the converged interval did nothing
    p->pass_open = false;
    const int64_t done = std::min<int64_t>(done_intervals * check_every, max_iter);
    APDE_TRY(plan_read_residuals(p, 0, done, hist, 0));  // solver.py:281 -- one copy for the whole history
    *iters = done;
    if (ksec) *ksec = total_ms * 1e-3;
    const double t_swept = now();
    const int rc = plan_download(p, grid);
    if (trace)
        fprintf(stderr, "[apde] solve_redblack %lldx%lld: create %.1f ms, upload %.1f ms, %lld sweeps %.1f ms (kernels %.1f ms), download %.1f ms\n",
                (long long)rows, (long long)cols, (t_created - t_begin) * 1e3, (t_uploaded - t_created) * 1e3,
                (long long)done, (t_swept - t_uploaded) * 1e3, total_ms, (now() - t_swept) * 1e3);
    return rc;
}

// ------------------------------------------------------------------------------------------
// one-shot host-buffer solve on several GPUs of one box, ONE process (CUDA peer copies over NVLink).
// Row strips exactly as in the one-process-per-GPU driver (analog_pde_solver/strips.py): per pass of
// <= temporal_depth sweeps every device relaxes its strip, then pushes its outermost `halo` owned rows
// into the neighbours' ghost rows (cudaMemcpyPeerAsync on its own stream) and the neighbours' next
// pass waits for that event.  The ghost rows written belong to the buffer the receiving device's
// running kernel does not read, so no device ever waits inside a pass.  Same arithmetic per node =>
// bitwise equal to the single-GPU result.
// ------------------------------------------------------------------------------------------
int solve_redblack_multi(int elem, double *grid, const double *nh, const double *nv, const double *src,
                         int64_t rows, int64_t cols, double tol, int64_t max_iter, int check_every,
                         int temporal_depth, int n_gpus, double *hist, int64_t *iters, double *ksec) {
    if (ksec) *ksec = 0.0;
    if (check_every <= 0) check_every = 32;
    if (temporal_depth <= 0) temporal_depth = default_temporal_depth(elem, nh != nullptr);
    temporal_depth = std::min(temporal_depth, max_temporal_depth());
    // the sm_100 devices of the box, by ordinal (what apde_device_count counts)
    std::vector<int> devs;
    {
        int have = 0;
        if (cudaGetDeviceCount(&have) != cudaSuccess) have = 0;
        cudaGetLastError();
        for (int d = 0; d < have; ++d) {
            int major = 0;
            if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10)
                devs.push_back(d);
        }
        cudaGetLastError();
    }
    int ndev = std::min<int>(n_gpus, (int)devs.size());
    const int64_t halo = 2 * (int64_t)temporal_depth;
    while (ndev > 1 && rows / ndev < std::max<int64_t>(halo, 8)) --ndev;  // strips must be thicker than their halo
    if (ndev <= 1 || max_iter <= 0 || rows < 3 || cols < 3)
        return solve_redblack(elem, grid, nh, nv, src, rows, cols, tol, max_iter, check_every, temporal_depth,
                              hist, iters, ksec);
    int dev0 = 0;
    APDE_CUDA(cudaGetDevice(&dev0));
    release_cached_plan();  // the single-GPU workspace (several GB on dev0) would only be in the way
    struct Strip {
        apde_plan *plan = nullptr;
        cudaStream_t st = nullptr;
        cudaEvent_t halo_sent = nullptr, t0 = nullptr, t1 = nullptr;
    };
    std::vector<Strip> sp((size_t)ndev);
    auto cleanup = [&]() {
        for (int g = 0; g < ndev; ++g) {
            cudaSetDevice(devs[g]);
            if (sp[g].halo_sent) cudaEventDestroy(sp[g].halo_sent);
            if (sp[g].t0) cudaEventDestroy(sp[g].t0);
            if (sp[g].t1) cudaEventDestroy(sp[g].t1);
            if (sp[g].st) cudaStreamDestroy(sp[g].st);
            delete sp[g].plan;
        }
        cudaSetDevice(dev0);
    };
#define MG_TRY(call)              \
    do {                          \
        int rc__ = (call);        \
        if (rc__ != APDE_OK) {    \
            cleanup();            \
            return rc__;          \
        }                         \
    } while (0)
#define MG_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));      \
            cleanup();                                                                             \
            return APDE_ECUDA;                                                                     \
        }                                                                                          \
    } while (0)
    for (int g = 0; g < ndev; ++g) {
        MG_CUDA(cudaSetDevice(devs[g]));
        // direct NVLink copies where the topology allows; peer access stays enabled for the life of the
        // process (enabling is idempotent and other plans of this process may rely on it)
        for (int o = g - 1; o <= g + 1; o += 2) {
            int can = 0;
            if (o >= 0 && o < ndev && cudaDeviceCanAccessPeer(&can, devs[g], devs[o]) == cudaSuccess && can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(devs[o], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                cudaGetLastError();
            }
        }
        apde_plan_desc d{};
        d.dtype_bytes = elem;
        d.has_noise = (nh != nullptr);
        d.has_source = (src != nullptr);
        d.temporal_depth = temporal_depth;
        d.grows = rows;
        d.cols = cols;
        d.row_begin = rows * g / ndev;
        d.row_end = rows * (g + 1) / ndev;
        MG_TRY(plan_create(&sp[g].plan, &d));
        // only the ping-pong streaming kernel keeps the ghost rows a neighbour pushes into out of the
        // running pass's reach; the in-place per-colour variant (APDE_RB_VARIANT=v0) would race
        if (sp[g].plan->variant == 0) sp[g].plan->variant = 1;
        MG_CUDA(cudaStreamCreateWithFlags(&sp[g].st, cudaStreamNonBlocking));
        MG_CUDA(cudaEventCreateWithFlags(&sp[g].halo_sent, cudaEventDisableTiming));
        MG_CUDA(cudaEventCreate(&sp[g].t0));
        MG_CUDA(cudaEventCreate(&sp[g].t1));
        MG_TRY(plan_upload_rows_impl(sp[g].plan, 0, rows, grid, nh, nv, src, false));  // queued, not waited for
    }
    for (int g = 0; g < ndev; ++g) {
        MG_CUDA(cudaSetDevice(devs[g]));
        MG_CUDA(cudaStreamSynchronize(0));
    }
    int64_t done = 0;
    bool stop = false;
    float total_ms = 0.f;
    std::vector<double> part((size_t)check_every);
    while (done < max_iter && !stop) {
        const int64_t n = std::min<int64_t>(check_every, max_iter - done);
        for (int g = 0; g < ndev; ++g) {
            MG_CUDA(cudaSetDevice(devs[g]));
            MG_CUDA(cudaEventRecord(sp[g].t0, sp[g].st));
        }
        for (int64_t k = 0; k < n; k += temporal_depth) {
            const int64_t m = std::min<int64_t>(temporal_depth, n - k);
            for (int g = 0; g < ndev; ++g) {
                MG_CUDA(cudaSetDevice(devs[g]));
                // the ghost rows this pass reads were written by the neighbours' previous halo pushes
                if (g > 0) MG_CUDA(cudaStreamWaitEvent(sp[g].st, sp[g - 1].halo_sent, 0));
                if (g + 1 < ndev) MG_CUDA(cudaStreamWaitEvent(sp[g].st, sp[g + 1].halo_sent, 0));
                MG_TRY(plan_run(sp[g].plan, m, done + k, 0, sp[g].st));
            }
            for (int g = 0; g < ndev; ++g) {
                MG_CUDA(cudaSetDevice(devs[g]));
                for (int side = 0; side < 2; ++side) {
                    const int o = side == 0 ? g - 1 : g + 1;
                    if (o < 0 || o >= ndev) continue;
                    void *send = nullptr, *recv_unused = nullptr, *osend_unused = nullptr, *orecv = nullptr;
                    int64_t nb = 0, onb = 0;
                    MG_TRY(plan_halo(sp[g].plan, side, &send, &recv_unused, &nb));
                    MG_TRY(plan_halo(sp[o].plan, 1 - side, &osend_unused, &orecv, &onb));
                    MG_CUDA(cudaMemcpyPeerAsync(orecv, devs[o], send, devs[g], (size_t)nb, sp[g].st));
                }
                MG_CUDA(cudaEventRecord(sp[g].halo_sent, sp[g].st));
            }
        }
        std::fill(hist + done, hist + done + n, 0.0);
        for (int g = 0; g < ndev; ++g) {
            MG_CUDA(cudaSetDevice(devs[g]));
            MG_CUDA(cudaEventRecord(sp[g].t1, sp[g].st));
            MG_TRY(plan_read_residuals(sp[g].plan, done, n, part.data(), sp[g].st));
            for (int64_t k = 0; k < n; ++k)
                if (part[(size_t)k] > hist[done + k]) hist[done + k] = part[(size_t)k];
            float ms = 0.f;
            MG_CUDA(cudaEventElapsedTime(&ms, sp[g].t0, sp[g].t1));
            if (g == 0) total_ms += ms;
        }
        for (int64_t k = 0; k < n; ++k)
            if (hist[done + k] < tol) stop = true;
        done += n;
    }
    for (int g = 0; g < ndev; ++g) {
        MG_CUDA(cudaSetDevice(devs[g]));
        MG_CUDA(cudaStreamSynchronize(sp[g].st));
        MG_TRY(plan_download_rows_impl(sp[g].plan, 0, 0, rows, grid, false));  // all devices copy back concurrently
    }
    for (int g = 0; g < ndev; ++g) {
        MG_CUDA(cudaSetDevice(devs[g]));
        MG_CUDA(cudaStreamSynchronize(0));
    }
    *iters = done;
    if (ksec) *ksec = total_ms * 1e-3;
    cleanup();
#undef MG_TRY
#undef MG_CUDA
    return APDE_OK;
}

}  // namespace apde

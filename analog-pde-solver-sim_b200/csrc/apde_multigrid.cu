// apde_multigrid.cu -- geometric multigrid V-cycle for the ideal mesh (no link mismatch), with the fused red-black
// kernels as the smoother (not in the reference).
//
// The reference relaxes  4 u_ij - (u_N + u_S + u_W + u_E) = -src_ij  by Gauss-Seidel (solver.py:350-356), which needs
// O(N^2) sweeps on an N x N grid.  A V(nu1, nu2) cycle reduces the error by a constant factor per cycle at the cost of
// ~ (nu1 + nu2) * 4/3 fine-grid sweeps, independent of N:
//     smooth nu1 sweeps (red-black, fused kernel)  ->  residual r = -src - 4u + sum(neighbours)
//     -> restrict to the next coarser grid (full weighting transposed, src_c = -P^T r; rediscretised 5-point operator)
//     -> recurse with a zero initial error -> prolong the coarse error bilinearly and add it -> smooth nu2 sweeps.
// Coarsening: a level of (r, c) nodes coarsens to ((r+1)/2, (c+1)/2) uniform nodes over the same rectangle while both
// stay >= 3: every second node when the side has 2^k+1-like sizes (nested, spacing ratio exactly 2), otherwise a
// non-nested grid with ratio slightly above 2 (4096 -> 2048 -> 1024 ...); the transfers are hat-function weights
// computed from the node positions, so every size works.  The coarsest level is relaxed by a fixed number of sweeps.
//
// Mismatched links (noise_h / noise_v given): the V-cycle then solves the NORMALISED system of the redblack_kcl mode,
// sum_k g_k (u - u_k) = -src -- the reference's divide-by-4 system is indefinite at the sizes where multigrid matters.
// Only the finest level carries the links (its smoother is the general per-colour kernel with the normalised update,
// its residual uses the conductances); the coarse levels keep the ideal 5-point operator, which a 1 % mismatch
// perturbs by 1 %: the coarse-grid correction stays accurate and the cycle count barely moves.
//
// Every level is an ordinary apde_plan (padded strip layout, plan dtype), so the smoother is literally
// plan_run(); the two transfer kernels below work between plans.  All arithmetic is explicit round-to-nearest with a
// fixed association, restated by the CPU checker of the tests, so cycles are bit-reproducible.
#include <algorithm>
#include <cstring>
#include <memory>
#include <vector>

#include "apde_plan.h"
#include "apde_rb_common.cuh"

namespace apde {

// r_ij = ((((uN + uS) + uW) + uE) - 4 u) - src   at interior nodes, 0 on the ring.
// With links (finest level of a mismatched mesh, Kirchhoff form):
// r_ij = ((((gN*uN + gS*uS) + gW*uW) + gE*uE) - (((gN + gS) + gW) + gE) * u) - src
template <typename T>
__device__ __forceinline__ T mg_residual(const T *__restrict__ u, const T *__restrict__ src, const T *__restrict__ gh,
                                         const T *__restrict__ gv, long long c, long long pitch) {
    T t;
    if (gh) {
        const T gn = gv[c - pitch], gs = gv[c], gw = gh[c - 1], ge = gh[c];
        t = Ar<T>::add(Ar<T>::mul(gn, u[c - pitch]), Ar<T>::mul(gs, u[c + pitch]));
        t = Ar<T>::add(t, Ar<T>::mul(gw, u[c - 1]));
        t = Ar<T>::add(t, Ar<T>::mul(ge, u[c + 1]));
        const T sg = Ar<T>::add(Ar<T>::add(Ar<T>::add(gn, gs), gw), ge);
        t = Ar<T>::sub(t, Ar<T>::mul(sg, u[c]));
    } else {
        t = Ar<T>::add(u[c - pitch], u[c + pitch]);
        t = Ar<T>::add(t, u[c - 1]);
        t = Ar<T>::add(t, u[c + 1]);
        t = Ar<T>::sub(t, Ar<T>::mul((T)4, u[c]));
    }
    if (src) t = Ar<T>::sub(t, src[c]);
    return t;
}

// ---- grid transfer with hat-function weights ---------------------------------------------------------------
// Fine node f and coarse node I of one direction sit at f*h and I*H over the same length, H/h = r =
// (n_fine-1)/(n_coarse-1) (exactly 2 for nested grids, 2.0005 for 4096 -> 2048).  P[f][I] = max(0, 1 - |f/r - I|)
// is linear interpolation; restriction is its transpose.  With the rediscretised 5-point operator the coarse
// equation [4,-1] e = (H^2/h^2) * (P^T r / (r_x r_y)) reduces to [4,-1] e = P^T r, i.e. src_c = -P^T r (for r = 2
// the weights are 1, 1/2, 1/4: full weighting times 4).  Weights are computed in fp64 from integer positions by the
// same expression on the device and in the CPU statement; sums run over ascending fine indices, rows outer.
__device__ __forceinline__ double mg_hat(int f, int I, double inv_r) {
    return fmax(0.0, __dsub_rn(1.0, fabs(__dsub_rn(__dmul_rn((double)f, inv_r), (double)I))));
}

// coarse source src_c(I,J) = -sum_f sum_g P[f][I] P[g][J] r(f,g); coarse u = 0 everywhere (zero initial error, zero ring)
template <typename T>
__global__ void mg_restrict_kernel(const T *__restrict__ uf, const T *__restrict__ sf, const T *__restrict__ ghf,
                                   const T *__restrict__ gvf, long long pitch_f, long long padl_f,
                                   int rows_f, int cols_f, T *__restrict__ uc0, T *__restrict__ uc1, T *__restrict__ sc,
                                   long long pitch_c, long long padl_c, int rows_c, int cols_c, double inv_rx, double inv_ry,
                                   double rx, double ry) {
    const long long n = (long long)rows_c * cols_c;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const int I = (int)(k / cols_c), J = (int)(k - (long long)I * cols_c);
        const long long cc = (long long)I * pitch_c + padl_c + J;
        uc0[cc] = (T)0;
        uc1[cc] = (T)0;
        double acc = 0.0;
        if (I > 0 && I < rows_c - 1 && J > 0 && J < cols_c - 1) {
            // fine nodes with a non-zero weight lie strictly inside (r*(I-1), r*(I+1))
            const int f0 = max(1, (int)floor(rx * (I - 1))), f1 = min(rows_f - 2, (int)ceil(rx * (I + 1)));
            const int g0 = max(1, (int)floor(ry * (J - 1))), g1 = min(cols_f - 2, (int)ceil(ry * (J + 1)));
            for (int f = f0; f <= f1; ++f) {
                const double wf = mg_hat(f, I, inv_rx);
                if (wf == 0.0) continue;
                double row = 0.0;
                for (int g = g0; g <= g1; ++g) {
                    const double wg = mg_hat(g, J, inv_ry);
                    if (wg == 0.0) continue;
                    const double r = (double)mg_residual<T>(uf, sf, ghf, gvf, (long long)f * pitch_f + padl_f + g, pitch_f);
                    row = __dadd_rn(row, __dmul_rn(wg, r));
                }
                acc = __dadd_rn(acc, __dmul_rn(wf, row));
            }
        }
        sc[cc] = (T)(-acc);
    }
}

// fine u += P e: bilinear between the coarse nodes I0 = floor(f/r), I0+1 (rows first, then columns)
// corr_max (optional): max |P e| over the interior, max-accumulated on the bit pattern (for the stop criterion)
template <typename T>
__global__ void mg_prolong_kernel(T *__restrict__ uf, long long pitch_f, long long padl_f, int rows_f, int cols_f,
                                  const T *__restrict__ ec, long long pitch_c, long long padl_c, int rows_c, int cols_c,
                                  double inv_rx, double inv_ry, T *corr_max) {
    const long long n = (long long)(rows_f - 2) * (cols_f - 2);
    T mx = (T)0;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const int i = 1 + (int)(k / (cols_f - 2)), j = 1 + (int)(k - (long long)(i - 1) * (cols_f - 2));
        const double xi = __dmul_rn((double)i, inv_rx), xj = __dmul_rn((double)j, inv_ry);
        const int I = min((int)floor(xi), rows_c - 2), J = min((int)floor(xj), cols_c - 2);
        const double ti = __dsub_rn(xi, (double)I), tj = __dsub_rn(xj, (double)J);
        const T *e = ec + (long long)I * pitch_c + padl_c + J;
        // top = (1-tj)*e00 + tj*e01 ; bot likewise on the next coarse row ; corr = (1-ti)*top + ti*bot
        const double top = __dadd_rn(__dmul_rn(__dsub_rn(1.0, tj), (double)e[0]), __dmul_rn(tj, (double)e[1]));
        const double bot = __dadd_rn(__dmul_rn(__dsub_rn(1.0, tj), (double)e[pitch_c]), __dmul_rn(tj, (double)e[pitch_c + 1]));
        const T corr = (T)__dadd_rn(__dmul_rn(__dsub_rn(1.0, ti), top), __dmul_rn(ti, bot));
        const long long c = (long long)i * pitch_f + padl_f + j;
        uf[c] = Ar<T>::add(uf[c], corr);
        mx = max_ignore_nan(mx, Ar<T>::abs(corr));
    }
    if (corr_max) {
        mx = warp_max(mx);
        if ((threadIdx.x & 31) == 0 && mx > (T)0) atomic_max_nonneg(corr_max, mx);
    }
}

namespace {

struct Level {
    apde_plan *plan = nullptr;
    int rows = 0, cols = 0;
    ~Level() { delete plan; }
};

template <typename T>
int restrict_to(apde_plan *f, apde_plan *c) {
    const int tg = f->sm_count * 8;
    const double rx = (double)(f->d.grows - 1) / (double)(c->d.grows - 1), ry = (double)(f->d.cols - 1) / (double)(c->d.cols - 1);
    mg_restrict_kernel<T><<<tg, 256>>>(f->U<T>(f->cur), f->S<T>(), f->GH<T>(), f->GV<T>(), f->pitch, f->padl, (int)f->d.grows, (int)f->d.cols,
                                       c->U<T>(0), c->U<T>(1), c->S<T>(), c->pitch, c->padl, (int)c->d.grows, (int)c->d.cols,
                                       1.0 / rx, 1.0 / ry, rx, ry);
    APDE_CUDA(cudaGetLastError());
    c->cur = 0;
    c->split_dirty = true;  // the coarse source was rewritten
    c->pass_open = false;
    c->filled = true;
    f->launches++;
    return APDE_OK;
}
template <typename T>
int prolong_from(apde_plan *f, apde_plan *c, bool want_max) {
    const int tg = f->sm_count * 8;
    T *corr_max = nullptr;
    if (want_max) {  // scratch bytes 40..47 (0..23 checksums, 32..35 the stop flag)
        corr_max = reinterpret_cast<T *>(static_cast<char *>(f->scratch.p) + 40);
        APDE_CUDA(cudaMemsetAsync(corr_max, 0, sizeof(T), 0));
    }
    const double rx = (double)(f->d.grows - 1) / (double)(c->d.grows - 1), ry = (double)(f->d.cols - 1) / (double)(c->d.cols - 1);
    mg_prolong_kernel<T><<<tg, 256>>>(f->U<T>(f->cur), f->pitch, f->padl, (int)f->d.grows, (int)f->d.cols,
                                      c->U<T>(c->cur), c->pitch, c->padl, (int)c->d.grows, (int)c->d.cols, 1.0 / rx, 1.0 / ry,
                                      corr_max);
    APDE_CUDA(cudaGetLastError());
    f->launches++;
    return APDE_OK;
}

}  // namespace

// A level of n nodes per side coarsens to (n + 1) / 2: nested (every second node) when n - 1 is even, otherwise a
// uniform grid over the same length with a spacing ratio slightly above 2 (4096 -> 2048: 4095/2047).
int64_t mg_coarse_size(int64_t n) { return (n + 1) / 2; }
int mg_level_count(int64_t rows, int64_t cols) {
    int n = 1;
    while (rows > 3 && cols > 3 && mg_coarse_size(rows) >= 3 && mg_coarse_size(cols) >= 3) {
        rows = mg_coarse_size(rows);
        cols = mg_coarse_size(cols);
        ++n;
    }
    return n;
}

// V-cycles until a whole cycle moves no node by tol or more (the reference's criterion, solver.py:283, applied per
// cycle instead of per sweep -- a single smoothing sweep barely moves a smooth error on a fine grid, so its
// max_change says little about convergence), or max_cycles.  hist[c] = an upper bound of cycle c's max change:
// the sum of its fine-grid smoothing sweeps' max_change values and of max |coarse-grid correction|, accumulated in
// fp64 in the order pre-smoothing sweeps, correction, post-smoothing sweeps.  *cycles = cycles run.
int solve_multigrid(int elem, double *grid, const double *nh, const double *nv, const double *src, int64_t rows,
                    int64_t cols, double tol, int64_t max_cycles, int nu1, int nu2, int coarse_sweeps, double *hist,
                    int64_t *cycles, double *ksec) {
    if (ksec) *ksec = 0.0;
    if (max_cycles <= 0) {
        *cycles = max_cycles;
        return APDE_OK;
    }
    if (rows < 3 || cols < 3) {
        const int64_t n = (0.0 < tol) ? 1 : max_cycles;
        for (int64_t k = 0; k < n; ++k) hist[k] = 0.0;
        *cycles = n;
        return APDE_OK;
    }
    if (nu1 <= 0) nu1 = 2;
    if (nu2 <= 0) nu2 = 2;
    if (coarse_sweeps <= 0) coarse_sweeps = 32;
    const int n_levels = mg_level_count(rows, cols);
    release_cached_plan();
    std::vector<std::unique_ptr<Level>> lv;
    int64_t r = rows, c = cols;
    for (int l = 0; l < n_levels; ++l) {
        auto L = std::make_unique<Level>();
        apde_plan_desc d{};
        d.dtype_bytes = elem;
        d.has_noise = (l == 0 && nh != nullptr) ? 1 : 0;  // links live on the finest level only (see the header)
        d.has_source = (l > 0 || src != nullptr) ? 1 : 0;
        d.temporal_depth = std::min(std::max(nu1, nu2), max_temporal_depth());
        d.grows = r;
        d.cols = c;
        d.row_begin = 0;
        d.row_end = r;
        APDE_TRY(plan_create(&L->plan, &d));
        L->rows = (int)r;
        L->cols = (int)c;
        lv.push_back(std::move(L));
        r = mg_coarse_size(r);
        c = mg_coarse_size(c);
    }
    apde_plan *fine = lv[0]->plan;
    if (nh) fine->kcl = true;  // mismatched mesh: the normalised (Kirchhoff) update, smoothed by the general kernel
    APDE_TRY(plan_upload(fine, grid, nh, nv, src));
    EventPair ev;
    APDE_TRY(ev.init());
    APDE_CUDA(cudaEventRecord(ev.a));
    auto smooth = [&](apde_plan *p, int n, int base) { return plan_run(p, n, base, 0, 0); };
    auto down = [&](apde_plan *f, apde_plan *cz) { return elem == 8 ? restrict_to<double>(f, cz) : restrict_to<float>(f, cz); };
    auto up = [&](apde_plan *f, apde_plan *cz, bool mx) {
        return elem == 8 ? prolong_from<double>(f, cz, mx) : prolong_from<float>(f, cz, mx);
    };
    std::vector<double> sw((size_t)(nu1 + nu2));
    int64_t done = 0;
    bool stop = false;
    while (done < max_cycles && !stop) {
        for (int l = 0; l + 1 < n_levels; ++l) {
            APDE_TRY(smooth(lv[l]->plan, nu1, 0));
            APDE_TRY(down(lv[l]->plan, lv[l + 1]->plan));
        }
        APDE_TRY(smooth(lv[n_levels - 1]->plan, n_levels > 1 ? coarse_sweeps : nu1 + nu2, 0));
        for (int l = n_levels - 2; l >= 0; --l) {
            APDE_TRY(up(lv[l]->plan, lv[l + 1]->plan, l == 0));
            APDE_TRY(smooth(lv[l]->plan, nu2, nu1));  // slots nu1.. : the pre-smoothing values of this cycle stay readable
        }
        // how far did this cycle move the fine grid, at most?
        APDE_TRY(plan_read_residuals(fine, 0, nu1 + nu2, sw.data(), 0));
        double corr = 0.0;
        if (n_levels > 1) {
            char raw[8] = {0};
            APDE_CUDA(cudaMemcpy(raw, static_cast<char *>(fine->scratch.p) + 40, (size_t)elem, cudaMemcpyDeviceToHost));
            if (elem == 8) memcpy(&corr, raw, 8);
            else {
                float f32 = 0.f;
                memcpy(&f32, raw, 4);
                corr = (double)f32;
            }
        }
        double moved = 0.0;
        for (int k = 0; k < nu1; ++k) moved += sw[(size_t)k];
        moved += corr;
        for (int k = 0; k < nu2; ++k) moved += sw[(size_t)(nu1 + k)];
        hist[done] = moved;
        if (moved < tol) stop = true;
        ++done;
    }
    APDE_CUDA(cudaEventRecord(ev.b));
    APDE_CUDA(cudaEventSynchronize(ev.b));
    float ms = 0.f;
    APDE_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
    if (ksec) *ksec = ms * 1e-3;
    *cycles = done;
    return plan_download(fine, grid);
}

}  // namespace apde

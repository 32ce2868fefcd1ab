/*
 * apde_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the relaxation hot path of
 * danieleschmidt/analog-pde-solver-sim, used only as the checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * Nothing under analog-pde-solver-sim_b200/ may link, import or call this file.
 *
 * Parity status: PINNED.  The lexicographic functions below are checked
 * bit-for-bit (grid bytes, residual history, iteration count) against the
 * unmodified Python reference run in the build container; the resulting golden
 * vectors are committed under tests/golden/ (generator: tests/golden/make_golden.py)
 * and re-checked by tests/test_oracle.py on every run.
 *
 * Build:  gcc -O2 -ffp-contract=off -mfma -fopenmp -shared -fPIC   (see oracle/Makefile)
 *         -ffp-contract=off is REQUIRED: the reference is numpy float64 scalar
 *         arithmetic, one IEEE rounding per operation, no fused multiply-add.
 *
 * Reference citations are file:line under the upstream repository root.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define IDX(i, j) ((size_t)(i) * (size_t)cols + (size_t)(j))

/* ------------------------------------------------------------------------
 * Lexicographic Gauss-Seidel, the reference algorithm.
 *   analog_pde_solver/solver.py:262-287  (AnalogPDESolver.solve sweep loop)
 *   analog_pde_solver/solver.py:344-368  (DigitalSolver.solve sweep loop)
 *
 * grid  : rows*cols, C order, boundary ring + initial guess in, solution out
 * nh    : rows*(cols-1)   _noise_h (solver.py:209), or NULL => all ones
 * nv    : (rows-1)*cols   _noise_v (solver.py:210), or NULL => all ones
 * src   : rows*cols       h^2 * f  (solver.py:253-259), or NULL => zeros
 * hist  : capacity max_iter, one max_change per sweep (solver.py:281)
 * returns iterations_run (solver.py:283-287)
 *
 * NULL nh/nv reproduces the digital expression (solver.py:350-356): x / 1.0 == x
 * bit for bit, so one routine covers both classes.  NULL src: x - 0.0 == x.
 * ---------------------------------------------------------------------- */
int64_t apde_oracle_gs_f64(double *grid, const double *nh, const double *nv,
                           const double *src, int64_t rows, int64_t cols,
                           double tol, int64_t max_iter, double *hist)
{
    const int64_t hc = cols - 1; /* pitch of nh */
    for (int64_t it = 0; it < max_iter; ++it) {
        double max_change = 0.0;                               /* :264 */
        for (int64_t i = 1; i < rows - 1; ++i) {               /* :266 */
            for (int64_t j = 1; j < cols - 1; ++j) {           /* :267 */
                double north = grid[IDX(i - 1, j)];
                double south = grid[IDX(i + 1, j)];
                double west = grid[IDX(i, j - 1)];
                double east = grid[IDX(i, j + 1)];
                if (nv) {                                      /* :269-270 */
                    north = north / nv[IDX(i - 1, j)];
                    south = south / nv[IDX(i, j)];
                }
                if (nh) {                                      /* :271-272 */
                    west = west / nh[(size_t)i * hc + (j - 1)];
                    east = east / nh[(size_t)i * hc + j];
                }
                double acc = north + south;                    /* :275, left-assoc */
                acc = acc + west;
                acc = acc + east;
                if (src) acc = acc - src[IDX(i, j)];
                double new_val = acc / 4.0;
                double change = fabs(new_val - grid[IDX(i, j)]); /* :276 */
                if (change > max_change) max_change = change;  /* :277-278, NaN ignored */
                grid[IDX(i, j)] = new_val;                     /* :279 */
            }
        }
        if (hist) hist[it] = max_change;                       /* :281 */
        if (max_change < tol) return it + 1;                   /* :283-285 */
    }
    return max_iter;                                           /* :286-287 */
}

/* ------------------------------------------------------------------------
 * Same algorithm, same bits, all host threads: column-block pipelined
 * lexicographic Gauss-Seidel.  Thread k owns a block of columns and walks
 * (sweep, row) in order; before row i of sweep t it waits until
 *   - thread k-1 has finished (t, i)      (new west value, solver.py:271)
 *   - thread k+1 has finished (t-1, i)    (old east value, solver.py:272)
 * so every operand has exactly the value the serial loop would see.  Used as
 * the multi-core CPU baseline; validated bitwise against apde_oracle_gs_f64.
 * No early stop: runs exactly n_sweeps (tol = 0 behaviour, solver.py:286-287).
 * ---------------------------------------------------------------------- */
int64_t apde_oracle_gs_f64_mt(double *grid, const double *nh, const double *nv,
                              const double *src, int64_t rows, int64_t cols,
                              int64_t n_sweeps, double *hist, int n_threads)
{
#ifndef _OPENMP
    (void)n_threads;
    return apde_oracle_gs_f64(grid, nh, nv, src, rows, cols, 0.0, n_sweeps, hist);
#else
    if (rows < 3 || cols < 3 || n_sweeps <= 0) {
        for (int64_t t = 0; t < n_sweeps; ++t) if (hist) hist[t] = 0.0;
        return n_sweeps;
    }
    int64_t ncol = cols - 2;
    if (n_threads < 1) n_threads = 1;
    if ((int64_t)n_threads > ncol / 8) n_threads = (int)(ncol / 8 > 0 ? ncol / 8 : 1);
    const int64_t hc = cols - 1;
    const int64_t nrow = rows - 2;
    /* progress[k] = number of (sweep,row) units thread k has completed; padded */
    volatile int64_t *progress = (volatile int64_t *)calloc((size_t)n_threads * 16, sizeof(int64_t));
    double *tmax = (double *)calloc((size_t)n_threads * (size_t)n_sweeps, sizeof(double));
#pragma omp parallel num_threads(n_threads)
    {
        const int k = omp_get_thread_num();
        const int nt = omp_get_num_threads();
        const int64_t j0 = 1 + ncol * k / nt, j1 = 1 + ncol * (k + 1) / nt;
        for (int64_t t = 0; t < n_sweeps; ++t) {
            double max_change = 0.0;
            for (int64_t r = 0; r < nrow; ++r) {
                const int64_t i = r + 1;
                const int64_t unit = t * nrow + r;
                if (k > 0)
                    while (__atomic_load_n(&progress[(size_t)(k - 1) * 16], __ATOMIC_ACQUIRE) < unit + 1) { }
                if (k + 1 < nt && t > 0)
                    while (__atomic_load_n(&progress[(size_t)(k + 1) * 16], __ATOMIC_ACQUIRE) < unit - nrow + 1) { }
                for (int64_t j = j0; j < j1; ++j) {
                    double north = grid[IDX(i - 1, j)];
                    double south = grid[IDX(i + 1, j)];
                    double west = grid[IDX(i, j - 1)];
                    double east = grid[IDX(i, j + 1)];
                    if (nv) { north = north / nv[IDX(i - 1, j)]; south = south / nv[IDX(i, j)]; }
                    if (nh) { west = west / nh[(size_t)i * hc + (j - 1)]; east = east / nh[(size_t)i * hc + j]; }
                    double acc = north + south;
                    acc = acc + west;
                    acc = acc + east;
                    if (src) acc = acc - src[IDX(i, j)];
                    double new_val = acc / 4.0;
                    double change = fabs(new_val - grid[IDX(i, j)]);
                    if (change > max_change) max_change = change;
                    grid[IDX(i, j)] = new_val;
                }
                __atomic_store_n(&progress[(size_t)k * 16], unit + 1, __ATOMIC_RELEASE);
            }
            tmax[(size_t)k * n_sweeps + t] = max_change;
        }
    }
    for (int64_t t = 0; t < n_sweeps; ++t) {
        double m = 0.0;
        for (int k = 0; k < n_threads; ++k)
            if (tmax[(size_t)k * n_sweeps + t] > m) m = tmax[(size_t)k * n_sweeps + t];
        if (hist) hist[t] = m;
    }
    free((void *)progress);
    free(tmax);
    return n_sweeps;
#endif
}

/* ------------------------------------------------------------------------
 * Red-black Gauss-Seidel oracle for the throughput path (NOT in the
 * reference; it is the CPU statement of what the CUDA red-black kernels
 * compute, so they can be checked bit-for-bit at a fixed sweep count).
 *
 * Colour 0 = (i + j) even is relaxed first, then colour 1, each from the
 * other colour's freshest values.  Per node, with g = 1/noise precomputed
 * in fp64 and rounded once to the working type (kernel K5):
 *     t = gN*uN; t = fma(gW,uW,t); t = fma(gE,uE,t); t = t - src; tq = t*0.25;
 *     new = fma(gS*0.25, uS, tq)
 *   no mismatch (gh == gv == NULL):
 *     tq = (((uN + uW) + uE) - src) * 0.25;  new = fma(uS, 0.25, tq)
 *   (the south neighbour enters last, through one fused multiply-add: it is
 *   the operand the streaming kernel's stage pipeline produces last)
 *     change = |new - old|
 * max_change of a sweep = max over both colours (NaN ignored, as the
 * reference's strict '>' does, solver.py:277).
 * Runs exactly n_sweeps; hist[t] = max_change of sweep t.
 * ---------------------------------------------------------------------- */
#define RB_BODY(T, FMA, FABS)                                                              \
    const int64_t hc = cols - 1;                                                           \
    for (int64_t t = 0; t < n_sweeps; ++t) {                                               \
        T max_change = (T)0;                                                               \
        for (int colour = 0; colour < 2; ++colour) {                                       \
            _Pragma("omp parallel for schedule(static) reduction(max : max_change)")       \
            for (int64_t i = 1; i < rows - 1; ++i) {                                       \
                int64_t j = 1 + ((i + 1 + colour) & 1);                                    \
                for (; j < cols - 1; j += 2) {                                             \
                    T un = u[IDX(i - 1, j)], us = u[IDX(i + 1, j)];                        \
                    T uw = u[IDX(i, j - 1)], ue = u[IDX(i, j + 1)];                        \
                    T tq, gsq;                                                             \
                    if (gh) {                                                              \
                        tq = gv[IDX(i - 1, j)] * un;                                       \
                        tq = FMA(gh[(size_t)i * hc + (j - 1)], uw, tq);                    \
                        tq = FMA(gh[(size_t)i * hc + j], ue, tq);                          \
                        gsq = gv[IDX(i, j)] * (T)0.25;                                     \
                    } else {                                                               \
                        tq = un + uw;                                                      \
                        tq = tq + ue;                                                      \
                        gsq = (T)0.25;                                                     \
                    }                                                                      \
                    if (src) tq = tq - src[IDX(i, j)];                                     \
                    tq = tq * (T)0.25;                                                     \
                    T nv_ = FMA(gsq, us, tq);                                              \
                    T ch = FABS(nv_ - u[IDX(i, j)]);                                       \
                    if (ch > max_change) max_change = ch;                                  \
                    u[IDX(i, j)] = nv_;                                                    \
                }                                                                          \
            }                                                                              \
        }                                                                                  \
        if (hist) hist[t] = max_change;                                                    \
    }                                                                                      \
    return n_sweeps;

int64_t apde_oracle_rb_f64(double *u, const double *gh, const double *gv, const double *src,
                           int64_t rows, int64_t cols, int64_t n_sweeps, double *hist)
{
    RB_BODY(double, fma, fabs)
}

int64_t apde_oracle_rb_f32(float *u, const float *gh, const float *gv, const float *src,
                           int64_t rows, int64_t cols, int64_t n_sweeps, float *hist)
{
    RB_BODY(float, fmaf, fabsf)
}

/* ------------------------------------------------------------------------
 * Extended red-black relaxation (NOT in the reference -- SURVEY section 8f item 3): over-relaxation and the
 * normalised ("KCL") conductance update, the CPU statement of rb_general_kernel (csrc/apde_redblack.cu).
 *
 *   plain    : t = gN*uN; t = fma(gW,uW,t); t = fma(gE,uE,t); t -= src; gs = fma(gS*0.25, uS, t*0.25)
 *              (exactly apde_oracle_rb_*: the reference's update, divisor 4, solver.py:269-275)
 *   kcl      : dinv = 1 / (((gN + gS) + gW) + gE);  gs = fma(gS*dinv, uS, t*dinv)
 *              -- Kirchhoff's current law at the node, sum(g_k (u_k - u)) = src: the weighted average the
 *              reference's docstring describes; diagonally dominant at any grid size (SURVEY finding 2)
 *   no links : g == 1: t = (uN + uW) + uE; kcl is the same as plain (dinv = 0.25 exactly)
 *   omega    : new = fma(omega, gs - old, old)       (omega == 1.0: new = gs, bit-identical to the plain kernels)
 *   change = |new - old|, max over both colours, NaN ignored.
 * ---------------------------------------------------------------------- */
#define RBX_BODY(T, FMA, FABS)                                                             \
    const int64_t hc = cols - 1;                                                           \
    const int sor = (omega != (T)1);                                                       \
    for (int64_t t = 0; t < n_sweeps; ++t) {                                               \
        T max_change = (T)0;                                                               \
        for (int colour = 0; colour < 2; ++colour) {                                       \
            _Pragma("omp parallel for schedule(static) reduction(max : max_change)")       \
            for (int64_t i = 1; i < rows - 1; ++i) {                                       \
                int64_t j = 1 + ((i + 1 + colour) & 1);                                    \
                for (; j < cols - 1; j += 2) {                                             \
                    T un = u[IDX(i - 1, j)], us = u[IDX(i + 1, j)];                        \
                    T uw = u[IDX(i, j - 1)], ue = u[IDX(i, j + 1)];                        \
                    T tq, gsq, scale = (T)0.25;                                            \
                    if (gh) {                                                              \
                        T gn = gv[IDX(i - 1, j)], gs_ = gv[IDX(i, j)];                     \
                        T gw = gh[(size_t)i * hc + (j - 1)], ge = gh[(size_t)i * hc + j];  \
                        if (kcl) scale = (T)1 / (((gn + gs_) + gw) + ge);                  \
                        tq = gn * un;                                                      \
                        tq = FMA(gw, uw, tq);                                              \
                        tq = FMA(ge, ue, tq);                                              \
                        gsq = gs_ * scale;                                                 \
                    } else {                                                               \
                        tq = un + uw;                                                      \
                        tq = tq + ue;                                                      \
                        gsq = scale;                                                       \
                    }                                                                      \
                    if (src) tq = tq - src[IDX(i, j)];                                     \
                    tq = tq * scale;                                                       \
                    T old = u[IDX(i, j)];                                                  \
                    T nv_ = FMA(gsq, us, tq);                                              \
                    if (sor) nv_ = FMA(omega, nv_ - old, old);                             \
                    T ch = FABS(nv_ - old);                                                \
                    if (ch > max_change) max_change = ch;                                  \
                    u[IDX(i, j)] = nv_;                                                    \
                }                                                                          \
            }                                                                              \
        }                                                                                  \
        if (hist) hist[t] = max_change;                                                    \
    }                                                                                      \
    return n_sweeps;

int64_t apde_oracle_rbx_f64(double *u, const double *gh, const double *gv, const double *src,
                            int64_t rows, int64_t cols, int64_t n_sweeps, double omega, int kcl, double *hist)
{
    RBX_BODY(double, fma, fabs)
}

int64_t apde_oracle_rbx_f32(float *u, const float *gh, const float *gv, const float *src,
                            int64_t rows, int64_t cols, int64_t n_sweeps, float omega, int kcl, float *hist)
{
    RBX_BODY(float, fmaf, fabsf)
}

/* g = 1/noise in fp64 (kernel K5); the fp32 path rounds the fp64 quotient once. */
void apde_oracle_recip_f64(const double *noise, double *g, int64_t n)
{
    for (int64_t k = 0; k < n; ++k) g[k] = 1.0 / noise[k];
}
void apde_oracle_recip_f32(const double *noise, float *g, int64_t n)
{
    for (int64_t k = 0; k < n; ++k) g[k] = (float)(1.0 / noise[k]);
}

int apde_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

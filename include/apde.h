/*
 * apde.h -- C-ABI of libapde.so, the B200 (sm_100a) backend for the relaxation hot
 * path of danieleschmidt/analog-pde-solver-sim.
 *
 * The reference has no FFI: its hot path is a pure-Python triple loop inside
 *   AnalogPDESolver.solve   analog_pde_solver/solver.py:262-287
 *   DigitalSolver.solve     analog_pde_solver/solver.py:344-368
 * so this boundary is what a maintainer would bind with ctypes from those two
 * methods (INTEGRATION.md shows the stub).  Plain pointers and sizes only.
 *
 * Conventions
 *   - every function returns 0 (APDE_OK) or a negative APDE_E* code;
 *     apde_last_error() gives the thread-local message of the last failure.
 *   - there is NO CPU fallback: without a usable sm_100 device the solve
 *     functions fail with APDE_ENODEV.
 *   - host arrays are C-order float64 exactly as the reference holds them:
 *       grid    rows*cols        boundary ring + initial guess in, solution out   (solver.py:250-251, :289)
 *       noise_h rows*(cols-1)    AnalogPDESolver._noise_h (solver.py:209) or NULL (=> 1.0, DigitalSolver)
 *       noise_v (rows-1)*cols    AnalogPDESolver._noise_v (solver.py:210) or NULL
 *       source  rows*cols        h^2 * f (solver.py:253-259) or NULL (=> Laplace)
 *       residual_history  capacity max_iter doubles, one max_change per sweep     (solver.py:281)
 *       iterations_run    number of sweeps executed                               (solver.py:283-287)
 */
#ifndef APDE_H
#define APDE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APDE_VERSION 100 /* 1.0.0, tracks the reference's __version__ (analog_pde_solver/__init__.py:11) */

enum {
    APDE_OK = 0,
    APDE_EINVAL = -1,  /* bad shape / NULL pointer / bad option */
    APDE_ENODEV = -2,  /* no CUDA device of compute capability 10.x */
    APDE_ECUDA = -3,   /* CUDA runtime error (message in apde_last_error) */
    APDE_ENOMEM = -4,  /* device or host allocation failed */
    APDE_ESTATE = -5   /* plan used in the wrong state */
};

int apde_version(void);
int apde_device_count(void);           /* number of sm_100 devices visible, 0 if none */
const char *apde_last_error(void);     /* thread-local, valid until the next call on this thread */

/* ------------------------------------------------------------------------------------------
 * EXACT path -- replaces the sweep loops solver.py:263-287 and :344-368 bit for bit.
 * Anti-diagonal wavefront lexicographic Gauss-Seidel, fp64, compiled -fmad=false.  Stops at the
 * first sweep whose max_change < tol (strict) like solver.py:283-285; tol <= 0 runs max_iter sweeps.
 * kernel_seconds (optional) = device time of the relaxation kernels only.
 * ---------------------------------------------------------------------------------------- */
int apde_solve_exact_f64(double *grid, const double *noise_h, const double *noise_v,
                         const double *source, int64_t rows, int64_t cols, double tol,
                         int64_t max_iter, double *residual_history, int64_t *iterations_run,
                         double *kernel_seconds);
/* Same, choosing the kernel: path 0 = automatic, 1 = one-CTA shared-memory kernel (fails with
 * APDE_EINVAL if the problem does not fit), 2 = multi-CTA ring kernel.  Results are identical. */
int apde_solve_exact_f64_path(double *grid, const double *noise_h, const double *noise_v,
                              const double *source, int64_t rows, int64_t cols, double tol,
                              int64_t max_iter, double *residual_history, int64_t *iterations_run,
                              double *kernel_seconds, int path);

/* ------------------------------------------------------------------------------------------
 * THROUGHPUT path -- red-black Gauss-Seidel (colour (i+j) even first), conductances g = 1/noise
 * precomputed on the device, `temporal_depth` sweeps fused per pass over HBM.  Same fixed point as
 * the reference; iterates differ (different update order), so iterations_run may differ by a few
 * sweeps.  Convergence is tested every `check_every` sweeps: iterations_run is a multiple of
 * check_every (or max_iter) and residual_history holds every executed sweep.
 *   _f64: fp64 arithmetic;  _f32: fp32 arithmetic (host arrays stay float64, converted on device).
 * check_every <= 0 => 32; temporal_depth <= 0 => library default (4); values above the
 * kernel family's maximum (4; 6 for APDE_RB_VARIANT=v2) are clamped.
 * A call that cannot stop early (tol <= 0, i.e. exactly max_iter sweeps) on a grid of >= 2^24 nodes and >= 12288 columns is STREAMED:
 * the host arrays go up in row chunks, the fused passes follow the upload front band by band (parallelogram tiling:
 * the bands of pass p are shifted up by (p-1)*2*temporal_depth rows) and finished bands come back while later ones
 * are still relaxed -- same grid and residual history bit for bit, about half the wall time from pinned host memory.
 * Environment: APDE_STREAMED=0 never / 1 whenever valid; APDE_STREAM_BAND, APDE_STREAM_RBH (rows per band / row block),
 * APDE_STREAM_LANES (1 or 2 compute streams).
 * ---------------------------------------------------------------------------------------- */
int apde_solve_redblack_f64(double *grid, const double *noise_h, const double *noise_v,
                            const double *source, int64_t rows, int64_t cols, double tol,
                            int64_t max_iter, int check_every, int temporal_depth,
                            double *residual_history, int64_t *iterations_run,
                            double *kernel_seconds);
int apde_solve_redblack_f32(double *grid, const double *noise_h, const double *noise_v,
                            const double *source, int64_t rows, int64_t cols, double tol,
                            int64_t max_iter, int check_every, int temporal_depth,
                            double *residual_history, int64_t *iterations_run,
                            double *kernel_seconds);

/* Fixed-sweep solve of a rows x cols array with a RESULT WINDOW: exactly n_sweeps sweeps (rows 0 and rows-1 and both
 * edge columns are fixed, as always), streamed like the calls above; residual_history[k] is the maximum change of sweep k
 * over the rows [win_lo, win_hi) only, and only those rows of `grid` are written back.  This is the building block of
 * communication-avoiding row strips: information moves 2 rows per sweep, so a strip extended by G >= 2*n_sweeps rows on
 * each interior side (an even number, to keep the red-black colouring) reproduces, inside its window, the sweeps of the
 * whole grid bit for bit -- no halo exchange; the ranks only max-combine their residual histories.  Needs
 * rows >= 2 bands (APDE_STREAM_BAND, default 1536) and the default kernel family. */
int apde_solve_redblack_window(int dtype_bytes, double *grid, const double *noise_h, const double *noise_v,
                               const double *source, int64_t rows, int64_t cols, int64_t n_sweeps, int temporal_depth,
                               int64_t win_lo, int64_t win_hi, double *residual_history, double *kernel_seconds);

/* The same one-shot solve spread over up to n_gpus devices of this box from ONE process: row strips,
 * halo rows pushed to the neighbours with CUDA peer copies over NVLink once per fused pass, per-sweep
 * residuals max-combined on the host.  dtype_bytes 8 = fp64, 4 = fp32.  Result is bitwise equal to
 * the single-GPU solve.  (The one-process-per-GPU route is apde_plan_* + torch.distributed.) */
int apde_solve_redblack_multi(int dtype_bytes, double *grid, const double *noise_h, const double *noise_v,
                              const double *source, int64_t rows, int64_t cols, double tol,
                              int64_t max_iter, int check_every, int temporal_depth, int n_gpus,
                              double *residual_history, int64_t *iterations_run, double *kernel_seconds);

/* ------------------------------------------------------------------------------------------
 * EXTENDED relaxation (not in the reference) -- red-black sweeps with
 *   omega       over-relaxation, new = old + omega * (gauss_seidel_value - old), 0 < omega < 2.  For the ideal mesh the
 *               optimum is 2 / (1 + sqrt(1 - rho^2)), rho = (cos(pi/(rows-1)) + cos(pi/(cols-1))) / 2, which cuts the
 *               sweep count from O(N^2) to O(N).
 *   normalised  != 0: each node divides by the sum of ITS four link conductances instead of the reference's constant 4
 *               (solver.py:275), i.e. Kirchhoff's current law sum g_k (u_k - u) = source.  That system is diagonally
 *               dominant at every grid size, whereas the reference's (4I - W)u = b turns indefinite -- and the
 *               iteration diverges -- for N >~ 200 at 1 % mismatch (SURVEY finding 2).  Without links it is the same
 *               update as the reference's.
 * omega == 1 and normalised == 0 is apde_solve_redblack_f64/f32.  Runs on the per-colour general kernel (one pass over
 * HBM per colour); check_every as above; residual_history / iterations_run as above.
 * ---------------------------------------------------------------------------------------- */
int apde_solve_redblack_ex(int dtype_bytes, double *grid, const double *noise_h, const double *noise_v,
                           const double *source, int64_t rows, int64_t cols, double tol, int64_t max_iter,
                           int check_every, double omega, int normalised, double *residual_history,
                           int64_t *iterations_run, double *kernel_seconds);

/* ------------------------------------------------------------------------------------------
 * MULTIGRID (not in the reference) -- geometric V(nu1, nu2) cycles.  Ideal mesh (noise_h ==
 * noise_v == NULL: DigitalSolver, or AnalogPDESolver with noise_level 0): the reference's system, the fused red-black
 * kernels as smoother.  Mismatched links: the NORMALISED (Kirchhoff) system of apde_solve_redblack_ex(normalised = 1) --
 * the links act on the finest level (general kernel as smoother), the coarse levels keep the ideal operator.  Either way:
 * restriction = transpose of the (bi)linear prolongation, rediscretised 5-point operator on every level.  A level of n
 * nodes per side coarsens to (n+1)/2 uniform nodes over the same length while both sides keep >= 3 nodes: nested
 * (every second node) for 2^k+1-like sides, otherwise non-nested with a spacing ratio slightly above 2 (4096 -> 2048),
 * the transfer weights being hat functions of the node positions -- every grid size works;
 * apde_multigrid_levels() says how many levels a shape gets.
 * A cycle costs about (nu1 + nu2) * 4/3 fine-grid sweeps and reduces the error by a factor independent of the grid
 * size.  residual_history[c] bounds how far cycle c moved any node: the sum of its fine-grid smoothing sweeps' max_change
 * and of max |coarse-grid correction|; the solve stops after the first cycle where that is < tol (the reference's
 * criterion, solver.py:283, applied to a whole cycle -- one smoothing sweep barely moves a smooth error on a fine grid,
 * so a per-sweep test would stop far too early) or after max_cycles; *cycles_run is the number of cycles.  nu1 / nu2 / coarse_sweeps <= 0 select 2 / 2 / 32.
 * ---------------------------------------------------------------------------------------- */
int apde_solve_multigrid(int dtype_bytes, double *grid, const double *noise_h, const double *noise_v,
                         const double *source, int64_t rows, int64_t cols, double tol, int64_t max_cycles, int nu1, int nu2,
                         int coarse_sweeps, double *residual_history, int64_t *cycles_run, double *kernel_seconds);
int apde_multigrid_levels(int64_t rows, int64_t cols);

/* The fused depth the one-shot solves use for temporal_depth <= 0 (dtype_bytes 8 / 4, with or without mismatch). */
int apde_default_temporal_depth(int dtype_bytes, int has_noise);

/* The red-black one-shot solves keep their device workspace for the next call of the same shape
 * (allocation of several GB costs 10-100 ms); this frees it.  APDE_NO_PLAN_CACHE=1 disables the cache. */
int apde_release_cache(void);

/* ------------------------------------------------------------------------------------------
 * Device-resident strip plans (benchmark scale and multi-GPU row strips).
 *
 * A plan owns one row strip of a global grows x cols problem on the CURRENT CUDA device:
 * global rows [row_begin, row_end) plus `halo` ghost rows on each interior side.  Nothing O(N^2)
 * has to cross PCIe: inputs can be generated on the device.  One process per GPU drives one plan;
 * the host (torch.distributed / NCCL) moves the halo rows between plans using the device
 * pointers returned by apde_plan_halo().
 * ---------------------------------------------------------------------------------------- */
typedef struct apde_plan apde_plan;

typedef struct {
    int32_t dtype_bytes;    /* 8 = fp64, 4 = fp32 */
    int32_t has_noise;      /* per-link conductances present */
    int32_t has_source;     /* Poisson source present */
    int32_t temporal_depth; /* max sweeps fused per pass, 1..4; halo = 2*temporal_depth rows */
    int64_t grows, cols;    /* global grid */
    int64_t row_begin;      /* first owned global row */
    int64_t row_end;        /* one past last owned global row */
} apde_plan_desc;

typedef struct {
    int32_t kind;           /* 0 = zeros; 1 = sin source  h^2 * (-2 pi^2 sin(pi x) sin(pi y)),
                               x = i/(grows-1), y = j/(cols-1), h = 1/(max(grows,cols)-1)
                               (tests/test_solver.py:28-32 + solver.py:254-259) */
    int32_t top_hot;        /* 1 => global row 0 = 1.0 (examples/laplace_heat_demo.py:32-38), else zero ring */
    double noise_level;     /* sigma of the multiplicative link mismatch 1 + sigma*N(0,1) */
    uint64_t seed;          /* counter-based Philox4x32-10 key; link id = counter (NOT numpy's PCG64 stream) */
} apde_fill_desc;

int apde_plan_create(apde_plan **plan, const apde_plan_desc *desc);
int apde_plan_destroy(apde_plan *plan);

/* Upload host float64 arrays covering the GLOBAL grid (only this strip's rows are read). */
int apde_plan_upload(apde_plan *plan, const double *grid, const double *noise_h,
                     const double *noise_v, const double *source);
/* Same with host arrays that hold only global rows [row_lo, row_hi) (noise_v: up to
 * min(row_hi, grows-1)); the window must cover the strip's owned + ghost rows.  Lets every rank of
 * a multi-GPU run keep just its own rows in (pinned) host memory. */
int apde_plan_upload_rows(apde_plan *plan, int64_t row_lo, int64_t row_hi, const double *grid,
                          const double *noise_h, const double *noise_v, const double *source);
/* Copy the owned rows that fall inside [row_lo, row_hi) into `out`, which holds exactly that window. */
int apde_plan_download_rows(apde_plan *plan, int64_t row_lo, int64_t row_hi, double *out);
/* Same for any of the plan's arrays, as float64: which 0 = current grid, 1 = source (h^2 f), 2 = gh, the
 * conductance 1/noise_h of link (i,j)-(i,j+1) (0 in the last column), 3 = gv, 1/noise_v of link
 * (i,j)-(i+1,j) (0 in the last row).  Lets a caller read back inputs that apde_plan_fill generated. */
int apde_plan_download_array(apde_plan *plan, int which, int64_t row_lo, int64_t row_hi, double *out);
/* Generate inputs on the device. */
int apde_plan_fill(apde_plan *plan, const apde_fill_desc *fill);
/* Copy this strip's owned rows into a GLOBAL host float64 grid. */
int apde_plan_download(apde_plan *plan, double *grid);

/* Select the extended update (see apde_solve_redblack_ex) for this plan's subsequent apde_plan_run calls. */
int apde_plan_set_relaxation(apde_plan *plan, double omega, int normalised);

/* Run n_sweeps red-black sweeps on `stream` (a cudaStream_t passed as void*, NULL = default),
 * fusing up to `temporal_depth` sweeps per pass.  n_sweeps must be <= the plan's temporal_depth
 * when the strip has neighbours (halo validity); single-strip plans accept any n_sweeps.
 * `phase`: 0 = whole strip, 1 = only the row blocks adjacent to neighbours (launch first, then
 * start the halo exchange), 2 = the remaining interior blocks.  Per-sweep max_change goes to the
 * device residual array at index sweep_base + s (max-accumulated).  Phase 0 zeroes its slots first;
 * for phased runs -- whose two launches may sit on different streams -- call
 * apde_plan_clear_residuals on the stream both launches are ordered after.  Asynchronous. */
int apde_plan_clear_residuals(apde_plan *plan, int64_t first, int64_t count, void *stream);
int apde_plan_run(apde_plan *plan, int64_t n_sweeps, int64_t sweep_base, int phase, void *stream);

/* After a group of sweeps: device pointers + byte counts of the rows to send to / receive from the
 * upper (side 0) and lower (side 1) neighbour.  Pointers are NULL on a side with no neighbour. */
int apde_plan_halo(apde_plan *plan, int side, void **send_ptr, void **recv_ptr, int64_t *nbytes);

/* Device pointer of the per-sweep max_change array (plan dtype) and its capacity. */
int apde_plan_residuals(apde_plan *plan, void **dev_ptr, int64_t *capacity);
/* Copy residuals [first, first+count) to host as float64 (synchronises `stream`). */
int apde_plan_read_residuals(apde_plan *plan, int64_t first, int64_t count, double *out, void *stream);
/* Kernels launched by this plan since creation (bench.py's gpu_launches). */
int64_t apde_plan_launch_count(const apde_plan *plan);
/* sum_j u[row, j] style checksum of owned rows (fp64 accumulate) for cheap cross-rank checks. */
int apde_plan_checksum(apde_plan *plan, double *sum_out, double *abs_max_out, void *stream);
/* Order-independent 64-bit checksum of the owned rows: wrap-around sum of (bit pattern x odd weight of the
 * global position).  Equal bits <=> (with overwhelming probability) equal grids; strips add up (mod 2^64)
 * to the single-strip value, so full-size multi-GPU / multi-kernel runs can be compared without a download. */
int apde_plan_checksum_bits(apde_plan *plan, uint64_t *bits_out, void *stream);

/* On-device pieces of EnergyModel.analog_energy_joules (solver.py:85-92): sum of squared voltage drops
 * over the horizontal links of the owned rows and over the vertical links below each owned row (needs
 * fresh ghost rows).  Sum the two outputs over all strips; energy = (sum_dh2/R + sum_dv2/R) * 5*R*1pF.
 * fp64 accumulation in a different order than numpy: agrees to rounding, not bit for bit. */
int apde_plan_energy_sums(apde_plan *plan, double *sum_dh2, double *sum_dv2, void *stream);

/* Device self-test of the exact path's link division (precomputed-reciprocal sequence vs the general
 * IEEE division) on n pseudo-random + adversarial operand pairs; *mismatches must come back 0. */
int apde_selftest_division(int64_t n, uint64_t seed, int64_t *mismatches);

/* Device-to-device copy on `stream` (same device or peer device, e.g. a neighbour strip's halo
 * rows when both plans live in one process).  Asynchronous. */
int apde_memcpy_dtod(void *dst, const void *src, int64_t nbytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* APDE_H */

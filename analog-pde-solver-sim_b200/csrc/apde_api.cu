// apde_api.cu -- the extern "C" boundary of libapde.so (declared in include/apde.h).
#include <stdarg.h>
#include <string.h>

#include <mutex>

#include "apde_plan.h"

namespace apde {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *get_error() { return g_err; }

int require_sm100(cudaDeviceProp *prop_out) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device available (%s); libapde has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return APDE_ENODEV;
    }
    int dev = 0;
    APDE_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    APDE_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        set_error("device %d (%s) is sm_%d%d; libapde is built for sm_100a only", dev, prop.name,
                  prop.major, prop.minor);
        return APDE_ENODEV;
    }
    if (prop_out) *prop_out = prop;
    return APDE_OK;
}

// one host thread at a time drives the one-shot solves
static std::mutex g_solve_mutex;

}  // namespace apde

using namespace apde;

extern "C" {

int apde_version(void) { return APDE_VERSION; }

int apde_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
    }
    return ok;
}

const char *apde_last_error(void) { return get_error(); }

static int check_solve_args(const char *fn, double *grid, const double *nh, const double *nv,
                            int64_t rows, int64_t cols, int64_t max_iter, double *hist,
                            int64_t *iters) {
    if (!grid || !iters || rows < 1 || cols < 1 || (max_iter > 0 && !hist)) {
        set_error("%s: bad argument (grid=%p rows=%lld cols=%lld hist=%p iters=%p)", fn, (void *)grid,
                  (long long)rows, (long long)cols, (void *)hist, (void *)iters);
        return APDE_EINVAL;
    }
    if ((nh == nullptr) != (nv == nullptr)) {
        set_error("%s: noise_h and noise_v must both be given or both be NULL", fn);
        return APDE_EINVAL;
    }
    return APDE_OK;
}

int apde_solve_exact_f64_path(double *grid, const double *noise_h, const double *noise_v,
                              const double *source, int64_t rows, int64_t cols, double tol,
                              int64_t max_iter, double *residual_history, int64_t *iterations_run,
                              double *kernel_seconds, int path) {
    APDE_TRY(check_solve_args("apde_solve_exact_f64", grid, noise_h, noise_v, rows, cols, max_iter,
                              residual_history, iterations_run));
    std::lock_guard<std::mutex> lock(g_solve_mutex);
    return solve_exact_f64(grid, noise_h, noise_v, source, rows, cols, tol, max_iter,
                           residual_history, iterations_run, kernel_seconds, path);
}

int apde_solve_exact_f64(double *grid, const double *noise_h, const double *noise_v,
                         const double *source, int64_t rows, int64_t cols, double tol,
                         int64_t max_iter, double *residual_history, int64_t *iterations_run,
                         double *kernel_seconds) {
    return apde_solve_exact_f64_path(grid, noise_h, noise_v, source, rows, cols, tol, max_iter,
                                     residual_history, iterations_run, kernel_seconds, 0);
}

int apde_solve_redblack_f64(double *grid, const double *noise_h, const double *noise_v,
                            const double *source, int64_t rows, int64_t cols, double tol,
                            int64_t max_iter, int check_every, int temporal_depth,
                            double *residual_history, int64_t *iterations_run,
                            double *kernel_seconds) {
    APDE_TRY(check_solve_args("apde_solve_redblack_f64", grid, noise_h, noise_v, rows, cols, max_iter,
                              residual_history, iterations_run));
    std::lock_guard<std::mutex> lock(g_solve_mutex);
    return solve_redblack(8, grid, noise_h, noise_v, source, rows, cols, tol, max_iter, check_every,
                          temporal_depth, residual_history, iterations_run, kernel_seconds);
}

int apde_solve_redblack_f32(double *grid, const double *noise_h, const double *noise_v,
                            const double *source, int64_t rows, int64_t cols, double tol,
                            int64_t max_iter, int check_every, int temporal_depth,
                            double *residual_history, int64_t *iterations_run,
                            double *kernel_seconds) {
    APDE_TRY(check_solve_args("apde_solve_redblack_f32", grid, noise_h, noise_v, rows, cols, max_iter,
                              residual_history, iterations_run));
    std::lock_guard<std::mutex> lock(g_solve_mutex);
    return solve_redblack(4, grid, noise_h, noise_v, source, rows, cols, tol, max_iter, check_every,
                          temporal_depth, residual_history, iterations_run, kernel_seconds);
}

int apde_solve_redblack_window(int dtype_bytes, double *grid, const double *noise_h, const double *noise_v,
                               const double *source, int64_t rows, int64_t cols, int64_t n_sweeps, int temporal_depth,
                               int64_t win_lo, int64_t win_hi, double *residual_history, double *kernel_seconds) {
    int64_t dummy = 0;
    APDE_TRY(check_solve_args("apde_solve_redblack_window", grid, noise_h, noise_v, rows, cols, n_sweeps,
                              residual_history, &dummy));
    if (dtype_bytes != 4 && dtype_bytes != 8) {
        set_error("apde_solve_redblack_window: dtype_bytes must be 4 or 8");
        return APDE_EINVAL;
    }
    std::lock_guard<std::mutex> lock(g_solve_mutex);
    return solve_redblack_window(dtype_bytes, grid, noise_h, noise_v, source, rows, cols, n_sweeps, temporal_depth,
                                 win_lo, win_hi, residual_history, kernel_seconds);
}

int apde_solve_redblack_multi(int dtype_bytes, double *grid, const double *noise_h, const double *noise_v,
                              const double *source, int64_t rows, int64_t cols, double tol,
                              int64_t max_iter, int check_every, int temporal_depth, int n_gpus,
                              double *residual_history, int64_t *iterations_run, double *kernel_seconds) {
    APDE_TRY(check_solve_args("apde_solve_redblack_multi", grid, noise_h, noise_v, rows, cols, max_iter,
                              residual_history, iterations_run));
    if (dtype_bytes != 4 && dtype_bytes != 8) {
        set_error("apde_solve_redblack_multi: dtype_bytes must be 4 or 8");
        return APDE_EINVAL;
    }
    std::lock_guard<std::mutex> lock(g_solve_mutex);
    return solve_redblack_multi(dtype_bytes, grid, noise_h, noise_v, source, rows, cols, tol, max_iter,
                                check_every, temporal_depth, n_gpus, residual_history, iterations_run,
                                kernel_seconds);
}

int apde_solve_redblack_ex(int dtype_bytes, double *grid, const double *noise_h, const double *noise_v,
                           const double *source, int64_t rows, int64_t cols, double tol, int64_t max_iter,
                           int check_every, double omega, int normalised, double *residual_history,
                           int64_t *iterations_run, double *kernel_seconds) {
    APDE_TRY(check_solve_args("apde_solve_redblack_ex", grid, noise_h, noise_v, rows, cols, max_iter,
                              residual_history, iterations_run));
    if (dtype_bytes != 4 && dtype_bytes != 8) {
        set_error("apde_solve_redblack_ex: dtype_bytes must be 4 or 8");
        return APDE_EINVAL;
    }
    std::lock_guard<std::mutex> lock(g_solve_mutex);
    return solve_redblack(dtype_bytes, grid, noise_h, noise_v, source, rows, cols, tol, max_iter, check_every, 0,
                          residual_history, iterations_run, kernel_seconds, omega, normalised != 0);
}

int apde_solve_multigrid(int dtype_bytes, double *grid, const double *noise_h, const double *noise_v,
                         const double *source, int64_t rows, int64_t cols, double tol, int64_t max_cycles, int nu1, int nu2,
                         int coarse_sweeps, double *residual_history, int64_t *cycles_run, double *kernel_seconds) {
    APDE_TRY(check_solve_args("apde_solve_multigrid", grid, noise_h, noise_v, rows, cols, max_cycles, residual_history,
                              cycles_run));
    if (dtype_bytes != 4 && dtype_bytes != 8) {
        set_error("apde_solve_multigrid: dtype_bytes must be 4 or 8");
        return APDE_EINVAL;
    }
    std::lock_guard<std::mutex> lock(g_solve_mutex);
    return solve_multigrid(dtype_bytes, grid, noise_h, noise_v, source, rows, cols, tol, max_cycles, nu1, nu2,
                           coarse_sweeps, residual_history, cycles_run, kernel_seconds);
}

int apde_multigrid_levels(int64_t rows, int64_t cols) { return rows >= 3 && cols >= 3 ? mg_level_count(rows, cols) : 1; }

int apde_plan_set_relaxation(apde_plan *plan, double omega, int normalised) {
    return plan_set_relaxation(plan, omega, normalised);
}

int apde_default_temporal_depth(int dtype_bytes, int has_noise) {
    return default_temporal_depth(dtype_bytes, has_noise != 0);
}

int apde_release_cache(void) {
    std::lock_guard<std::mutex> lock(g_solve_mutex);
    release_cached_plan();
    return APDE_OK;
}

int apde_plan_create(apde_plan **plan, const apde_plan_desc *desc) { return plan_create(plan, desc); }

int apde_plan_destroy(apde_plan *plan) {
    if (plan) {
        cudaSetDevice(plan->device);
        delete plan;
    }
    return APDE_OK;
}

int apde_plan_upload(apde_plan *plan, const double *grid, const double *noise_h,
                     const double *noise_v, const double *source) {
    return plan_upload(plan, grid, noise_h, noise_v, source);
}
int apde_plan_upload_rows(apde_plan *plan, int64_t row_lo, int64_t row_hi, const double *grid,
                          const double *noise_h, const double *noise_v, const double *source) {
    return plan_upload_rows(plan, row_lo, row_hi, grid, noise_h, noise_v, source);
}
int apde_plan_download_rows(apde_plan *plan, int64_t row_lo, int64_t row_hi, double *out) {
    return plan_download_rows(plan, row_lo, row_hi, out);
}
int apde_plan_download_array(apde_plan *plan, int which, int64_t row_lo, int64_t row_hi, double *out) {
    return plan_download_array(plan, which, row_lo, row_hi, out);
}
int apde_plan_fill(apde_plan *plan, const apde_fill_desc *fill) { return plan_fill(plan, fill); }
int apde_plan_download(apde_plan *plan, double *grid) { return plan_download(plan, grid); }

int apde_plan_run(apde_plan *plan, int64_t n_sweeps, int64_t sweep_base, int phase, void *stream) {
    return plan_run(plan, n_sweeps, sweep_base, phase, (cudaStream_t)stream);
}

int apde_plan_clear_residuals(apde_plan *plan, int64_t first, int64_t count, void *stream) {
    return plan_clear_residuals(plan, first, count, (cudaStream_t)stream);
}

int apde_plan_halo(apde_plan *plan, int side, void **send_ptr, void **recv_ptr, int64_t *nbytes) {
    return plan_halo(plan, side, send_ptr, recv_ptr, nbytes);
}

int apde_plan_residuals(apde_plan *plan, void **dev_ptr, int64_t *capacity) {
    if (!plan || !dev_ptr || !capacity) {
        set_error("apde_plan_residuals: NULL argument");
        return APDE_EINVAL;
    }
    *dev_ptr = plan->resid.p;
    *capacity = plan->resid_cap;
    return APDE_OK;
}

int apde_plan_read_residuals(apde_plan *plan, int64_t first, int64_t count, double *out, void *stream) {
    return plan_read_residuals(plan, first, count, out, (cudaStream_t)stream);
}

int64_t apde_plan_launch_count(const apde_plan *plan) { return plan ? plan->launches : 0; }

int apde_plan_checksum(apde_plan *plan, double *sum_out, double *abs_max_out, void *stream) {
    return plan_checksum(plan, sum_out, abs_max_out, nullptr, (cudaStream_t)stream);
}
int apde_plan_checksum_bits(apde_plan *plan, uint64_t *bits_out, void *stream) {
    if (!bits_out) {
        set_error("apde_plan_checksum_bits: NULL argument");
        return APDE_EINVAL;
    }
    return plan_checksum(plan, nullptr, nullptr, bits_out, (cudaStream_t)stream);
}

int apde_plan_energy_sums(apde_plan *plan, double *sum_dh2, double *sum_dv2, void *stream) {
    return plan_energy_sums(plan, sum_dh2, sum_dv2, (cudaStream_t)stream);
}

int apde_selftest_division(int64_t n, uint64_t seed, int64_t *mismatches) {
    if (n < 0 || !mismatches) {
        set_error("apde_selftest_division: bad argument");
        return APDE_EINVAL;
    }
    return selftest_division(n, seed, mismatches);
}

int apde_memcpy_dtod(void *dst, const void *src, int64_t nbytes, void *stream) {
    if (!dst || !src || nbytes < 0) {
        set_error("apde_memcpy_dtod: bad argument");
        return APDE_EINVAL;
    }
    APDE_CUDA(cudaMemcpyAsync(dst, src, (size_t)nbytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return APDE_OK;
}

}  // extern "C"

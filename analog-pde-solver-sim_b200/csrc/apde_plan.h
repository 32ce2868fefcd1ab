// Device-resident row-strip plan for the red-black throughput path (internal header).
#pragma once
#include "apde_common.h"

struct apde_plan {
    apde_plan_desc d{};
    int device = 0;
    int sm_count = 0;
    int elem = 8;             // dtype bytes
    // geometry of the local (strip) arrays -----------------------------------------------------
    int64_t halo = 0;         // ghost rows per interior side = 2 * temporal_depth
    int64_t ghost_top = 0;    // ghost rows actually present above / below
    int64_t ghost_bot = 0;
    int64_t lrows = 0;        // ghost_top + owned + ghost_bot
    int64_t row0 = 0;         // global row index of local row 0
    int64_t padl = 0;         // zero columns left of column 0 (alignment + tile overhang)
    int64_t pitch = 0;        // elements per local row
    // buffers (all lrows x pitch, plan dtype) --------------------------------------------------
    apde::DevBuf u[2];        // ping-pong grids; u[cur] is current
    apde::DevBuf src, gh, gv; // source, 1/noise_h, 1/noise_v  (gh[i][j]: link (i,j)-(i,j+1); gv[i][j]: (i,j)-(i+1,j))
    // colour-split copies of the three operand planes for the streaming kernel: within every row the even padded
    // columns come first, then (half a pitch on) the odd ones, so the operands of one colour of a row are contiguous
    // (a warp's operand load touches 2-3 L1 lines instead of 4-8).  Rebuilt from src/gh/gv when split_dirty.
    apde::DevBuf src_cs, gh_cs, gv_cs;
    bool split_dirty = true;
    apde::DevBuf resid;       // per-sweep max_change, plan dtype
    apde::DevBuf scratch;     // checksum outputs etc.
    apde::DevBuf stage;       // float64 staging for host<->device conversion (grown on demand, reused)
    // extended relaxation: over-relaxation factor and normalised ("KCL") conductance update.
    // omega == 1 && !kcl is the reference's update and runs on the fused kernels; anything else runs on the
    // per-colour general kernel.
    double omega = 1.0;
    bool kcl = false;
    int cur = 0;
    bool pass_open = false;   // a phase-1 launch wrote part of u[1-cur]; phase 2 completes it and flips
    int variant = 1;          // 2 = sweep-pipelined shared-memory kernel, 1 = register-window streaming kernel,
                              // 0 = per-colour in-place kernels (cross-check)
    int lane_vectors = 1;     // 16-byte vectors per lane per row in the streaming kernel (1 or 2)
    int64_t row_block = 256;  // rows per work item of the streaming kernel (v2 sizes its blocks itself unless forced)
    bool row_block_forced = false;  // APDE_ROW_BLOCK was given
    int64_t edge_rows = 128;  // v2: rows next to a neighbour strip that form the phase-1 row blocks
    int64_t resid_cap = 0;
    const int *stop_flag = nullptr;  // device-visible flag: when non-null and set, every sweep kernel of this plan
                                     // returns at once (the device-side convergence stop of the one-shot solves)
    int64_t launches = 0;
    bool filled = false;

    int64_t rowpad = 48;      // zero rows above and below every array: the streaming kernel may read
                              // (never write) a few rows past either end without clamping
    size_t plane_elems() const { return (size_t)(lrows + 2 * rowpad) * (size_t)pitch; }
    // pointers to local row 0 (past the top padding)
    template <typename T> T *U(int k) const { return u[k].as<T>() + rowpad * pitch; }
    template <typename T> T *S() const { return src.p ? src.as<T>() + rowpad * pitch : nullptr; }
    template <typename T> T *GH() const { return gh.p ? gh.as<T>() + rowpad * pitch : nullptr; }
    template <typename T> T *GV() const { return gv.p ? gv.as<T>() + rowpad * pitch : nullptr; }
    template <typename T> T *SC() const { return src_cs.p ? src_cs.as<T>() + rowpad * pitch : nullptr; }
    template <typename T> T *GHC() const { return gh_cs.p ? gh_cs.as<T>() + rowpad * pitch : nullptr; }
    template <typename T> T *GVC() const { return gv_cs.p ? gv_cs.as<T>() + rowpad * pitch : nullptr; }
};

namespace apde {
// implemented in apde_redblack.cu
int plan_create(apde_plan **out, const apde_plan_desc *d);
int plan_upload(apde_plan *p, const double *grid, const double *nh, const double *nv, const double *src);
int plan_fill(apde_plan *p, const apde_fill_desc *f);
int plan_download(apde_plan *p, double *grid);
int plan_upload_rows(apde_plan *p, int64_t win_lo, int64_t win_hi, const double *grid, const double *nh, const double *nv, const double *src);
int plan_download_rows(apde_plan *p, int64_t win_lo, int64_t win_hi, double *out);
int plan_download_array(apde_plan *p, int which, int64_t win_lo, int64_t win_hi, double *out);
int plan_run(apde_plan *p, int64_t n_sweeps, int64_t sweep_base, int phase, cudaStream_t st);
int plan_clear_residuals(apde_plan *p, int64_t first, int64_t count, cudaStream_t st);
int plan_halo(apde_plan *p, int side, void **send_ptr, void **recv_ptr, int64_t *nbytes);
int plan_read_residuals(apde_plan *p, int64_t first, int64_t count, double *out, cudaStream_t st);
int plan_energy_sums(apde_plan *p, double *sum_h2, double *sum_v2, cudaStream_t st);
int plan_checksum(apde_plan *p, double *sum_out, double *absmax_out, uint64_t *bits_out, cudaStream_t st);
int plan_ensure_resid(apde_plan *p, int64_t cap);
int plan_refresh_split(apde_plan *p, cudaStream_t st);
void release_cached_plan();
int default_temporal_depth(int elem, bool noise);
int default_rb_variant();
int max_temporal_depth();
int solve_redblack_multi(int elem, double *grid, const double *nh, const double *nv, const double *src,
                         int64_t rows, int64_t cols, double tol, int64_t max_iter, int check_every,
                         int temporal_depth, int n_gpus, double *hist, int64_t *iters, double *ksec);
int solve_redblack(int elem, double *grid, const double *nh, const double *nv, const double *src,
                   int64_t rows, int64_t cols, double tol, int64_t max_iter, int check_every,
                   int temporal_depth, double *hist, int64_t *iters, double *ksec, double omega = 1.0,
                   bool kcl = false);
int solve_redblack_window(int elem, double *grid, const double *nh, const double *nv, const double *src, int64_t rows,
                          int64_t cols, int64_t n_sweeps, int temporal_depth, int64_t win_lo, int64_t win_hi, double *hist,
                          double *ksec);
int plan_set_relaxation(apde_plan *p, double omega, int normalised);
// implemented in apde_multigrid.cu
int mg_level_count(int64_t rows, int64_t cols);
int solve_multigrid(int elem, double *grid, const double *nh, const double *nv, const double *src, int64_t rows,
                    int64_t cols, double tol, int64_t max_cycles, int nu1, int nu2, int coarse_sweeps, double *hist,
                    int64_t *cycles, double *ksec);
// implemented in apde_exact.cu
int selftest_division(int64_t n, uint64_t seed, int64_t *mismatches);
int solve_exact_f64(double *grid, const double *nh, const double *nv, const double *src,
                    int64_t rows, int64_t cols, double tol, int64_t max_iter, double *hist,
                    int64_t *iters, double *ksec, int force_path);
}  // namespace apde

// apde_rb_pipe.inl -- sweep-pipelined red-black kernel ("v2"), included by apde_rb_pipe_f{64,32}.cu with
// RBP_T (element type) and RBP_FN (exported launcher name) defined.
//
// out = RB^TD(in): TD full red-black sweeps fused in ONE pass over HBM, like the register-window kernel
// (apde_rb_stream.inl), but the TD sweeps are spread over the WARPS of a CTA instead of the registers of one
// warp:
//
//   HBM --bulk-async copy (TMA engine, one elected lane)--> input ring in shared memory
//       --> sweep 0 warps --> ring 0 --> sweep 1 warps --> ring 1 --> ... --> sweep TD-1 warps --> HBM
//
// A CTA owns a tile of 120*W columns (+ one 4-column halo group each side) and streams DOWN the rows of a row
// block.  It has TD*W warps -- warp (k, w) runs sweep k on column strip w; lanes 0..3 of warp 0 double as the
// producer (one bulk copy + up to three bulk L2 prefetches per step; a dedicated producer warp would be a fifth
// warp on one SM sub-partition and cost every thread a fifth of its registers).
// Every warp keeps a sliding window of 6 rows x 4 columns per lane in registers (5 live), relaxes colour 0 of
// row L-1 and colour 1 of row L-2 per step (L = the row that just arrived), and hands row L-2 to the next
// sweep through a double-buffered row slot in shared memory.  Consecutive sweeps run 4 rows apart, which
// (a) makes the colour parity of a step the same compile-time constant for every sweep and (b) lets a warp
// fetch its next input row one step before it needs it.  One __syncthreads() per step keeps the CTA in
// lockstep: slot (n & 1) is written in step n and read in step n + 1.
//
// What this buys over one-warp-owns-all-sweeps:
//   * registers per thread no longer grow with TD (the window is 6 rows at any depth): 16 warps per SM
//     instead of 8, dependent chains of two stages instead of 2*TD, depth up to 6;
//   * the W strips of a tile exchange their edge columns through the shared row slots, so the column halo
//     that is recomputed is 2*(TD-1) columns per TILE side, not 2*TD per 128-column warp strip;
//   * u rows enter through cp.async.bulk (SASS UBLKCP) 11 rows ahead and source / conductance rows are pulled
//     into L2 by cp.async.bulk.prefetch.L2 -- no per-thread prefetch instructions.
//
// MEASURED (B200, 16384^2, profiles/r2_pipe_*): bit-exact, but NOT faster than the register-window kernel -- fp64
// Poisson 314 GLUP/s at depth 4 (v1: 447 at depth 3), Laplace 531 (v1: 736), fp64 + mismatch 188 (v1: 194).  A step
// hands over only one row (4 nodes per lane), so the per-step costs -- the CTA barrier (37 % of all warp stall
// samples), fetch / emit / predicate bookkeeping (~120 issued instructions per warp-step, i.e. 30 per node) -- are
// not amortised; replacing the CTA barrier by per-slot mbarrier pairs (sweeps free to drift) measured 15-25 % slower
// still, and 2 columns per lane with twice the warps 25 % slower.  The kernel therefore ships as the alternative
// family APDE_RB_VARIANT=v2 (all tests run on it too); the default stays v1.
//
// Arithmetic per node is rb_partial / rb_finish (apde_rb_common.cuh), the same expression tree as every
// other red-black kernel and the CPU checker, so results are bit-identical across kernels and depths.
#include <algorithm>
#include <type_traits>

#include "apde_ptx.cuh"
#include "apde_rb_common.cuh"

namespace apde {
namespace {

using T = RBP_T;

#ifndef RBP_DI
#define RBP_DI 16
#endif
#ifndef RBP_L2PF
#define RBP_L2PF 16
#endif
#ifndef RBP_OPF
#define RBP_OPF 0
#endif
constexpr int DI = RBP_DI;       // slots of the bulk-copy input ring; rows are requested DI-1 steps ahead
constexpr int L2PF = RBP_L2PF;   // rows ahead that source / conductance rows are pulled from HBM into L2
constexpr int OPF = RBP_OPF;     // 1: operand rows are loaded into registers one step before their first use
#ifndef RBP_L1PF
#define RBP_L1PF 2
#endif
constexpr int L1PF = RBP_L1PF;   // steps ahead that an operand row is pulled from L2 into L1
constexpr int UNROLL = 6;        // step-loop unroll = register-window rows (even: static colour parity)
constexpr int SWEEP_LAG = 4;     // rows between consecutive sweeps

struct PipeArgs {
    const T *in;
    T *out;
    const T *src, *gh, *gv;
    T *resid;  // slot of the first sweep of this pass
    const int *stop;  // optional device flag: set => the solve has converged, do nothing
    long long pitch, padl;
    int lrows, cols;
    long long row0;
    int a_lo, a_hi, b_lo, b_hi;  // local-row ranges this launch covers: [a_lo, a_hi) and [b_lo, b_hi)
    int rbh;                     // row-block height
    int nblk_a, nblk_b, n_tiles, n_items;
};

template <int I, int N, typename F>
__device__ __forceinline__ void pipe_static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        pipe_static_for<I + 1, N>(f);
    }
}

__device__ __forceinline__ double pipe_abs(double x) {
    return __hiloint2double(__double2hiint(x) & 0x7fffffff, __double2loint(x));
}
__device__ __forceinline__ float pipe_abs(float x) { return fabsf(x); }
// NaN never wins and mx is never NaN: identical to the checker's strict '>' (solver.py:277 upstream)
__device__ __forceinline__ double pipe_max(double mx, double v) { return fmax(mx, v); }
__device__ __forceinline__ float pipe_max(float mx, float v) { return fmaxf(mx, v); }

// Columns per lane.  fp64: 2 (one 16-byte vector) -- half the registers of a 4-column lane, so twice the warps
// fit an SM sub-partition (the kernel is bound by dependent-issue latency, not by instruction count), and every
// shared-memory row access is one conflict-free 16-byte vector per lane.  fp32: 4 (also one vector).
#ifndef RBP_C64
#define RBP_C64 4
#endif
constexpr int PIPE_C = sizeof(T) == 8 ? RBP_C64 : 4;

// Warps per SM the register budget is sized for.  The register file is 16 K registers per SM sub-partition, so
// 8 / 6 / 4 warps per sub-partition leave 64 / 80 / 128 registers per thread.
#ifndef RBP_WT_PLAIN
#define RBP_WT_PLAIN 16
#endif
#ifndef RBP_WT_NOISE
#define RBP_WT_NOISE (sizeof(T) == 8 ? 12 : 16)
#endif
__host__ __device__ constexpr int pipe_warp_target(bool noise) { return noise ? RBP_WT_NOISE : RBP_WT_PLAIN; }
// Column strips per CTA: as many as the warp target allows at this depth, at most what keeps a tile (30*C*W + 2*C
// columns) inside the arrays' right-hand padding (480 stored columns).
__host__ __device__ constexpr int pipe_w(int td, bool noise) {
#ifdef RBP_W
    return RBP_W;
#else
    const int wmax = 480 / (30 * PIPE_C), w = pipe_warp_target(noise) / td;
    return w > wmax ? wmax : (w < 2 ? 2 : w);
#endif
}
__host__ __device__ constexpr int pipe_ctas(int td, bool noise) {
    const int warps = td * pipe_w(td, noise), target = pipe_warp_target(noise);
    return target / warps > 0 ? target / warps : 1;
}
__host__ __device__ constexpr int pipe_hc(int td) { return ((2 * (td - 1) + PIPE_C - 1) / PIPE_C) * PIPE_C; }

template <int TD, bool NOISE>
struct PipeGeo {
    static constexpr int W = pipe_w(TD, NOISE);
    static constexpr int CTAS = pipe_ctas(TD, NOISE);
    static constexpr int C = PIPE_C;                           // columns per lane (one group)
    static constexpr int NV = C * (int)sizeof(T) / 16;         // 16-byte vectors per lane per row
    static constexpr int NG = 30 * W + 2;                      // groups per CTA row (1 halo group each side)
    static constexpr int HC = pipe_hc(TD);                     // columns per tile side lost to the open edge, group aligned
    static constexpr int HG = HC / C;
    static constexpr int S = 30 * C * W - 2 * HC;              // columns a tile stores
    static constexpr int ROWB = NG * C * (int)sizeof(T);       // bytes of one CTA row
    static constexpr int SLOT = (ROWB + 127) / 128 * 128;
    static constexpr int NWARP = TD * W;
    static constexpr int NTHR = 32 * NWARP;
    static constexpr int SMEM = DI * SLOT + (TD - 1) * 2 * SLOT + DI * 8;
};

template <int TD, bool NOISE, bool SRC>
__global__ void __launch_bounds__(PipeGeo<TD, NOISE>::NTHR, PipeGeo<TD, NOISE>::CTAS) rb_pipe_kernel(const PipeArgs a) {
    using Geo_ = PipeGeo<TD, NOISE>;
    constexpr int PW = Geo_::W;
    constexpr int C = Geo_::C, NV = Geo_::NV, NG = Geo_::NG, HC = Geo_::HC, HG = Geo_::HG, S = Geo_::S;
    constexpr int ROWB = Geo_::ROWB, SLOT = Geo_::SLOT, NTHR = Geo_::NTHR;
    constexpr int PER = 16 / (int)sizeof(T);  // elements per 16-byte vector

    if (a.stop && *a.stop) return;  // device-side stop (uniform over the grid: the flag only changes between launches)
    extern __shared__ __align__(128) unsigned char pipe_smem[];
    const unsigned s_in = ptx::smem_addr(pipe_smem);
    const unsigned s_ring = s_in + DI * SLOT;
    const unsigned s_bar = s_ring + (TD - 1) * 2 * SLOT;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long pitch = a.pitch;

    // zero every row slot once (halo groups of the inter-sweep rings are never written) and arm the barriers
    for (unsigned o = threadIdx.x * 16u; o < (unsigned)(DI * SLOT + (TD - 1) * 2 * SLOT); o += NTHR * 16u)
        asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(s_in + o), "r"(0u) : "memory");
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < DI; ++s) ptx::mbar_init(s_bar + 8 * s, 1);
        ptx::mbar_fence_init();
    }
    __syncthreads();

    unsigned tma_n = 0;  // rows that went through the input ring so far (slot = n % DI, parity = (n / DI) & 1)
    T mxk = (T)0;        // this warp's running max_change of sweep k over the nodes it owns

    const int k = warp / PW, w = warp % PW;  // sweep, column strip (one strip of every sweep per SM sub-partition)
    const int grp = 30 * w + lane;           // this lane's 4-column group within the CTA row
    const bool first = (k == 0), last = (k == TD - 1);
    // input slot addressing: sweep 0 reads the bulk-copy ring (rows in natural column order), later sweeps read
    // the previous sweep's ring (vector-major: all lanes' vector 0, then all lanes' vector 1 -- conflict free)
    const unsigned in_base = first ? s_in + grp * (16u * NV) : s_ring + (unsigned)(k - 1) * 2u * SLOT + grp * 16u;
    const unsigned in_vstride = first ? 16u : 16u * NG;
    const unsigned out_base = s_ring + (unsigned)k * 2u * SLOT + grp * 16u;  // unused by the last sweep
    const bool emit_lane = lane >= 1 && lane <= 30;
    // producer duties, one lane each on four different SM sub-partitions:
    //   warp 0: one bulk copy of the CTA row of u per step, DI-1 rows ahead of its use
    //   warps 1..3: one bulk L2 prefetch per step of the source / gh / gv row L2PF rows ahead
    const int duty = (lane == 0 && warp < 4) ? warp : -1;

    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        // ---- geometry of the work item (uniform over the CTA) --------------------------------------------
        const int tile = item % a.n_tiles;
        int blk = item / a.n_tiles;
        int rb0, rb1;
        if (blk < a.nblk_a) {
            rb0 = a.a_lo + blk * a.rbh;
            rb1 = min(rb0 + a.rbh, a.a_hi);
        } else {
            blk -= a.nblk_a;
            rb0 = a.b_lo + blk * a.rbh;
            rb1 = min(rb0 + a.rbh, a.b_hi);
        }
        const int ra = max(rb0 - 2 * TD, 0);              // first streamed row (never relaxed)
        const int re = min(rb1 + 2 * TD, a.lrows) - 1;    // last streamed row (never relaxed)
        const int rs = ra - (int)((a.row0 + ra) & 1);     // even global row: static colour parity
        const int r_last = rb1 + 1 + SWEEP_LAG * (TD - 1);  // step at which the last sweep emits row rb1 - 1
        const int n_steps = ((r_last - rs + 1 + UNROLL - 1) / UNROLL) * UNROLL;
        const long long x0 = (long long)tile * S - HC - C;  // grid column of group 0 (multiple of C, may be negative)
        const unsigned upd_span = (unsigned)max(re - ra - 1, 0), own_span = (unsigned)(rb1 - rb0);

        // ---- producer state: the next row to request / prefetch, kept incrementally ---------------------------
        const T *prod_ptr = nullptr;   // warp 0: next u row to copy; warps 1..3: next operand row to prefetch
        int prod_left = 0;             // rows still to request (warp 0) / prefetch (warps 1..3)
        int prod_clamp = 0;            // warp 0: rows until the pointer stops advancing (rows past `re` re-read row re)
        unsigned prod_cnt = tma_n;     // ring counter of the next request
        if (duty == 0) {
            prod_ptr = a.in + (long long)rs * pitch + a.padl + x0;
            prod_left = n_steps + 1;
            prod_clamp = re - rs;
            for (int q = 0; q < DI - 1 && prod_left > 0; ++q) {
                const unsigned slot = prod_cnt & (DI - 1);
                ptx::mbar_expect_tx(s_bar + 8 * slot, ROWB);
                ptx::bulk_load(s_in + slot * SLOT, prod_ptr, ROWB, s_bar + 8 * slot);
                ++prod_cnt;
                --prod_left;
                if (prod_clamp > 0) {
                    prod_ptr += pitch;
                    --prod_clamp;
                }
            }
        } else if (duty > 0) {
            const T *base = duty == 1 ? a.src : duty == 2 ? a.gh : a.gv;
            const bool have = duty == 1 ? SRC : NOISE;
            if (have) {
                // rows ra .. re, starting L2PF rows ahead of the first sweep's first use
                const int r0 = max(ra, rs);
                prod_ptr = base + (long long)r0 * pitch + a.padl + x0;
                prod_left = re - r0 + 1;
                for (int q = 0; q < L2PF && prod_left > 0; ++q) {
                    ptx::bulk_prefetch_l2(prod_ptr, ROWB);
                    prod_ptr += pitch;
                    --prod_left;
                }
            }
        }

        // L1 prefetch pointer (warps 1..3): lane l covers the l-th 128-byte line of the operand row first used in
        // step L1PF (row rs + L1PF - 1 by sweep 0); rows outside the array's padding are never touched
        const char *l1pf_ptr = nullptr;
        if (warp >= 1 && warp <= 3 && lane * 128 < ROWB + 128) {
            const T *base = warp == 1 ? a.src : warp == 2 ? a.gh : a.gv;
            if (warp == 1 ? SRC : NOISE)
                l1pf_ptr = reinterpret_cast<const char *>(base + (long long)(rs + L1PF - 1) * pitch + a.padl + x0) + lane * 128;
        }

        // ---- this lane's columns ----------------------------------------------------------------------------
        const long long jl0 = x0 + C * grp;      // grid column of the lane's first column (even)
        const bool lane_valid = lane >= 1 && lane <= 30 && grp >= 1 + HG && grp < NG - 1 - HG && jl0 < a.cols;
        unsigned colmask = 0;  // bit q: column jl0+q is an interior column
#pragma unroll
        for (int q = 0; q < C; ++q)
            if (jl0 + q >= 1 && jl0 + q <= a.cols - 2) colmask |= 1u << q;
        const bool masked = __any_sync(0xffffffffu, colmask != (1u << C) - 1u);  // warp-uniform: some column here never moves

        const long long col_off = a.padl + jl0;
        // operand pointers at row L (the row arriving in step 0 is rs - SWEEP_LAG*k), advanced by one row per step
        const long long row_l0 = (long long)(rs - SWEEP_LAG * k);
        const T *psrc = a.src + row_l0 * pitch + col_off;
        const T *pgh = a.gh + row_l0 * pitch + col_off;
        const T *pgv = a.gv + row_l0 * pitch + col_off;
        T *pout = a.out + (row_l0 - 2) * pitch + col_off;  // row L - 2
        int row_l = (int)row_l0;                            // L of the current step

        // register window: uin = rows as fetched, unw = the same rows after this sweep's update.  Every element is
        // written exactly once in each array, so no value is ever updated in place (no register copies).
        T uin[UNROLL][C], unw[UNROLL][C];
#pragma unroll
        for (int s = 0; s < UNROLL; ++s)
#pragma unroll
            for (int q = 0; q < C; ++q) uin[s][q] = unw[s][q] = (T)0;
        // operand windows (rows relative to L; sizes divide UNROLL so slots are compile-time)
        constexpr int SW = 2 + OPF, GHW = 2 + OPF, GVW = OPF ? 6 : 3;
        T sv[SRC ? SW : 1][C], gh[NOISE ? GHW : 1][C], ghw[NOISE ? GHW : 1], gv[NOISE ? GVW : 1][C];
#pragma unroll
        for (int q = 0; q < C; ++q) {
#pragma unroll
            for (int s = 0; s < (SRC ? SW : 1); ++s) sv[s][q] = (T)0;
#pragma unroll
            for (int s = 0; s < (NOISE ? GHW : 1); ++s) gh[s][q] = (T)0;
#pragma unroll
            for (int s = 0; s < (NOISE ? GVW : 1); ++s) gv[s][q] = (T)0;
        }
#pragma unroll
        for (int s = 0; s < (NOISE ? GHW : 1); ++s) ghw[s] = (T)0;
        T mx = (T)0;
        unsigned wait_cnt = tma_n;  // sweep 0: ring counter of the next row to wait for

        // fetch of the next arriving row into window slot `dst`; `par` = parity of the step it arrives in
        auto fetch_row = [&](T (&dst)[C], int par) {
            unsigned addr;
            if (first) {
                const unsigned slot = wait_cnt & (DI - 1);
                ptx::mbar_wait(s_bar + 8 * slot, (wait_cnt / DI) & 1u);
                addr = in_base + slot * SLOT;
                ++wait_cnt;
            } else {
                addr = in_base + (unsigned)par * SLOT;  // the previous sweep emitted it two steps before its arrival
            }
#pragma unroll
            for (int v = 0; v < NV; ++v) ptx::lds16(&dst[v * PER], addr + v * in_vstride);
        };
        auto row_active = [&](int i) { return (unsigned)(i - 1 - ra) < upd_span; };  // ra < i < re
        auto row_owned = [&](int i) { return (unsigned)(i - rb0) < own_span; };

        fetch_row(uin[0], 0);
        bool act2 = row_active(row_l - 2), cnt2 = row_owned(row_l - 2);  // carried: this step's row L-2 is the last step's L-1
        for (int nb = 0; nb < n_steps; nb += UNROLL) {
            pipe_static_for<0, UNROLL>([&](auto rr_tag) {
                constexpr int rr = decltype(rr_tag)::value;
                // window slots: row L+d lives in slot (rr + d) mod 6
                constexpr int s_nxt = (rr + 1) % UNROLL, s_l = rr, s_m1 = (rr + 5) % UNROLL, s_m2 = (rr + 4) % UNROLL,
                              s_m3 = (rr + 3) % UNROLL;
                constexpr int p = (rr + 1) & 1;     // column parity relaxed by both stages of this step
                constexpr int qe = p == 0 ? C - 1 : 0;  // the column whose value the neighbour lane needs
                // ---- producer duties ------------------------------------------------------------------------
                if (duty >= 0 && prod_left > 0) {
                    if (duty == 0) {
                        const unsigned slot = prod_cnt & (DI - 1);
                        ptx::mbar_expect_tx(s_bar + 8 * slot, ROWB);
                        ptx::bulk_load(s_in + slot * SLOT, prod_ptr, ROWB, s_bar + 8 * slot);
                        ++prod_cnt;
                        if (prod_clamp > 0) {
                            prod_ptr += pitch;
                            --prod_clamp;
                        }
                    } else {
                        ptx::bulk_prefetch_l2(prod_ptr, ROWB);
                        prod_ptr += pitch;
                    }
                    --prod_left;
                }
                __syncwarp();
                // warps 1..3, all lanes: one instruction pulls the whole CTA row of an operand array, two rows ahead of
                // the first sweep's use, from L2 into this SM's L1 (the later sweeps then hit in L1 too)
                if (l1pf_ptr) {
#ifdef RBP_L1PF_LOAD
                    // a real (never consumed) load: allocates the line in L1 for certain
                    unsigned dummy;
                    asm volatile("ld.global.nc.L1::evict_last.b32 %0, [%1];" : "=r"(dummy) : "l"(l1pf_ptr));
#else
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(l1pf_ptr));
#endif
                    l1pf_ptr += pitch * (long long)sizeof(T);
                }
                // ---- the row of the NEXT step (its producer wrote it in the previous step / the bulk copy was
                // requested DI-1 steps ago): nothing waits on it until the next step --------------------------------
                fetch_row(uin[s_nxt], (rr + 1) & 1);
                // the one cross-lane operand of each stage-row, requested before anything depends on it
                T e1, e2;
                if (p == 0) {
                    e1 = __shfl_up_sync(0xffffffffu, uin[s_m1][qe], 1);
                    e2 = __shfl_up_sync(0xffffffffu, unw[s_m2][qe], 1);
                } else {
                    e1 = __shfl_down_sync(0xffffffffu, uin[s_m1][qe], 1);
                    e2 = __shfl_down_sync(0xffffffffu, unw[s_m2][qe], 1);
                }
                const bool act1 = row_active(row_l - 1), cnt1 = row_owned(row_l - 1);
                // ---- operands: row L-1+OPF is loaded now; row L-2 was loaded one step earlier ----------
                {
                    constexpr int d = -1 + OPF;  // row offset loaded in this step
                    const bool need = OPF ? row_active(row_l) : act1;                       // that row will be relaxed
                    const bool need_v = (unsigned)(row_l + d - ra) <= (unsigned)(re - ra);  // gv row also serves the row below
                    if (SRC && need) {
#pragma unroll
                        for (int v = 0; v < NV; ++v)
                            ptx::ldg16(&sv[(rr + d + SW * 6) % SW][v * PER], psrc + (long long)d * pitch + v * PER);
                    }
                    if (NOISE && need) {
#pragma unroll
                        for (int v = 0; v < NV; ++v)
                            ptx::ldg16(&gh[(rr + d + GHW * 6) % GHW][v * PER], pgh + (long long)d * pitch + v * PER);
                        ghw[(rr + d + GHW * 6) % GHW] = __ldg(pgh + (long long)d * pitch - 1);
                    }
                    if (NOISE && need_v) {
#pragma unroll
                        for (int v = 0; v < NV; ++v)
                            ptx::ldg16(&gv[(rr + d + GVW * 6) % GVW][v * PER], pgv + (long long)d * pitch + v * PER);
                    }
                }
                // ---- the two half-sweep stages of this step -----------------------------------------------------
                //   colour 0 of row L-1: all four neighbours still hold the previous sweep's values (uin);
                //   colour 1 of row L-2: every neighbour already holds this sweep's colour-0 value (unw).
                // FAST = both rows are relaxed and counted and no column of this warp is a fixed ring column: the
                // steady state, free of selects and branches.  Otherwise the generic code (predicated per node).
                auto stages = [&](auto fast_tag) {
                    constexpr bool FAST = decltype(fast_tag)::value;
#pragma unroll
                    for (int q = p; q < C; q += 2) {
                        const T lf = (q > 0) ? uin[s_m1][q - 1] : e1;
                        const T rt = (q < C - 1) ? uin[s_m1][q + 1] : e1;
                        T gn = (T)1, gs = (T)1, gw = (T)1, ge = (T)1, sq = (T)0;
                        if (NOISE) {
                            gn = gv[(rr - 2 + GVW * 6) % GVW][q];
                            gs = gv[(rr - 1 + GVW * 6) % GVW][q];
                            gw = (q > 0) ? gh[(rr - 1 + GHW * 6) % GHW][q - 1] : ghw[(rr - 1 + GHW * 6) % GHW];
                            ge = gh[(rr - 1 + GHW * 6) % GHW][q];
                        }
                        if (SRC) sq = sv[(rr - 1 + SW * 6) % SW][q];
                        const T tq = rb_partial<T, NOISE, SRC>(uin[s_m2][q], lf, rt, gn, gw, ge, sq);
                        T nw = rb_finish<T>(tq, rb_south_weight<T, NOISE>(gs), uin[s_l][q]);
                        if (!FAST) nw = (act1 && ((colmask >> q) & 1u)) ? nw : uin[s_m1][q];
                        unw[s_m1][q] = nw;
                        const T ch = pipe_abs(Ar<T>::sub(nw, uin[s_m1][q]));
                        if ((FAST || cnt1) && ch > mx) mx = ch;  // NaN never wins, like the strict '>' of solver.py:277
                    }
#pragma unroll
                    for (int q = p; q < C; q += 2) {
                        const T lf = (q > 0) ? unw[s_m2][q - 1] : e2;
                        const T rt = (q < C - 1) ? unw[s_m2][q + 1] : e2;
                        T gn = (T)1, gs = (T)1, gw = (T)1, ge = (T)1, sq = (T)0;
                        if (NOISE) {
                            gn = gv[(rr - 3 + GVW * 6) % GVW][q];
                            gs = gv[(rr - 2 + GVW * 6) % GVW][q];
                            gw = (q > 0) ? gh[(rr - 2 + GHW * 6) % GHW][q - 1] : ghw[(rr - 2 + GHW * 6) % GHW];
                            ge = gh[(rr - 2 + GHW * 6) % GHW][q];
                        }
                        if (SRC) sq = sv[(rr - 2 + SW * 6) % SW][q];
                        const T tq = rb_partial<T, NOISE, SRC>(unw[s_m3][q], lf, rt, gn, gw, ge, sq);
                        T nw = rb_finish<T>(tq, rb_south_weight<T, NOISE>(gs), unw[s_m1][q]);
                        if (!FAST) nw = (act2 && ((colmask >> q) & 1u)) ? nw : uin[s_m2][q];
                        unw[s_m2][q] = nw;
                        const T ch = pipe_abs(Ar<T>::sub(nw, uin[s_m2][q]));
                        if ((FAST || cnt2) && ch > mx) mx = ch;
                    }
                };
                if (act1 && act2 && cnt1 && cnt2 && !masked) stages(std::true_type{});
                else stages(std::false_type{});
                // ---- row L-2 has passed both colours of this sweep: hand it on ---------------------------
                if (!last) {
                    if (emit_lane) {
#pragma unroll
                        for (int v = 0; v < NV; ++v)
                            ptx::sts16(out_base + (unsigned)(rr & 1) * SLOT + v * (16u * NG), &unw[s_m2][v * PER]);
                    }
                } else if (lane_valid && cnt2) {
#pragma unroll
                    for (int v = 0; v < NV; ++v) ptx::stg16(pout + v * PER, &unw[s_m2][v * PER]);
                }
                act2 = act1;
                cnt2 = cnt1;
                if (SRC) psrc += pitch;
                if (NOISE) {
                    pgh += pitch;
                    pgv += pitch;
                }
                pout += pitch;
                ++row_l;
#ifndef RBP_NOSYNC  // (timing experiment only: without the barrier the results are wrong)
                __syncthreads();
#endif
            });
        }
        tma_n += (unsigned)n_steps + 1;
        if (lane_valid && mx > mxk) mxk = mx;
    }
    const T v = warp_max(mxk);
    if (lane == 0 && v > (T)0) atomic_max_nonneg(a.resid + k, v);
}

template <int TD, bool NOISE, bool SRC>
int pipe_launch(apde_plan *p, const PipeArgs &a, cudaStream_t st) {
    using Geo_ = PipeGeo<TD, NOISE>;
    static int per_sm_dev[64] = {0};  // per device: the shared-memory opt-in belongs to one context
    int &per_sm = per_sm_dev[p->device & 63];
    if (per_sm == 0) {
        int n = 0;
        APDE_CUDA(cudaFuncSetAttribute(rb_pipe_kernel<TD, NOISE, SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo_::SMEM));
        APDE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, rb_pipe_kernel<TD, NOISE, SRC>, Geo_::NTHR, Geo_::SMEM));
        per_sm = n > 0 ? n : 1;
    }
    int grid = std::min(a.n_items, p->sm_count * per_sm);
    if (grid < 1) grid = 1;
    rb_pipe_kernel<TD, NOISE, SRC><<<grid, Geo_::NTHR, Geo_::SMEM, st>>>(a);
    p->launches++;
    return APDE_OK;
}

template <bool NOISE, bool SRC>
int pipe_depth(apde_plan *p, const PipeArgs &a, int td, cudaStream_t st) {
    switch (td) {
        case 1: return pipe_launch<1, NOISE, SRC>(p, a, st);
        case 2: return pipe_launch<2, NOISE, SRC>(p, a, st);
        case 3: return pipe_launch<3, NOISE, SRC>(p, a, st);
        case 4: return pipe_launch<4, NOISE, SRC>(p, a, st);
        case 5: return pipe_launch<5, NOISE, SRC>(p, a, st);
        case 6: return pipe_launch<6, NOISE, SRC>(p, a, st);
    }
    set_error("rb_pipe: unsupported fused depth %d", td);
    return APDE_EINVAL;
}

constexpr int pipe_tile_cols(int td, bool noise) { return 30 * PIPE_C * pipe_w(td, noise) - 2 * pipe_hc(td); }

}  // namespace

// Runs n_sweeps sweeps as passes of <= temporal_depth fused sweeps; ping-pongs u[cur] -> u[1-cur].
// phase 1 = only the rows next to a neighbour strip (their output is what the halo exchange sends),
// phase 2 = the remaining rows, phase 0 = all.  With phases the caller runs a single pass.
int RBP_FN(apde_plan *p, int64_t n_sweeps, int64_t sweep_base, int phase, cudaStream_t st) {
    const Geo g = make_geo(p);
    if (g.cols < 3 || g.lrows < 3 || n_sweeps == 0) return APDE_OK;
    const int max_td = std::min(p->d.temporal_depth, 6);
    if (phase != 0 && n_sweeps > max_td) {
        set_error("rb_pipe: phased runs take one pass (n_sweeps <= %d)", max_td);
        return APDE_EINVAL;
    }
    int64_t done = 0;
    while (done < n_sweeps) {
        const int td = (int)std::min<int64_t>(max_td, n_sweeps - done);
        const int src_buf = p->cur, dst_buf = 1 - p->cur;
        PipeArgs a;
        a.in = p->U<T>(src_buf);
        a.out = p->U<T>(dst_buf);
        a.src = p->S<T>();
        a.gh = p->GH<T>();
        a.gv = p->GV<T>();
        a.resid = p->resid.as<T>() + sweep_base + done;
        a.stop = p->stop_flag;
        a.pitch = g.pitch;
        a.padl = g.padl;
        a.lrows = (int)g.lrows;
        a.cols = (int)g.cols;
        a.row0 = g.row0;
        const int tile_cols = pipe_tile_cols(td, p->d.has_noise != 0);
        a.n_tiles = (int)((g.cols + tile_cols - 1) / tile_cols);
        const int own_lo = (int)g.own_lo, own_hi = (int)g.own_hi, owned = own_hi - own_lo;
        // Rows next to a neighbour strip form their own short row blocks (phase 1): at least the `halo` rows the
        // exchange ships, a few more so the block is not all pipeline fill.  The rest (phase 2) is cut into as many
        // row blocks as fill the GPU once with the column tiles -- tall blocks keep the 2*TD-row halo and the
        // pipeline fill/drain negligible.
        const bool top_nb = p->ghost_top > 0, bot_nb = p->ghost_bot > 0;
        int edge = (int)std::max<int64_t>(p->halo, p->edge_rows);
        int e_top = top_nb ? std::min(edge, owned) : 0;
        int e_bot = bot_nb ? std::min(edge, owned - e_top) : 0;
        const int mid_lo = own_lo + e_top, mid_hi = own_hi - e_bot, mid = mid_hi - mid_lo;
        int rbh;
        if (p->row_block_forced) {
            rbh = (int)p->row_block;  // APDE_ROW_BLOCK (tests, tuning)
        } else {
            const int slots = p->sm_count * pipe_ctas(td, p->d.has_noise != 0);
            const int nrb = std::max(1, slots / std::max(a.n_tiles, 1));
            rbh = std::max((mid + nrb - 1) / nrb, 64);
        }
        rbh = std::max(rbh, (int)p->halo);
        auto nblk = [&](int rows) { return rows > 0 ? (rows + rbh - 1) / rbh : 0; };
        int rc = APDE_OK;
        auto launch = [&](int a_lo, int a_hi, int b_lo, int b_hi) {
            a.a_lo = a_lo; a.a_hi = a_hi; a.b_lo = b_lo; a.b_hi = b_hi;
            a.rbh = rbh;
            a.nblk_a = nblk(a_hi - a_lo);
            a.nblk_b = nblk(b_hi - b_lo);
            a.n_items = (a.nblk_a + a.nblk_b) * a.n_tiles;
            if (a.n_items <= 0) return;
            const bool N = p->d.has_noise, S = p->d.has_source;
            rc = N ? (S ? pipe_depth<true, true>(p, a, td, st) : pipe_depth<true, false>(p, a, td, st))
                   : (S ? pipe_depth<false, true>(p, a, td, st) : pipe_depth<false, false>(p, a, td, st));
        };
        if (phase == 0) {
            // edge blocks first (they are short), then the interior: one launch, two row ranges + the middle
            if (e_top + e_bot > 0) launch(own_lo, mid_lo, mid_hi, own_hi);
            if (rc == APDE_OK) launch(mid_lo, mid_hi, 0, 0);
        } else if (phase == 1) {
            launch(own_lo, mid_lo, mid_hi, own_hi);
        } else {
            launch(mid_lo, mid_hi, 0, 0);
        }
        if (rc != APDE_OK) return rc;
        if (phase == 1) {
            p->pass_open = true;  // buffers flip when phase 2 completes the pass
        } else {
            p->cur = dst_buf;
            p->pass_open = false;
        }
        done += td;
    }
    APDE_CUDA(cudaGetLastError());
    return APDE_OK;
}

}  // namespace apde

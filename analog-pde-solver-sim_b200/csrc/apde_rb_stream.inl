// apde_rb_stream.inl -- temporally blocked red-black sweep, included by apde_rb_stream_f{64,32}.cu
// with RBS_T (element type) and RBS_FN (exported launcher name) defined.
//
// out = RB^TD(in): TD full red-black sweeps fused in ONE pass over HBM (u read once, written once;
// source / conductances read once from HBM and re-read from L1/L2 by the later stages).
//
// Work item = (row block, column strip), one item per warp at a time, no shared memory and no
// block barriers: a warp streams DOWN the rows of its strip keeping a sliding window of rows in
// registers.  2*TD half-sweep stages are pipelined one row apart:
//     iteration r : load row r+PF;  stage s relaxes colour (s&1) of row r-1-s  (s = 0..2TD-1);
//                   row r-2TD has passed every stage -> stored.
// Stage s of row i needs rows i-1,i,i+1 after stage s-1: row i+1 got it earlier in the same
// iteration, rows i-1 and i in the previous ones; a stage only writes its own colour and only
// reads the other colour, so updating the window in place is exact.
// Lane l owns C adjacent columns (one 16-byte vector); vertical neighbours are the lane's own
// registers, horizontal neighbours are its own registers or ONE warp shuffle per stage-row.
// With the stream start aligned to an even global row, the column parity updated in iteration r
// is (r+1)&1 for every stage, so all register indices are compile-time constants.
// Redundant halo: 2TD rows above/below the row block and HCA >= 2TD columns left/right of the strip
// are recomputed (garbage creeps in one node per stage from the open edges and never reaches
// the stored region).  The Dirichlet ring and the padding are never updated.
#include <algorithm>
#include <type_traits>

#include "apde_rb_common.cuh"

namespace apde {
namespace {

using T = RBS_T;

#ifndef RBS_BLOCK
#define RBS_BLOCK 128
#endif
constexpr int RBS_THREADS = RBS_BLOCK;  // no block-level cooperation: any multiple of 32 works
#ifndef RBS_PF_ROWS
#define RBS_PF_ROWS 2
#endif
constexpr int RBS_PF = RBS_PF_ROWS;  // rows prefetched ahead (kept even so the unroll period stays even)
#ifndef RBS_LAG
#define RBS_LAG 1
#endif
// Rows between consecutive half-sweep stages.  1: stage s+1 consumes what stage s produced in the
// same iteration (one long dependent chain, smallest register window).  2: every stage of an
// iteration is independent of the others (NS-way instruction-level parallelism, NS more window rows).
constexpr int LAG = RBS_LAG;

struct StreamArgs {
    const T *in;
    T *out;
    const T *src, *gh, *gv;  // colour-split operand planes (apde_plan.h): even columns, then odd columns half a pitch on
    T *resid;  // slot of the first sweep of this pass
    const int *stop;  // optional device flag: set => the solve has converged, do nothing
    long long pitch, padl, lrows, cols, grows, row0;
    long long own_lo, own_hi;      // the local rows this launch cuts into row blocks (a strip, a strip-edge or interior region, a band)
    long long rbh;                 // row-block height
    long long blk_a0, blk_an;      // row blocks [blk_a0, blk_a0+blk_an) ...
    long long blk_b0, blk_bn;      // ... and [blk_b0, blk_b0+blk_bn)
    long long n_strips, n_items;
};

template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

#ifndef RBS_OPF_ROWS
#define RBS_OPF_ROWS 3
#endif
constexpr int RBS_OPF = RBS_OPF_ROWS;  // iterations an operand row is prefetched into L1 before its first use
#ifndef RBS_L2PF
#define RBS_L2PF 12
#endif
constexpr int L2PF_ROWS = RBS_L2PF;  // rows ahead that are pulled from HBM into L2 (source-only kernels)
// Resident blocks per SM the compiler must allow (caps registers).  Measured on B200 (tools/rb_sweep.py):
// the fp32 source-only kernels gain from 12 warps/SM (168 registers, no spills); every other variant
// loses more to spills than it gains from occupancy.
constexpr int rbs_min_blocks(int elem_bytes, int lane_cols, bool noise) {
#ifdef RBS_MINBLOCKS_ALL
    return RBS_MINBLOCKS_ALL;
#elif defined(RBS_F32_MINBLOCKS)
    return (elem_bytes == 4 && lane_cols == 4 && !noise) ? RBS_F32_MINBLOCKS : 1;
#else
#ifndef RBS_F64N_MINBLOCKS
#define RBS_F64N_MINBLOCKS 1
#endif
    if (elem_bytes == 8 && lane_cols == 2 && noise) return RBS_F64N_MINBLOCKS;
    return (elem_bytes == 4 && lane_cols == 4 && !noise) ? 3 : 1;
#endif
}

// Shifted or rotating register window (see the kernel).  Measured on B200, 16384^2 (profiles/r2_rb_sweep_shift.txt):
// shifted wins wherever registers were the wall -- fp64 Poisson depth 4: 548 vs 496 GLUP/s (rotating, depth 3), fp64
// with mismatch 215 vs 210, fp32 with mismatch 462 vs 441 -- and loses where they were not: fp32 Poisson 1004-1068
// (depth 4-6) vs 1076 (rotating, depth 3, 168 registers).
#ifndef RBS_SHIFT_MODE
#define RBS_SHIFT_MODE -1  // -1: per variant (default), 0: always rotating, 1: always shifted
#endif
#ifndef RBS_OPS
#define RBS_OPS 2  // 0: no asynchronous shared-memory feed, 1: kernels with conductances, 2: every kernel
#endif
// asynchronous shared-memory feed (described below): needs one 16-byte vector per lane
__host__ __device__ constexpr bool rbs_ops(int elem_bytes, int lane_cols, bool noise) {
#ifdef RBS_TMA
    return false;
#else
    // one 16-byte vector per lane, or two without conductances (their six segments per row would not fit)
    return (RBS_OPS == 2 || (RBS_OPS == 1 && noise)) && (lane_cols * elem_bytes == 16 || (lane_cols * elem_bytes == 32 && !noise)) &&
           RBS_SHIFT_MODE != 0;
#endif
}
__host__ __device__ constexpr bool rbs_shift(int elem_bytes, int lane_cols, bool noise) {
    return RBS_SHIFT_MODE >= 0 ? RBS_SHIFT_MODE != 0 : (elem_bytes == 8 || noise || rbs_ops(elem_bytes, lane_cols, noise));
}

// Operand planes read by the kernel: colour-split copies (apde_plan.h) or the natural ones.  Split wins wherever it was
// measured except fp32 Poisson, whose 168-register variant spills with the vector operand loads (887 vs 1040 GLUP/s).
#ifndef RBS_SPLIT_MODE
#define RBS_SPLIT_MODE -1  // -1: per variant (default), 0: natural everywhere, 1: split everywhere
#endif
__host__ __device__ constexpr bool rbs_split(int elem_bytes, int lane_cols, bool noise) {
    return rbs_ops(elem_bytes, lane_cols, noise) || (RBS_SPLIT_MODE >= 0 ? RBS_SPLIT_MODE != 0 : (elem_bytes == 8 || noise));
}

// Asynchronous feed through shared memory (kernels with conductances, one 16-byte vector per lane).  Every iteration
// the warp copies, with cp.async (16 bytes per lane per instruction, no register and no global-load return path
// involved), the operand row stage 0 will need RBS_OPS_AHEAD iterations later -- even / odd halves of source, gh and
// gv, 272 bytes each, contiguous in a ring slot -- plus the next u row of the strip; one commit group per row,
// cp.async.wait_group + __syncwarp before use.  All stages read their operands with ld.shared at constant offsets
// from NS + 1 slot addresses, and the u row enters the register window straight from shared memory.
// Why: with register loads half of all warp cycles sat on the first use after each batch of loads (ncu source view):
// the first touches of a row come from DRAM and every later load of the warp returns behind them.
#ifndef RBS_OPS_AHEAD
#define RBS_OPS_AHEAD 4
#endif
// bytes per half-row segment: the strip's 32 * C / 2 elements + 16 of alignment slack / west link (vb = bytes per lane per u row)
__host__ __device__ constexpr int rbs_ops_seg(int vb) { return 16 * vb + 16; }
__host__ __device__ constexpr int rbs_ops_rows(int td) { return 2 * td + 1 + RBS_OPS_AHEAD; }
__host__ __device__ constexpr int rbs_ops_row_bytes(int vb, bool src, bool noise) { return ((src ? 2 : 0) + (noise ? 4 : 0)) * rbs_ops_seg(vb); }
__host__ __device__ constexpr int rbs_ops_warp_bytes(int vb, int td, bool src, bool noise) {
    return (rbs_ops_rows(td) * rbs_ops_row_bytes(vb, src, noise) + (RBS_OPS_AHEAD + 1) * 32 * vb + 127) / 128 * 128;
}
__device__ __forceinline__ void cp_async16(unsigned dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void prefetch_l1(const void *p) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}
// |x| on the bit pattern (integer pipe) instead of an fp64-pipe instruction
__device__ __forceinline__ double abs_bits(double x) {
    return __hiloint2double(__double2hiint(x) & 0x7fffffff, __double2loint(x));
}
__device__ __forceinline__ float abs_bits(float x) { return fabsf(x); }

// 16-byte vector loads/stores typed as the element type: the values land directly in (and leave
// directly from) the registers of the window -- no int->fp repacking moves that would make the warp
// wait for the load right after issuing it.
__device__ __forceinline__ void ld_vec16(double *w, const double *p) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(w[0]), "=d"(w[1]) : "l"(p));
}
__device__ __forceinline__ void ld_vec16(float *w, const float *p) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(w[0]), "=f"(w[1]), "=f"(w[2]), "=f"(w[3])
                 : "l"(p));
}
__device__ __forceinline__ void st_vec16(double *p, const double *w) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(w[0]), "d"(w[1]) : "memory");
}
__device__ __forceinline__ void st_vec16(float *p, const float *w) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(w[0]), "f"(w[1]),
                 "f"(w[2]), "f"(w[3])
                 : "memory");
}

// ---- optional bulk-async (TMA, cp.async.bulk) feed of the u rows: per warp a FIFO ring of row
// segments in shared memory, filled RBS_TMA_AHEAD rows ahead by one elected lane, completion tracked by
// one mbarrier per slot.  Keeps many more bytes in flight per SM than the register prefetch can.
#ifndef RBS_TMA_SLOTS
#define RBS_TMA_SLOTS 8
#endif
#ifndef RBS_TMA_AHEAD
#define RBS_TMA_AHEAD 6
#endif
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void lds_vec16(double *w, unsigned addr) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(w[0]), "=d"(w[1]) : "r"(addr));
}
__device__ __forceinline__ void lds_vec16(float *w, unsigned addr) {
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(w[0]), "=f"(w[1]), "=f"(w[2]), "=f"(w[3]) : "r"(addr));
}
template <int C>
__device__ __forceinline__ void lds_row(T (&w)[C], unsigned addr) {
    constexpr int PER = 16 / (int)sizeof(T);
#pragma unroll
    for (int k = 0; k < C / PER; ++k) lds_vec16(&w[k * PER], addr + 16 * k);
}


// Operand load that stays where it is written: a volatile asm is not sunk towards its first use, so
// an iteration's operand loads really are in flight before the arithmetic starts.
__device__ __forceinline__ double ld_operand(const double *p) {
#ifdef RBS_PINNED_OPERAND_LOADS
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ float ld_operand(const float *p) {
#ifdef RBS_PINNED_OPERAND_LOADS
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

// N consecutive operands of one colour (colour-split planes).  ALIGNED: p is aligned to N elements -> one vector load.
__device__ __forceinline__ void ld_operand_vec2(double *o, const double *p) {
    asm("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(o[0]), "=d"(o[1]) : "l"(p));
}
__device__ __forceinline__ void ld_operand_vec2(float *o, const float *p) {
    asm("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(o[0]), "=f"(o[1]) : "l"(p));
}
template <int N, bool ALIGNED>
__device__ __forceinline__ void ld_operands(T *o, const T *p) {
    if constexpr (ALIGNED && N % 2 == 0) {
#pragma unroll
        for (int k = 0; k < N; k += 2) ld_operand_vec2(&o[k], p + k);
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) o[k] = ld_operand(p + k);
    }
}

__device__ __forceinline__ void lds_1(double *o, unsigned a) { asm volatile("ld.shared.f64 %0, [%1];" : "=d"(o[0]) : "r"(a)); }
__device__ __forceinline__ void lds_1(float *o, unsigned a) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o[0]) : "r"(a)); }
__device__ __forceinline__ void lds_2(double *o, unsigned a) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(o[0]), "=d"(o[1]) : "r"(a));
}
__device__ __forceinline__ void lds_2(float *o, unsigned a) {
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(o[0]), "=f"(o[1]) : "r"(a));
}
template <int N, bool ALIGNED>
__device__ __forceinline__ void lds_operands(T *o, unsigned a) {
    if constexpr (ALIGNED && N % 2 == 0) {
#pragma unroll
        for (int k = 0; k < N; k += 2) lds_2(&o[k], a + k * (unsigned)sizeof(T));
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) lds_1(&o[k], a + k * (unsigned)sizeof(T));
    }
}

template <int C>
__device__ __forceinline__ void load_row(T (&w)[C], const T *p) {
    constexpr int PER = 16 / (int)sizeof(T);
    static_assert(C % PER == 0, "lane vector must be a multiple of 16 bytes");
#pragma unroll
    for (int k = 0; k < C / PER; ++k) ld_vec16(&w[k * PER], p + k * PER);
}

template <int C>
__device__ __forceinline__ void store_row(T *p, const T (&w)[C]) {
    constexpr int PER = 16 / (int)sizeof(T);
#pragma unroll
    for (int k = 0; k < C / PER; ++k) st_vec16(p + k * PER, &w[k * PER]);
}

template <int C, int TD, bool NOISE, bool SRC>
__global__ void __launch_bounds__(RBS_THREADS, rbs_min_blocks((int)sizeof(T), C, NOISE)) rb_stream_kernel(const StreamArgs a) {
    constexpr int NS = 2 * TD;             // half-sweep stages
    constexpr int DEPTH_ROWS = 1 + LAG * (NS - 1);            // newest-to-final distance in rows
    constexpr int PF = rbs_ops((int)sizeof(T), C, NOISE) ? 0 : RBS_PF;  // u rows loaded ahead of use (none with the OPS feed)
    constexpr int NW = ((DEPTH_ROWS + 2 + 1) / 2) * 2 + PF;  // register window rows (even)
    // Two ways to keep the sliding window in registers (SHIFT chosen per variant, rbs_shift()):
    //   rotating: row r+off lives in slot (off + rr) mod NW; nothing ever moves, but the loop body must be unrolled
    //             NW times for the slot numbers to be constants -- 6600 instructions (103 KB) for fp64 depth 3, far
    //             beyond the instruction cache (ncu: no_instruction 0.55 stall cycles per issue);
    //   shifted : row r+off lives in slot off + DEPTH_ROWS + 1 + (rr & 1) and the whole window moves down two slots
    //             every second iteration: unrolled 2x (1744 instructions), (NW-1)*C register moves per two iterations,
    //             and ~35 fewer live registers -- depth 4 then fits without spills (fp64 Poisson 230 registers).
    constexpr bool SHIFT = rbs_shift((int)sizeof(T), C, NOISE);
    constexpr int WSL = SHIFT ? NW + 1 : NW;  // slots
    constexpr int UNR = SHIFT ? 2 : NW;       // unroll period
    static_assert(!SHIFT || LAG == 1, "the shifted window assumes one row between stages");
#ifndef RBS_FT
#define RBS_FT 0
#endif
    constexpr bool SPLIT = rbs_split((int)sizeof(T), C, NOISE);
    constexpr bool OPS = rbs_ops((int)sizeof(T), C, NOISE);
    constexpr bool FT = RBS_FT && SHIFT && NOISE && SPLIT && !OPS;
    // OPS rings per warp: OR operand row slots of OROW bytes ([src e|o] gh e|o gv e|o, 272 bytes each), then OU u-row slots
    constexpr int VB = C * (int)sizeof(T);  // bytes per lane per u row
    constexpr int OA = RBS_OPS_AHEAD, OR = rbs_ops_rows(TD), OROW = rbs_ops_row_bytes(VB, SRC, NOISE), OSEG = rbs_ops_seg(VB);
    constexpr int OUB = 32 * VB, OSC = OSEG / 16;  // bytes per u-row slot, 16-byte chunks per segment
    constexpr int OU = OA + 1, OCH = OROW / 16, OCP = (OCH + 31) / 32;  // 16-byte chunks per operand row, copies per lane
    constexpr int O_GH = SRC ? 2 * OSEG : 0, O_GV = O_GH + 2 * OSEG;
    constexpr int OCPA = OCP > 0 ? OCP : 1;
    static_assert(!OPS || VB == 16 || VB == 32, "OPS: one or two 16-byte vectors per lane");
#define RBS_SLOT(off) (SHIFT ? ((off) + DEPTH_ROWS + 1 + (rr & 1)) : (((off) + rr + 8 * NW) % NW))
    constexpr int CW = 32 * C;             // columns per warp strip
    constexpr int HCA = ((NS + C - 1) / C) * C;  // column halo, lane aligned
    constexpr int VW = CW - 2 * HCA;       // columns a strip stores
    static_assert(VW > 0 && NW % 2 == 0, "bad tile");
    // prefetch distances, measured per variant on B200 (tools/rb_sweep.py): with conductances the extra
    // prefetch instructions cost more than they hide
#ifndef RBS_NOISE_L2PF
#define RBS_NOISE_L2PF 0
#endif
    constexpr int L2PF = NOISE ? RBS_NOISE_L2PF : L2PF_ROWS;
    constexpr int OPF = NOISE ? RBS_OPF : 2 * RBS_OPF;

    if (a.stop && *a.stop) return;  // device-side stop: a sweep after the converged check interval is a no-op
    const int lane = threadIdx.x & 31;
    const long long warp_id = (long long)blockIdx.x * (RBS_THREADS / 32) + (threadIdx.x >> 5);
    const long long n_warps = (long long)gridDim.x * (RBS_THREADS / 32);
    const long long pitch = a.pitch, half = a.pitch >> 1;

    T mx[TD];
#pragma unroll
    for (int t = 0; t < TD; ++t) mx[t] = (T)0;

    extern __shared__ __align__(128) unsigned char rbs_smem[];
    unsigned o_ring = 0, o_uring = 0;  // shared-memory addresses of this warp's operand ring and u-row ring
    unsigned o_islot = 0, o_wslot = 0, o_iu = 0, o_wu = 0;  // next operand / u slot to fill, and to consume
    if constexpr (OPS) {
        o_ring = (unsigned)__cvta_generic_to_shared(rbs_smem) + (unsigned)(threadIdx.x >> 5) * (unsigned)rbs_ops_warp_bytes(C * (int)sizeof(T), TD, SRC, NOISE);
        o_uring = o_ring + OR * OROW;
    }
#ifdef RBS_TMA
    constexpr int TS = RBS_TMA_SLOTS, TA = RBS_TMA_AHEAD;
    constexpr unsigned ROWB = 32u * C * (unsigned)sizeof(T);  // bytes of one warp row segment
    static_assert(TA + 1 <= TS, "TMA ring too shallow");
    const unsigned wsm = (unsigned)__cvta_generic_to_shared(rbs_smem) + (unsigned)(threadIdx.x >> 5) * (TS * ROWB + 8 * TS);
    const unsigned bars = wsm + TS * ROWB;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < TS; ++k) mbar_init(bars + 8 * k, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned issued = 0, consumed = 0;  // FIFO counters over the whole kernel (slot = n % TS, parity = (n / TS) & 1)
#endif
    for (long long item = warp_id; item < a.n_items; item += n_warps) {
        const long long bsel = item / a.n_strips;
        const long long sidx = item - bsel * a.n_strips;
        const long long blk = bsel < a.blk_an ? a.blk_a0 + bsel : a.blk_b0 + (bsel - a.blk_an);
        // local row indices fit 32 bits; only element offsets are 64-bit (carried in the pointers)
        const int rb0 = (int)(a.own_lo + blk * a.rbh);
        const int rb1 = (int)min((long long)rb0 + a.rbh, a.own_hi);
        const long long jl0 = sidx * VW - HCA + (long long)C * lane;  // first column of this lane (even)
        const bool lane_valid = (C * lane >= HCA) && (C * lane < CW - HCA) && (jl0 < a.cols);
        unsigned colmask = 0;  // bit q: column jl0+q is an interior column
#pragma unroll
        for (int q = 0; q < C; ++q)
            if (jl0 + q >= 1 && jl0 + q <= a.cols - 2) colmask |= 1u << q;

        const int ra = max(rb0 - NS, 0);                          // first streamed row (never relaxed)
        const int re = (int)min((long long)rb1 + NS, a.lrows) - 1;  // last streamed row (never relaxed)
        const int rs = ra - (int)((a.row0 + ra) & 1);             // even global row: static colour parity
        const int r_last = rb1 - 1 + DEPTH_ROWS;
        // rows in (ra, re) are relaxed (the global ring rows and the outermost ghost rows always sit
        // at ra / re, never inside); rows in [rb0, rb1) are stored and counted
        const unsigned upd_span = (unsigned)max(re - ra - 1, 0), own_span = (unsigned)(rb1 - rb0);

        const long long lane_off = a.padl + jl0;
        const T *pin = a.in + (long long)rs * pitch + lane_off;            // row r
        T *pout = a.out + (long long)(rs - DEPTH_ROWS) * pitch + lane_off; // row r - DEPTH_ROWS
        // operands: row r - 1 (stage 0's row) of the colour-split planes, at the lane's first column pair
        const long long st0 = (long long)(rs - 1) * pitch + (SPLIT ? ((a.padl + jl0) >> 1) : lane_off);
        const T *psrc = a.src + st0, *pgh = a.gh + st0, *pgv = a.gv + st0;
        // OPS: the strip's half-row segment starts at element o_m0 of a split half row, rounded down to 16 bytes and one
        // element further left (the west link of the strip's first even node)
        constexpr int EPV = 16 / (int)sizeof(T);
        const long long o_m0 = (a.padl + sidx * VW - HCA) >> 1;
        const long long o_ms = (o_m0 - 1) & ~(long long)(EPV - 1);
        const unsigned o_lane = (unsigned)((o_m0 - o_ms) + (C / 2) * lane) * (unsigned)sizeof(T);  // this lane's byte offset in a segment
        // rows fed: stage 0's row of every iteration, rs - 1 .. o_last, in order, each with the u row below it
        const int o_last = rs + ((r_last - rs) / UNR + 1) * UNR - 2;
        int o_next = rs - 1;
        long long o_goff = (long long)(rs - 1) * pitch;  // element offset of the next operand row in a plane
        const T *o_src[OPS ? OCPA : 1];                  // chunk 32k + lane of an operand row: address in row 0
        if constexpr (OPS) {
#pragma unroll
            for (int k = 0; k < OCP; ++k) {
                const int c = min(32 * k + lane, OCH - 1), seg = c / OSC, part = c - seg * OSC;
                const int arr = seg >> 1;
                const T *plane = SRC ? (arr == 0 ? a.src : arr == 1 ? a.gh : a.gv) : (arr == 0 ? a.gh : a.gv);  // (no NOISE: src only)
                o_src[k] = plane + ((seg & 1) ? half : 0) + o_ms + part * EPV;
            }
        }
        auto ops_feed = [&]() {
            if (o_next <= o_last) {
                const unsigned dst = o_ring + o_islot * OROW + 16u * lane;
#pragma unroll
                for (int k = 0; k < OCP; ++k)
                    if (32 * k + lane < OCH) cp_async16(dst + 512u * k, o_src[k] + o_goff);
#pragma unroll
                for (int k = 0; k < VB / 16; ++k)  // the lane's own vectors of the u row
                    cp_async16(o_uring + o_iu * (unsigned)OUB + (unsigned)VB * lane + 16u * k, a.in + o_goff + pitch + lane_off + k * EPV);
                o_islot = (o_islot + 1 == OR) ? 0 : o_islot + 1;
                o_iu = (o_iu + 1 == OU) ? 0 : o_iu + 1;
                o_goff += pitch;
                ++o_next;
            }
            cp_async_commit();  // always: the group count per iteration stays fixed
        };
        unsigned oa[OPS ? NS + 1 : 1];  // oa[k]: this lane's address in the ring slot of operand row (stage 0's row) - k
        if constexpr (OPS) {
#pragma unroll
            for (int k = 0; k <= NS; ++k) oa[k] = o_ring + o_lane;
            __syncwarp();  // every lane is done with the previous item's slots
#pragma unroll 1
            for (int k = 0; k < OA; ++k) ops_feed();
        }
#ifdef RBS_PF_FEW_LANES
        // one prefetch per 128-byte line instead of one per lane: lane l < PFL takes line (l >> 1) of half (l & 1)
        constexpr int PFL = 2 * (32 * (C / 2) * (int)sizeof(T) / 128 + 1);
        const bool pf_lane = lane < PFL;
        const long long pfo = ((lane & 1) ? half : 0) + (long long)(lane >> 1) * (128 / (int)sizeof(T)) - (C / 2) * lane;
#else
        constexpr bool pf_lane = true;
        // prefetches of split rows: even lanes cover the even-column half, odd lanes the other
        const long long pfo = (SPLIT && (lane & 1)) ? half : 0;
#endif

        T w[WSL][C];
#pragma unroll
        for (int k = 0; k < WSL; ++k)
#pragma unroll
            for (int q = 0; q < C; ++q) w[k][q] = (T)0;
#ifdef RBS_TMA
        // rows rs .. r_load_end are fed in order: issue TA rows ahead of their consumption
        // exactly the rows the item consumes: RBS_PF in the prologue + one per iteration of the chunked loop
        // (nothing may still be in flight towards this CTA's shared memory when the item -- or the kernel -- ends)
        const int r_load_end = rs + PF + ((r_last - rs) / UNR + 1) * UNR - 1;
        const T *pin0 = a.in + (long long)rs * pitch + (a.padl + sidx * VW - HCA);  // lane 0's address of row rs
        int next_issue = rs;
        auto tma_issue_upto = [&](int row_limit) {
            // warp-uniform loop; lane 0 arms the barrier and launches the copy
            while (next_issue <= row_limit && next_issue <= r_load_end) {
                if (lane == 0) {
                    const unsigned slot = issued % TS;
                    mbar_expect_tx(bars + 8 * slot, ROWB);
                    tma_load_1d(wsm + slot * ROWB, pin0 + (long long)(next_issue - rs) * pitch, ROWB, bars + 8 * slot);
                }
                ++issued;
                ++next_issue;
            }
        };
        auto tma_consume = [&](T (&dst)[C]) {
            const unsigned slot = consumed % TS, par = (consumed / TS) & 1u;
            mbar_wait(bars + 8 * slot, par);
            lds_row<C>(dst, wsm + slot * ROWB + (unsigned)lane * (C * (unsigned)sizeof(T)));
            ++consumed;
        };
        __syncwarp();
        tma_issue_upto(rs + PF - 1 + TA);
#pragma unroll
        for (int k = 0; k < PF; ++k) tma_consume(w[k]);
#else
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            load_row<C>(w[SHIFT ? k + DEPTH_ROWS + 1 : k], pin + k * pitch);  // rows past either end are zero padding
        }
#endif
        // register pipeline of the first touches (colour-split operands: what stage 0 reads, and stage 1's source, was
        // never read before): loaded one loop trip before use, so their DRAM latency is covered by a trip of work
        constexpr int FTN = FT ? UNR : 1, FTC = FT ? C / 2 : 1;
        T ft_sv[FTN][FTC], ft_gn[FTN][FTC], ft_gs[FTN][FTC], ft_ge[FTN][FTC], ft_gw[FTN][FTC], ft_sv1[FTN][FTC];
        if constexpr (FT) {
#pragma unroll
            for (int rr = 0; rr < UNR; ++rr) {
                const int p = (rr + 1) & 1;
                const long long same = (long long)rr * pitch + (p ? half : 0), other = (long long)rr * pitch + (p ? 0 : half - 1);
                if (SRC) ld_operands<FTC, true>(ft_sv[rr], psrc + same);
                if (SRC) ld_operands<FTC, true>(ft_sv1[rr], psrc + same - pitch);
                ld_operands<FTC, true>(ft_gn[rr], pgv + same - pitch);
                ld_operands<FTC, true>(ft_gs[rr], pgv + same);
                ld_operands<FTC, true>(ft_ge[rr], pgh + same);
                if (p) ld_operands<FTC, true>(ft_gw[rr], pgh + other);
                else ld_operands<FTC, false>(ft_gw[rr], pgh + other);
            }
        }
        unsigned act = 0, cnt = 0;  // bit s: the row stage s handles this iteration is relaxed / counted
        T mxl[TD];
#pragma unroll
        for (int t = 0; t < TD; ++t) mxl[t] = (T)0;

        // A steady-state copy of the loop body without the keep-the-old-value selects, the relaxed / counted bit tests
        // and the store range test (interior strips, rows well inside the block) was built and measured in round 2:
        // 643 instead of 830 instructions per two rows, and the same speed on every variant (fp64 Poisson 662 vs 674,
        // fp32 1286 vs 1279, with conductances 425 / 893 vs 424 / 904 GLUP/s) -- the two resident warps per scheduler
        // are bound by their dependent chains, not by issue slots.  Removed again.
        for (int rc = rs; rc <= r_last; rc += UNR) {
#pragma unroll
            for (int rr = 0; rr < UNR; ++rr) {
                const int r = rc + rr;
                // unconditional (padding rows make every address valid): the load lands straight in
                // its window slot and nothing waits on it until the row is first used
#ifdef RBS_TMA
                __syncwarp();  // every lane is done with the slot that is about to be refilled
                tma_issue_upto(r + PF + TA);
                tma_consume(w[RBS_SLOT(PF)]);
#else
                if constexpr (OPS) {
                    cp_async_wait<OA - 1>();  // this lane's part of the oldest group (operand row r - 1, u row r) has landed
                    __syncwarp();             // ... so has everyone's; and every lane is done with the previous iteration's reads
                    ops_feed();               // refills the slots last read in the previous iteration
                    lds_row<C>(w[RBS_SLOT(PF)], o_uring + o_wu * (unsigned)OUB + (unsigned)VB * lane);
                    o_wu = (o_wu + 1 == OU) ? 0 : o_wu + 1;
#pragma unroll
                    for (int k = NS; k > 0; --k) oa[k] = oa[k - 1];
                    oa[0] = o_ring + o_wslot * OROW + o_lane;
                    o_wslot = (o_wslot + 1 == OR) ? 0 : o_wslot + 1;
                } else {
                    load_row<C>(w[RBS_SLOT(PF)], pin + PF * pitch);
                }
#endif
                act = (act << 1) | (((unsigned)(r - 2 - ra) < upd_span) ? 1u : 0u);
                cnt = (cnt << 1) | (((unsigned)(r - 1 - rb0) < own_span) ? 1u : 0u);
                // column parity relaxed by stage s in this iteration: (rr + 1 + (LAG-1)*s) & 1
                // Operand rows (source / conductances) are pulled into L1 RBS_OPF iterations before
                // their first use, then every stage's operands are loaded at the top of the
                // iteration (L1 hits) so the loads overlap the dependent stage chain.
                if (L2PF > 0 && r + L2PF <= re) {  // not past the item's last row (another warp's block)
                    // two-level software prefetch: HBM -> L2 far ahead (hides DRAM latency at low
                    // occupancy), L2 -> L1 / registers a few rows ahead
#ifndef RBS_TMA
                    prefetch_l2(pin + (long long)L2PF * pitch);
#endif
                    if (pf_lane && !OPS) {
                        if (SRC) prefetch_l2(psrc + (long long)L2PF * pitch + pfo);
                        if (NOISE) {
                            prefetch_l2(pgh + (long long)L2PF * pitch + pfo);
                            prefetch_l2(pgv + (long long)L2PF * pitch + pfo);
                        }
                    }
                }
#ifndef RBS_NO_L1PF
                if (pf_lane && !OPS) {
                    if (SRC) prefetch_l1(psrc + (long long)OPF * pitch + pfo);
                    if (NOISE) {
                        prefetch_l1(pgh + (long long)OPF * pitch + pfo);
                        prefetch_l1(pgv + (long long)OPF * pitch + pfo);
                    }
                }
#endif
                static_assert(C % 2 == 0, "a lane holds whole column pairs");
                constexpr int CH = C / 2;
                // operands of the first LA stages are fetched up front, the rest one stage ahead
                // (all of them when only the source is read; conductances would cost too many registers)
#ifndef RBS_NOISE_LA
#define RBS_NOISE_LA 1
#endif
                constexpr int LA = NOISE ? RBS_NOISE_LA : NS;
                T o_sv[NS][CH], o_gn[NOISE ? NS : 1][CH], o_gs[NOISE ? NS : 1][CH], o_gw[NOISE ? NS : 1][CH], o_ge[NOISE ? NS : 1][CH];
                // operands of stage s, `ahead` rows further down than this iteration's
                auto ld_stage = [&](auto s_tag, long long ahead, T *sv, T *gn, T *gs, T *ge, T *gw, bool with_sv, bool with_links,
                                    bool with_gs) {
                    constexpr int s = decltype(s_tag)::value;
                    const long long ro = (ahead - (long long)(LAG * s)) * pitch;
                    const int p = (rr + 1 + (LAG - 1) * s) & 1;
                    // colour p of a row: nodes q = p, p+2, ... of the lane sit at [p * half + 0 .. CH) of the split row;
                    // their west links are the other colour's: same index for p = 1, one to the left for p = 0.
                    // Unconditional: rows of inactive stages are padding or real rows, both readable.
                    if constexpr (OPS) {
                        // shared-memory ring: row slots oa[s] (this stage's row) and oa[s + 1] (the row above)
                        const unsigned b = oa[s], bn = oa[s + 1];
                        if (SRC && with_sv) lds_operands<CH, true>(sv, b + (p ? OSEG : 0));
                        if (NOISE && with_links) {
                            lds_operands<CH, true>(gn, bn + O_GV + (p ? OSEG : 0));
                            if (with_gs) lds_operands<CH, true>(gs, b + O_GV + (p ? OSEG : 0));
                            lds_operands<CH, true>(ge, b + O_GH + (p ? OSEG : 0));
                            if (p) lds_operands<CH, true>(gw, b + O_GH);
                            else lds_operands<CH, false>(gw, b + O_GH + OSEG - (unsigned)sizeof(T));
                        }
                        return;
                    }
                    if constexpr (!SPLIT) {
                        // natural planes: node q of the lane at [q], its west link at [q - 1]
#pragma unroll
                        for (int q = p; q < C; q += 2) {
                            if (SRC && with_sv) sv[q >> 1] = ld_operand(psrc + ro + q);
                            if (NOISE && with_links) {
                                gn[q >> 1] = ld_operand(pgv + ro - pitch + q);
                                if (with_gs) gs[q >> 1] = ld_operand(pgv + ro + q);
                                gw[q >> 1] = ld_operand(pgh + ro + q - 1);
                                ge[q >> 1] = ld_operand(pgh + ro + q);
                            }
                        }
                        return;
                    }
                    const long long same = ro + (p ? half : 0), other = ro + (p ? 0 : half - 1);
                    if (SRC && with_sv) ld_operands<CH, true>(sv, psrc + same);
                    if (NOISE && with_links) {
                        ld_operands<CH, true>(gn, pgv + same - pitch);
                        if (with_gs) ld_operands<CH, true>(gs, pgv + same);
                        ld_operands<CH, true>(ge, pgh + same);
                        if (p) ld_operands<CH, true>(gw, pgh + other);
                        else ld_operands<CH, false>(gw, pgh + other);
                    }
                };
                auto fetch = [&](auto s_tag) {
                    constexpr int s = decltype(s_tag)::value;
                    // one row between stages: a stage's south link is the previous stage's north link
                    constexpr bool own_gs = !(LAG == 1 && s > 0);
                    if constexpr (FT && s == 0) {
                        // first touches of the iteration come out of the registers loaded one loop trip ago ...
#pragma unroll
                        for (int k = 0; k < CH; ++k) {
                            if (SRC) o_sv[0][k] = ft_sv[rr][k];
                            o_gn[0][k] = ft_gn[rr][k];
                            o_gs[0][k] = ft_gs[rr][k];
                            o_ge[0][k] = ft_ge[rr][k];
                            o_gw[0][k] = ft_gw[rr][k];
                        }
                        // ... and the same loads for the next trip go out now (a whole trip of latency cover)
                        ld_stage(s_tag, UNR, ft_sv[rr], ft_gn[rr], ft_gs[rr], ft_ge[rr], ft_gw[rr], true, true, true);
                    } else if constexpr (FT && s == 1) {
#pragma unroll
                        for (int k = 0; k < CH; ++k)
                            if (SRC) o_sv[1][k] = ft_sv1[rr][k];
                        ld_stage(s_tag, UNR, ft_sv1[rr], nullptr, nullptr, nullptr, nullptr, true, false, false);
                        ld_stage(s_tag, 0, o_sv[s], o_gn[s], o_gs[s], o_ge[s], o_gw[s], false, true, own_gs);
                    } else {
                        ld_stage(s_tag, 0, o_sv[s], o_gn[s], o_gs[s], o_ge[s], o_gw[s], true, true, own_gs);
                    }
                    if constexpr (NOISE && !own_gs) {
#pragma unroll
                        for (int k = 0; k < CH; ++k) o_gs[s][k] = o_gn[s > 0 ? s - 1 : 0][k];
                    }
                };
                static_for<0, LA>(fetch);
                // Part 1 (no dependence between stages): shuffles and everything of the update except the
                // south neighbour, which the previous stage of this iteration is about to produce.
                T tq[NS][CH], gsq[NS][CH];
                static_for<0, NS>([&](auto s_tag) {
                    constexpr int s = decltype(s_tag)::value;
                    if constexpr (s + LA < NS) fetch(std::integral_constant<int, s + LA>{});
                    const int p = (rr + 1 + (LAG - 1) * s) & 1;
                    const int si = RBS_SLOT(-1 - LAG * s);
                    const int su = RBS_SLOT(-2 - LAG * s);
                    T edge;  // the one cross-lane operand of this stage-row
                    if (p == 0) edge = __shfl_up_sync(0xffffffffu, w[si][C - 1], 1);
                    else edge = __shfl_down_sync(0xffffffffu, w[si][0], 1);
#pragma unroll
                    for (int q = p; q < C; q += 2) {
                        const T lf = (q > 0) ? w[si][q - 1] : edge;
                        const T rt = (q < C - 1) ? w[si][q + 1] : edge;
                        T gn = (T)1, gs = (T)1, gw = (T)1, ge = (T)1, sv = (T)0;
                        if (NOISE) {
                            gn = o_gn[s][q >> 1];
                            gs = o_gs[s][q >> 1];
                            gw = o_gw[s][q >> 1];
                            ge = o_ge[s][q >> 1];
                        }
                        if (SRC) sv = o_sv[s][q >> 1];
                        tq[s][q >> 1] = rb_partial<T, NOISE, SRC>(w[su][q], lf, rt, gn, gw, ge, sv);
                        gsq[s][q >> 1] = rb_south_weight<T, NOISE>(gs);
                    }
                });
                // Part 2 (the serial chain): one fused multiply-add + select per stage.
                static_for<0, NS>([&](auto s_tag) {
                    constexpr int s = decltype(s_tag)::value;
                    // stage s relaxes colour s&1 of row r-1-LAG*s; everything is predicated, no branches
                    const int p = (rr + 1 + (LAG - 1) * s) & 1;
                    const int si = RBS_SLOT(-1 - LAG * s);
                    const int sd = RBS_SLOT(-LAG * s);
                    const bool act_s = (act >> (LAG * s)) & 1u;
                    const bool cnt_s = (cnt >> (LAG * s)) & 1u;
                    const unsigned okmask = act_s ? colmask : 0u;
#pragma unroll
                    for (int q = p; q < C; q += 2) {
                        const T old = w[si][q];
                        const T nw = rb_finish<T>(tq[s][q >> 1], gsq[s][q >> 1], w[sd][q]);
                        const T keep = ((okmask >> q) & 1u) ? nw : old;
                        w[si][q] = keep;
                        const T ch = abs_bits(Ar<T>::sub(keep, old));
                        if (cnt_s && ch > mxl[s >> 1]) mxl[s >> 1] = ch;
                    }
                });
                // row r-DEPTH_ROWS has passed every stage
                if (lane_valid && (unsigned)(r - DEPTH_ROWS - rb0) < own_span)
                    store_row<C>(pout, w[RBS_SLOT(-DEPTH_ROWS)]);
                pin += pitch;
                pout += pitch;
                psrc += pitch;
                pgh += pitch;
                pgv += pitch;
            }
            if (SHIFT) {
                // two iterations on: every row of the window is two slots further from the newest one
#pragma unroll
                for (int k = 0; k + 2 < WSL; ++k)
#pragma unroll
                    for (int q = 0; q < C; ++q) w[k][q] = w[k + 2][q];
            }
        }
        if (lane_valid) {
#pragma unroll
            for (int t = 0; t < TD; ++t) mx[t] = max_ignore_nan(mx[t], mxl[t]);
        }
    }
#pragma unroll
    for (int t = 0; t < TD; ++t) {
        const T v = warp_max(mx[t]);
        if (lane == 0 && v > (T)0) atomic_max_nonneg(a.resid + t, v);
    }
}

// What the launcher needs to cut the owned rows into row blocks (the block height is chosen per kernel
// instantiation, because it depends on how many warps of that instantiation are resident).
struct BlockReq {
    long long owned, halo;
    bool top_nb, bot_nb;
    int phase;
    long long forced_rbh;  // > 0: APDE_ROW_BLOCK
    long long band_lo = 0, band_hi = 0, band_rbh = 0;  // band_hi > band_lo: relax exactly these local rows (streamed solve)
};

// Row-block height.  One launch runs n_blocks * n_strips work items on W resident warps in ceil(items / W) rounds of
// (rbh + 2*NS + start-up) row iterations each; taller blocks recompute fewer halo rows (2*NS per block) but the
// last round must not run half empty.  Pick the height with the lowest modelled time (measured on B200, 16384^2
// fp64 depth 3: 256 rows 469 GLUP/s, 1024 rows 500 GLUP/s; fp32, 12 warps per SM: 1024 rows 892, 256 rows 1003).
inline long long pick_row_block(long long owned, long long n_strips, long long warps, int ns, long long halo) {
    long long best = std::max<long long>(256, halo);
    double best_cost = 1e300;
    const long long nb_lo = std::max<long long>(1, (owned + 4095) / 4096), nb_hi = std::max<long long>(nb_lo, std::min<long long>(owned / 64, 512));
    for (long long nb = nb_lo; nb <= nb_hi; ++nb) {
        const long long rbh = std::max<long long>((owned + nb - 1) / nb, std::max<long long>(halo, 8));
        const long long blocks = (owned + rbh - 1) / rbh;
        const long long rounds = (blocks * n_strips + warps - 1) / warps;
        const double cost = (double)rounds * (double)(rbh + 2 * ns + 12);
        if (cost < best_cost * 0.999) {
            best_cost = cost;
            best = rbh;
        }
    }
    return best;
}

template <int C, int TD, bool NOISE, bool SRC>
int launch_pass(apde_plan *p, StreamArgs &a, const BlockReq &rq, cudaStream_t st) {
#ifdef RBS_TMA
    constexpr size_t smem = (size_t)(RBS_THREADS / 32) * (RBS_TMA_SLOTS * 32 * C * sizeof(T) + 8 * RBS_TMA_SLOTS);
#else
    constexpr size_t smem = rbs_ops((int)sizeof(T), C, NOISE) ? (size_t)(RBS_THREADS / 32) * rbs_ops_warp_bytes(C * (int)sizeof(T), TD, SRC, NOISE) : 0;
#endif
    // per device: the dynamic shared-memory opt-in is an attribute of the function IN ONE CONTEXT (a second GPU of the
    // same process -- apde_solve_redblack_multi -- needs its own call, or its launches fail with "invalid argument")
    static int blocks_per_sm_dev[64] = {0};
    int &blocks_per_sm = blocks_per_sm_dev[p->device & 63];
    if (blocks_per_sm == 0) {
        int n = 0;
        if (smem > 48 * 1024)
            APDE_CUDA(cudaFuncSetAttribute(rb_stream_kernel<C, TD, NOISE, SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        APDE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, rb_stream_kernel<C, TD, NOISE, SRC>, RBS_THREADS, smem));
        blocks_per_sm = n > 0 ? n : 1;
    }
    const long long warps_per_block = RBS_THREADS / 32;
    const long long resident_warps = (long long)p->sm_count * blocks_per_sm * warps_per_block;
    // ---- row blocks ---------------------------------------------------------------------------------------------
    // Unphased: the owned rows in blocks of the modelled-best height.  Phased (multi-GPU pass): phase 1 relaxes ONE SHORT
    // block per side that has a neighbour strip -- it holds the `halo` outermost owned rows the exchange ships -- and phase
    // 2 the rows in between.  Round 2 cut edge blocks from the same partition as everything else (1024 rows at 16384^2), so
    // the "edge" kernel ran as long as the whole pass and the halo exchange started when the interior was done
    // (profiles/r2s3_step_timeline.txt: edge 1.62 ms of a 1.72 ms pass); 64-row edge blocks finish in a tenth of a pass
    // and the exchange hides behind the interior.  Results do not depend on the partition (every row block is bitwise
    // the same function of the input buffer).
    const long long own_lo = a.own_lo, own_hi = a.own_hi;
    const long long eh = rq.forced_rbh > 0 ? std::max(rq.forced_rbh, rq.halo) : std::max<long long>(rq.halo, 64);
    const long long et = rq.top_nb ? eh : 0, eb = rq.bot_nb ? eh : 0;
    const bool split = rq.phase != 0 && et + eb > 0 && et + eb < rq.owned;  // else: phase 1 takes the whole strip
    // one launch over the local rows [lo, hi) in blocks of `bh` rows (the kernel itself is the round-2 one: its row-block
    // prologue is left alone -- a two-origin variant of it cost the fp64 Poisson instantiation 5 %)
    auto launch = [&](long long lo, long long hi, long long bh) -> int {
        if (hi <= lo) return APDE_OK;
        a.own_lo = lo;
        a.own_hi = hi;
        a.rbh = std::max<long long>(bh, 1);
        a.blk_a0 = 0;
        a.blk_an = (hi - lo + a.rbh - 1) / a.rbh;
        a.blk_b0 = a.blk_bn = 0;
        a.n_items = a.blk_an * a.n_strips;
        long long grid = (long long)p->sm_count * blocks_per_sm;
        const long long need = (a.n_items + warps_per_block - 1) / warps_per_block;
        if (grid > need) grid = need;
        if (grid < 1) grid = 1;
        rb_stream_kernel<C, TD, NOISE, SRC><<<(unsigned)grid, RBS_THREADS, smem, st>>>(a);
        p->launches++;
        return APDE_OK;
    };
    auto best_height = [&](long long rows) {
        return rq.forced_rbh > 0 ? std::max(rq.forced_rbh, rq.halo) : pick_row_block(rows, a.n_strips, resident_warps, 2 * TD, rq.halo);
    };
    if (rq.band_hi > rq.band_lo) return launch(rq.band_lo, rq.band_hi, std::max<long long>(rq.band_rbh, 8));
    if (rq.phase == 1 && split) {  // one short block per side with a neighbour (two small launches)
        APDE_TRY(launch(own_lo, own_lo + et, et));
        return launch(own_hi - eb, own_hi, eb);
    }
    if (rq.phase == 2 && !split) return APDE_OK;  // phase 1 did everything
    if (rq.phase == 2) return launch(own_lo + et, own_hi - eb, best_height(own_hi - eb - own_lo - et));
    return launch(own_lo, own_hi, best_height(own_hi - own_lo));
}

template <int C, bool NOISE, bool SRC>
int launch_depth(apde_plan *p, StreamArgs &a, const BlockReq &rq, int td, cudaStream_t st) {
    switch (td) {
        case 1: return launch_pass<C, 1, NOISE, SRC>(p, a, rq, st);
        case 2: return launch_pass<C, 2, NOISE, SRC>(p, a, rq, st);
        case 3: return launch_pass<C, 3, NOISE, SRC>(p, a, rq, st);
        case 4: return launch_pass<C, 4, NOISE, SRC>(p, a, rq, st);
    }
    set_error("rb_stream: unsupported fused depth %d", td);
    return APDE_EINVAL;
}

constexpr int RBS_C1 = 16 / (int)sizeof(T);  // one 16-byte vector per lane
constexpr int RBS_C2 = 32 / (int)sizeof(T);  // two

template <int C>
int launch_any(apde_plan *p, StreamArgs &a, const BlockReq &rq, int td, cudaStream_t st) {
    const bool N = p->d.has_noise, S = p->d.has_source;
    if (N && S) return launch_depth<C, true, true>(p, a, rq, td, st);
    if (N) return launch_depth<C, true, false>(p, a, rq, td, st);
    if (S) return launch_depth<C, false, true>(p, a, rq, td, st);
    return launch_depth<C, false, false>(p, a, rq, td, st);
}

// kernel arguments of one pass of td fused sweeps reading u[src_buf], writing u[dst_buf]; the row blocks are filled in
// by launch_pass
static StreamArgs make_stream_args(apde_plan *p, const Geo &g, int src_buf, int dst_buf, int64_t resid_slot, int td) {
    StreamArgs a;
    a.in = p->U<T>(src_buf);
    a.out = p->U<T>(dst_buf);
    const int lane_cols = p->lane_vectors == 2 ? RBS_C2 : RBS_C1;
    const bool split = rbs_split((int)sizeof(T), lane_cols, p->d.has_noise);
    a.src = split ? p->SC<T>() : p->S<T>();
    a.gh = split ? p->GHC<T>() : p->GH<T>();
    a.gv = split ? p->GVC<T>() : p->GV<T>();
    a.resid = p->resid.as<T>() + resid_slot;
    a.stop = p->stop_flag;
    a.pitch = g.pitch;
    a.padl = g.padl;
    a.lrows = g.lrows;
    a.cols = g.cols;
    a.grows = g.grows;
    a.row0 = g.row0;
    a.own_lo = g.own_lo;
    a.own_hi = g.own_hi;
    const int HCA = ((2 * td + lane_cols - 1) / lane_cols) * lane_cols;
    const int VW = 32 * lane_cols - 2 * HCA;
    a.n_strips = (g.cols + VW - 1) / VW;
    return a;
}

}  // namespace

// Runs n_sweeps sweeps as passes of <= temporal_depth fused sweeps; ping-pongs u[cur] -> u[1-cur].
// phase 1 = only the row blocks next to a neighbour strip (their output is what the halo exchange
// sends), phase 2 = the other row blocks, phase 0 = all.  With phases the caller runs a single pass.
int RBS_FN(apde_plan *p, int64_t n_sweeps, int64_t sweep_base, int phase, cudaStream_t st) {
    const Geo g = make_geo(p);
    if (g.cols < 3 || g.lrows < 3 || n_sweeps == 0) return APDE_OK;
    const int max_td = std::min(p->d.temporal_depth, 4);
    if (phase != 0 && n_sweeps > max_td) {
        set_error("rb_stream: phased runs take one pass (n_sweeps <= %d)", max_td);
        return APDE_EINVAL;
    }
    int64_t done = 0;
    while (done < n_sweeps) {
        const int td = (int)std::min<int64_t>(max_td, n_sweeps - done);
        // a phase-2 call completes the pass a phase-1 call started: same buffers
        const int src_buf = p->cur, dst_buf = 1 - p->cur;
        StreamArgs a = make_stream_args(p, g, src_buf, dst_buf, sweep_base + done, td);
        BlockReq rq;
        rq.owned = g.own_hi - g.own_lo;
        rq.halo = p->halo;
        rq.top_nb = p->ghost_top > 0;
        rq.bot_nb = p->ghost_bot > 0;
        rq.phase = phase;
        rq.forced_rbh = p->row_block_forced ? p->row_block : 0;
        {
            const int rc = p->lane_vectors == 2 ? launch_any<RBS_C2>(p, a, rq, td, st) : launch_any<RBS_C1>(p, a, rq, td, st);
            if (rc != APDE_OK) return rc;
        }
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


// One pass of td <= 4 fused sweeps over the local rows [row_lo, row_hi) only, reading u[src_buf] and writing u[1 - src_buf],
// in row blocks of `rbh` rows.  The caller (the streamed one-shot solve, apde_redblack.cu) owns the schedule: p->cur is
// neither read nor changed.  Rows outside the band are untouched in both buffers.
int RBS_BAND_FN(apde_plan *p, int td, int64_t resid_slot, int src_buf, int64_t row_lo, int64_t row_hi, int64_t rbh, cudaStream_t st) {
    const Geo g = make_geo(p);
    if (g.cols < 3 || g.lrows < 3 || td <= 0 || row_hi <= row_lo) return APDE_OK;
    if (td > 4 || row_lo < g.own_lo || row_hi > g.own_hi) {
        set_error("rb_stream band: bad pass (depth %d, rows [%lld, %lld))", td, (long long)row_lo, (long long)row_hi);
        return APDE_EINVAL;
    }
    StreamArgs a = make_stream_args(p, g, src_buf, 1 - src_buf, resid_slot, td);
    BlockReq rq;
    rq.owned = g.own_hi - g.own_lo;
    rq.halo = p->halo;
    rq.top_nb = rq.bot_nb = false;
    rq.phase = 0;
    rq.forced_rbh = 0;
    rq.band_lo = row_lo;
    rq.band_hi = row_hi;
    rq.band_rbh = rbh;
    return p->lane_vectors == 2 ? launch_any<RBS_C2>(p, a, rq, td, st) : launch_any<RBS_C1>(p, a, rq, td, st);
}

#undef RBS_SLOT

}  // namespace apde

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

#include "apde_rb_common.cuh"

namespace apde {
namespace {

using T = RBS_T;

constexpr int RBS_THREADS = 128;
constexpr int RBS_PF = 2;  // rows prefetched ahead (kept even so the unroll period stays even)

struct StreamArgs {
    const T *in;
    T *out;
    const T *src, *gh, *gv;
    T *resid;  // slot of the first sweep of this pass
    long long pitch, padl, lrows, cols, grows, row0;
    long long own_lo, own_hi;
    long long rbh;                 // row-block height
    long long blk_a0, blk_an;      // row blocks [blk_a0, blk_a0+blk_an) ...
    long long blk_b0, blk_bn;      // ... and [blk_b0, blk_b0+blk_bn)
    long long n_strips, n_items;
};

template <int C>
__device__ __forceinline__ void load_row(T (&w)[C], const T *p) {
    constexpr int BYTES = C * (int)sizeof(T);
    static_assert(BYTES % 16 == 0, "lane vector must be a multiple of 16 bytes");
#pragma unroll
    for (int k = 0; k < BYTES / 16; ++k) {
        int4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(reinterpret_cast<const int4 *>(p) + k));
        if (sizeof(T) == 8) {
            w[2 * k + 0] = (T)__hiloint2double(v.y, v.x);
            w[2 * k + 1] = (T)__hiloint2double(v.w, v.z);
        } else {
            w[4 * k + 0] = (T)__int_as_float(v.x);
            w[4 * k + 1] = (T)__int_as_float(v.y);
            w[4 * k + 2] = (T)__int_as_float(v.z);
            w[4 * k + 3] = (T)__int_as_float(v.w);
        }
    }
}

template <int C>
__device__ __forceinline__ void store_row(T *p, const T (&w)[C]) {
    constexpr int BYTES = C * (int)sizeof(T);
#pragma unroll
    for (int k = 0; k < BYTES / 16; ++k) {
        int4 v;
        if (sizeof(T) == 8) {
            v.x = __double2loint((double)w[2 * k + 0]);
            v.y = __double2hiint((double)w[2 * k + 0]);
            v.z = __double2loint((double)w[2 * k + 1]);
            v.w = __double2hiint((double)w[2 * k + 1]);
        } else {
            v.x = __float_as_int((float)w[4 * k + 0]);
            v.y = __float_as_int((float)w[4 * k + 1]);
            v.z = __float_as_int((float)w[4 * k + 2]);
            v.w = __float_as_int((float)w[4 * k + 3]);
        }
        asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(
                         reinterpret_cast<int4 *>(p) + k),
                     "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                     : "memory");
    }
}

template <int C, int TD, bool NOISE, bool SRC>
__global__ void __launch_bounds__(RBS_THREADS) rb_stream_kernel(const StreamArgs a) {
    constexpr int NS = 2 * TD;             // half-sweep stages
    constexpr int NW = NS + 2 + RBS_PF;    // register window rows (even)
    constexpr int CW = 32 * C;             // columns per warp strip
    constexpr int HCA = ((NS + C - 1) / C) * C;  // column halo, lane aligned
    constexpr int VW = CW - 2 * HCA;       // columns a strip stores
    static_assert(VW > 0 && NW % 2 == 0, "bad tile");

    const int lane = threadIdx.x & 31;
    const long long warp_id = (long long)blockIdx.x * (RBS_THREADS / 32) + (threadIdx.x >> 5);
    const long long n_warps = (long long)gridDim.x * (RBS_THREADS / 32);

    T mx[TD];
#pragma unroll
    for (int t = 0; t < TD; ++t) mx[t] = (T)0;

    for (long long item = warp_id; item < a.n_items; item += n_warps) {
        const long long bsel = item / a.n_strips;
        const long long sidx = item - bsel * a.n_strips;
        const long long blk = bsel < a.blk_an ? a.blk_a0 + bsel : a.blk_b0 + (bsel - a.blk_an);
        const long long rb0 = a.own_lo + blk * a.rbh;
        const long long rb1 = min(rb0 + a.rbh, a.own_hi);
        const long long jl0 = sidx * VW - HCA + (long long)C * lane;  // first column of this lane (even)
        const bool lane_valid = (C * lane >= HCA) && (C * lane < CW - HCA) && (jl0 < a.cols);
        unsigned colmask = 0;  // bit q: column jl0+q is an interior column
#pragma unroll
        for (int q = 0; q < C; ++q)
            if (jl0 + q >= 1 && jl0 + q <= a.cols - 2) colmask |= 1u << q;

        const long long ra = max(rb0 - NS, 0ll);             // first streamed row (never updated)
        const long long re = min(rb1 + NS, a.lrows) - 1;     // last streamed row (never updated)
        const long long rs = ra - ((a.row0 + ra) & 1);       // even global row: static colour parity
        const long long r_last = rb1 - 1 + NS;

        const long long lane_off = a.padl + jl0;
        T w[NW][C];
#pragma unroll
        for (int k = 0; k < NW; ++k)
#pragma unroll
            for (int q = 0; q < C; ++q) w[k][q] = (T)0;
#pragma unroll
        for (int k = 0; k < RBS_PF; ++k) {
            const long long r = rs + k;
            if (r >= 0 && r <= re) load_row<C>(w[k], a.in + r * a.pitch + lane_off);
        }

        for (long long base = 0; rs + base <= r_last; base += NW) {
#pragma unroll
            for (int rr = 0; rr < NW; ++rr) {
                const long long r = rs + base + rr;
                if (r <= r_last) {
                    {   // prefetch row r+PF into its future slot
                        const long long rp = r + RBS_PF;
                        if (rp <= re) load_row<C>(w[(rr + RBS_PF) % NW], a.in + rp * a.pitch + lane_off);
                    }
                    const int p = (rr + 1) & 1;  // column parity relaxed in this iteration (all stages)
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        const long long i = r - 1 - s;
                        if (i > ra && i < re) {
                            const int si = (rr - 1 - s + 4 * NW) % NW;
                            const int su = (si + NW - 1) % NW, sd = (si + 1) % NW;
                            const long long gi = a.row0 + i;
                            const bool row_upd = gi > 0 && gi < a.grows - 1;
                            const bool row_cnt = lane_valid && i >= rb0 && i < rb1;
                            const long long e0 = i * a.pitch + lane_off;
                            // the one cross-lane operand of this stage-row
                            T edge;
                            if (p == 0) {
                                edge = __shfl_up_sync(0xffffffffu, w[si][C - 1], 1);
                            } else {
                                edge = __shfl_down_sync(0xffffffffu, w[si][0], 1);
                            }
#pragma unroll
                            for (int q = p; q < C; q += 2) {
                                const T lf = (q > 0) ? w[si][q - 1] : edge;
                                const T rt = (q < C - 1) ? w[si][q + 1] : edge;
                                T gn = (T)1, gs = (T)1, gw = (T)1, ge = (T)1, sv = (T)0;
                                if (NOISE) {
                                    gn = __ldg(a.gv + e0 - a.pitch + q);
                                    gs = __ldg(a.gv + e0 + q);
                                    gw = __ldg(a.gh + e0 + q - 1);
                                    ge = __ldg(a.gh + e0 + q);
                                }
                                if (SRC) sv = __ldg(a.src + e0 + q);
                                const T old = w[si][q];
                                const T nw = rb_relax<T, NOISE, SRC>(w[su][q], w[sd][q], lf, rt, gn, gs, gw, ge, sv);
                                if (row_upd && ((colmask >> q) & 1u)) {
                                    w[si][q] = nw;
                                    if (row_cnt) mx[s >> 1] = max_ignore_nan(mx[s >> 1], Ar<T>::abs(Ar<T>::sub(nw, old)));
                                }
                            }
                        }
                    }
                    const long long o = r - NS;  // this row has passed all stages
                    if (lane_valid && o >= rb0 && o < rb1)
                        store_row<C>(a.out + o * a.pitch + lane_off, w[(rr - NS + 4 * NW) % NW]);
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < TD; ++t) {
        const T v = warp_max(mx[t]);
        if (lane == 0 && v > (T)0) atomic_max_nonneg(a.resid + t, v);
    }
}

template <int C, int TD, bool NOISE, bool SRC>
int launch_pass(apde_plan *p, const StreamArgs &a, cudaStream_t st) {
    static int blocks_per_sm = 0;
    if (blocks_per_sm == 0) {
        int n = 0;
        APDE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, rb_stream_kernel<C, TD, NOISE, SRC>, RBS_THREADS, 0));
        blocks_per_sm = n > 0 ? n : 1;
    }
    const long long warps_per_block = RBS_THREADS / 32;
    long long grid = (long long)p->sm_count * blocks_per_sm;
    const long long need = (a.n_items + warps_per_block - 1) / warps_per_block;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    rb_stream_kernel<C, TD, NOISE, SRC><<<(unsigned)grid, RBS_THREADS, 0, st>>>(a);
    p->launches++;
    return APDE_OK;
}

template <int C, bool NOISE, bool SRC>
int launch_depth(apde_plan *p, const StreamArgs &a, int td, cudaStream_t st) {
    switch (td) {
        case 1: return launch_pass<C, 1, NOISE, SRC>(p, a, st);
        case 2: return launch_pass<C, 2, NOISE, SRC>(p, a, st);
        case 3: return launch_pass<C, 3, NOISE, SRC>(p, a, st);
        case 4: return launch_pass<C, 4, NOISE, SRC>(p, a, st);
    }
    set_error("rb_stream: unsupported fused depth %d", td);
    return APDE_EINVAL;
}

constexpr int RBS_C = 16 / (int)sizeof(T);  // one 16-byte vector per lane

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
        StreamArgs a;
        a.in = p->u[src_buf].as<T>();
        a.out = p->u[dst_buf].as<T>();
        a.src = p->src.as<T>();
        a.gh = p->gh.as<T>();
        a.gv = p->gv.as<T>();
        a.resid = p->resid.as<T>() + sweep_base + done;
        a.pitch = g.pitch;
        a.padl = g.padl;
        a.lrows = g.lrows;
        a.cols = g.cols;
        a.grows = g.grows;
        a.row0 = g.row0;
        a.own_lo = g.own_lo;
        a.own_hi = g.own_hi;
        const long long owned = g.own_hi - g.own_lo;
        a.rbh = p->row_block;
        const long long nblk = (owned + a.rbh - 1) / a.rbh;
        const int HCA = ((2 * td + RBS_C - 1) / RBS_C) * RBS_C;
        const int VW = 32 * RBS_C - 2 * HCA;
        a.n_strips = (g.cols + VW - 1) / VW;
        // edge blocks: first / last row block when that side has a neighbour strip
        const bool top_nb = p->ghost_top > 0, bot_nb = p->ghost_bot > 0;
        long long ea = (top_nb ? 1 : 0), eb = (bot_nb ? 1 : 0);
        if (ea + eb > nblk) { ea = std::min<long long>(ea, nblk); eb = nblk - ea; }
        if (phase == 0) {
            a.blk_a0 = 0; a.blk_an = nblk; a.blk_b0 = 0; a.blk_bn = 0;
        } else if (phase == 1) {
            a.blk_a0 = 0; a.blk_an = ea; a.blk_b0 = nblk - eb; a.blk_bn = eb;
        } else {
            a.blk_a0 = ea; a.blk_an = nblk - ea - eb; a.blk_b0 = 0; a.blk_bn = 0;
        }
        a.n_items = (a.blk_an + a.blk_bn) * a.n_strips;
        if (a.n_items > 0) {
            const bool N = p->d.has_noise, S = p->d.has_source;
            int rc;
            if (N && S) rc = launch_depth<RBS_C, true, true>(p, a, td, st);
            else if (N) rc = launch_depth<RBS_C, true, false>(p, a, td, st);
            else if (S) rc = launch_depth<RBS_C, false, true>(p, a, td, st);
            else rc = launch_depth<RBS_C, false, false>(p, a, td, st);
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

}  // namespace apde

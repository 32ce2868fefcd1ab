// apde_exact.cu -- bit-exact lexicographic Gauss-Seidel on sm_100a.
//
// THIS TRANSLATION UNIT MUST BE COMPILED WITH  -fmad=false .
//
// It replaces the reference's sweep loops
//   AnalogPDESolver.solve  analog_pde_solver/solver.py:263-287
//   DigitalSolver.solve    analog_pde_solver/solver.py:344-368
// and reproduces every iterate bit for bit.  The reference updates nodes in row-major
// order in place, so node (i,j) of sweep t reads NEW north/west (sweep t) and OLD
// south/east (sweep t-1).  All nodes with the same  s = 2t + i + j  are independent:
//   north/west of (t,i,j) have s-1, south/east of sweep t-1 have s-1, own old value s-2.
// Both kernels below walk s ("time-skewed anti-diagonal wavefront"), so many sweeps are
// in flight at once, and every operand is exactly the value the serial loop would read.
//
// Arithmetic per node (solver.py:269-279), one IEEE rounding per operation:
//   n = u[i-1,j]/nv[i-1,j]; s = u[i+1,j]/nv[i,j]; w = u[i,j-1]/nh[i,j-1]; e = u[i,j+1]/nh[i,j]
//   new = ((((n+s)+w)+e)-src)/4.0 ; change = |new-old| ; max_change = max (NaN ignored)
// The fp64 '/' of CUDA is IEEE-correct; /4.0 is exact either way.
//
// K1  exact_small_kernel : whole problem in one CTA's shared memory, all sweeps of a batch
//                          pipelined over s, early stop + rollback inside the kernel.
// K2  exact_ring_kernel  : large grids, anti-diagonal-major layout in HBM/L2, one sweep per
//                          CTA, CTAs chained in a ring by acquire/release progress counters
//                          (CTA of sweep t trails CTA of sweep t-1 by one diagonal).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "apde_common.h"

namespace apde {

// ------------------------------------------------------------------------------------------
// IEEE-exact division by a link value with a precomputed reciprocal.
//
// A link is stored as {b, y}: b = the reference's mismatch factor, y = RN(1/b) computed once at
// upload (__drcp_rn, correctly rounded) or 0.0 when b is not "safe" (see link_reciprocal).
// For |a| in [2^-900, 2^900] (round 2: was 2^+-700; a field decaying away from a hot edge spends ~4 binades per row,
// so the wider window moves 50 more rows of a 4096^2 grid off the general division):
//     q0 = RN(a*y)                 within 1.5 ulp of a/b (y has relative error <= 2^-53)
//     r0 = fma(-q0, b, a)          exact: fits 53 bits, and its last bit has exponent >= e_a - 104 >= -1004 > -1074 while
//                                  |q0| >= 2^-900 * 2^-100 stays normal for the screened links |b| in [2^-100, 2^100]
//     q1 = fma(r0, y, q0)          = RN(a/b + (r0/b)*eps): faithful
//     r1 = fma(-q1, b, a)          exact
//     q2 = fma(r1, y, q1)          = RN(a/b + (r1/b)*eps) = RN(a/b)   (Markstein's theorem: holds for a
//                                  correctly rounded y unless b's significand is all ones -- screened)
// i.e. 1 DMUL + 4 DFMA instead of the ~30-instruction general division.  Zero numerators give the
// correctly signed zero; every other case (tiny/huge/NaN/Inf numerators, unscreened links) takes
// the general IEEE division.  apde_selftest_division() compares both paths bit for bit on the GPU.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double link_reciprocal(double b) {
    const unsigned hi = (unsigned)__double2hiint(b), lo = (unsigned)__double2loint(b);
    const unsigned e = (hi >> 20) & 0x7ffu;                     // biased exponent
    const bool all_ones = ((hi & 0xfffffu) == 0xfffffu) && (lo == 0xffffffffu);
    const bool safe = (e - 923u) <= 200u && !all_ones;          // |b| in [2^-100, 2^100], finite, normal
    return safe ? __drcp_rn(b) : 0.0;
}

__device__ __forceinline__ double exact_div(double a, double2 link) {
    const double b = link.x, y = link.y;
    if (y != 0.0) {
        const unsigned hi = (unsigned)__double2hiint(a);
        const unsigned e = (hi >> 20) & 0x7ffu;
        if ((e - 123u) <= 1800u) {  // |a| in [2^-900, 2^900]
            double q = __dmul_rn(a, y);
            double r = __fma_rn(-q, b, a);
            q = __fma_rn(r, y, q);
            r = __fma_rn(-q, b, a);
            return __fma_rn(r, y, q);
        }
        if (a == 0.0)  // +-0 / finite non-zero b
            return __hiloint2double((int)((hi ^ (unsigned)__double2hiint(b)) & 0x80000000u), 0);
    }
    return __ddiv_rn(a, b);
}

// ------------------------------------------------------------------------------------------
// the node update, shared by K1 and K2
// ------------------------------------------------------------------------------------------
// Branch-free form of exact_div: the quotient of the screened fast path, or the correctly signed zero, selected by
// predicates; anything else (unscreened link, tiny / huge / non-finite numerator) raises `slow` and the caller redoes
// the node with the general division.  No branch per division, so the four divisions of a node interleave in the fp64
// pipe instead of running one behind the other.
__device__ __forceinline__ double band_div(double a, double2 link, bool &slow) {
    const double b = link.x, y = link.y;
    const unsigned hi = (unsigned)__double2hiint(a);
    const bool inr = ((hi & 0x7ff00000u) - (123u << 20)) <= (1800u << 20);  // |a| in [2^-900, 2^900]
    double q = __dmul_rn(a, y);
    double r = __fma_rn(-q, b, a);
    q = __fma_rn(r, y, q);
    r = __fma_rn(-q, b, a);
    q = __fma_rn(r, y, q);
    const double z = __hiloint2double((int)((hi ^ (unsigned)__double2hiint(b)) & 0x80000000u), 0);
    slow = slow || (y == 0.0) || !(inr || a == 0.0);
    return inr ? q : z;
}

// the node with four general IEEE divisions (what exact_div falls back to; same bits by construction), out of line so
// that the rare path costs the hot loop no registers
template <bool HAS_SRC>
__device__ __noinline__ double relax_node_general(double un, double us, double uw, double ue, double bn, double bs, double bw,
                                                  double be, double src) {
    double acc = __dadd_rn(__ddiv_rn(un, bn), __ddiv_rn(us, bs));
    acc = __dadd_rn(acc, __ddiv_rn(uw, bw));
    acc = __dadd_rn(acc, __ddiv_rn(ue, be));
    if (HAS_SRC) acc = __dsub_rn(acc, src);
    return __dmul_rn(acc, 0.25);
}

// R = rows per thread ("walkers"): thread r owns rows r, r + NT, ..., r + (R-1)*NT of the band (band height H = R*NT);
// walker k of thread r is the band's virtual thread v = k*NT + r and runs v steps behind virtual thread 0, so the R
// nodes a thread relaxes in one step are independent of each other (instruction-level parallelism across the four
// divisions x R nodes) and share the step's barrier, flag exchange and loop bookkeeping.
template <bool HAS_NOISE, bool HAS_SRC>
__device__ __forceinline__ double relax_node(double un, double us, double uw, double ue,
                                             double2 ln, double2 ls, double2 lw, double2 le,
                                             double src) {
    if (HAS_NOISE) {
#ifndef APDE_BRANCHY_DIV
        bool slow = false;
        const double qn = band_div(un, ln, slow), qs = band_div(us, ls, slow);  // solver.py:269-270
        const double qw = band_div(uw, lw, slow), qe = band_div(ue, le, slow);  // :271-272
        if (slow) return relax_node_general<HAS_SRC>(un, us, uw, ue, ln.x, ls.x, lw.x, le.x, src);  // rare, out of line
        un = qn;
        us = qs;
        uw = qw;
        ue = qe;
#else
        un = exact_div(un, ln);  // solver.py:269
        us = exact_div(us, ls);  // :270
        uw = exact_div(uw, lw);  // :271
        ue = exact_div(ue, le);  // :272
#endif
    }
    double acc = __dadd_rn(un, us);  // :275, left-associative
    acc = __dadd_rn(acc, uw);
    acc = __dadd_rn(acc, ue);
    if (HAS_SRC) acc = __dsub_rn(acc, src);
    return __dmul_rn(acc, 0.25);  // == /4.0 exactly
}

// {b} -> {b, RN(1/b) or 0}
__global__ void make_links_kernel(const double *__restrict__ b, double2 *__restrict__ out, long long n) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x) {
        const double v = b[k];
        out[k] = make_double2(v, link_reciprocal(v));
    }
}

// self-test: exact_div vs the general division on pseudo-random and adversarial operands
__global__ void division_selftest_kernel(long long n, unsigned long long seed, unsigned long long *mismatches) {
    unsigned long long bad = 0;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x) {
        // splitmix64 stream
        unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(2 * k + 1);
        auto next = [&]() {
            z += 0x9E3779B97F4A7C15ull;
            unsigned long long x = z;
            x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
            x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
            return x ^ (x >> 31);
        };
        const unsigned long long ra = next(), rb = next(), rc = next();
        // denominator: random significand, with a share of near-all-ones / near-power-of-two patterns
        unsigned long long sig = rb & 0xFFFFFFFFFFFFFull;
        const unsigned mode = (unsigned)(rc & 15u);
        if (mode == 0) sig = 0xFFFFFFFFFFFFFull - (rc >> 60);          // all ones and neighbours
        if (mode == 1) sig = (rc >> 60);                                // 1.0 and neighbours
        if (mode == 2) sig = 0xFFFFFFFFFFFFFull - ((rc >> 40) & 0xFFFFF);
        long long eb = 1023 + (long long)((rc >> 8) % 5) - 2;           // 2^-2 .. 2^2
        if (mode == 3) eb = 1023 + (long long)((rc >> 8) % 300) - 150;  // includes unscreened exponents
        if (mode == 4) eb = 1023 + (((rc >> 8) & 1) ? 97 : -100) + (long long)((rc >> 9) % 4);  // the ends of the screen
        const unsigned long long sb = (rc >> 7) & 1ull;
        const double b = __longlong_as_double((long long)((sb << 63) | ((unsigned long long)eb << 52) | sig));
        // numerator: mostly mid-range, some tiny / huge / zero / subnormal / inf / nan
        double a;
        const unsigned ma = (unsigned)((rc >> 20) & 31u);
        if (ma == 0) a = 0.0;
        else if (ma == 1) a = -0.0;
        else if (ma == 2) a = __longlong_as_double((long long)(ra & 0x800FFFFFFFFFFFFFull));  // subnormal
        else if (ma == 3) a = __longlong_as_double((long long)ra);                              // any bits
        else {
            long long ea = 1023 + (long long)((ra >> 53) % 2040) - 1020;  // both sides of the 2^+-900 screen
            if (ma == 5) ea = 1023 - 904 + (long long)((ra >> 53) % 8);  // around the lower end of the numerator window
            if (ma == 6) ea = 1023 + 896 + (long long)((ra >> 53) % 8);  // around the upper end
            if (ma >= 8) ea = 1023 + (long long)((ra >> 53) % 40) - 20;
            a = __longlong_as_double((long long)((ra & 0x800FFFFFFFFFFFFFull) | ((unsigned long long)ea << 52)));
        }
        const double2 link = make_double2(b, link_reciprocal(b));
        const double fast = exact_div(a, link), ref = __ddiv_rn(a, b);
        const bool same = (__double_as_longlong(fast) == __double_as_longlong(ref)) || (fast != fast && ref != ref);
        if (!same) ++bad;
        bool slow = false;  // the branch-free form: whatever it does not hand to the general division must be right
        const double bq = band_div(a, link, slow);
        if (!slow && __double_as_longlong(bq) != __double_as_longlong(ref)) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// ==========================================================================================
// K1: one CTA, problem resident in shared memory
// ==========================================================================================
constexpr int K1_MAX_BATCH = 256;

struct K1Params {
    double *grid;          // rows*cols row-major (in: BC + initial guess, out: solution)
    const double2 *nh;     // rows*(cols-1) {b, 1/b} or null
    const double2 *nv;     // (rows-1)*cols {b, 1/b} or null
    const double *src;     // rows*cols or null
    double *snapshot;      // rows*cols scratch (rollback state)
    double *hist;          // capacity max_iter
    long long *iters_out;  // iterations_run
    int rows, cols;
    int batch;             // sweeps pipelined per drained batch (<= K1_MAX_BATCH)
    long long max_iter;
    double tol;
};

template <bool HAS_NOISE, bool HAS_SRC>
__device__ __forceinline__ void k1_run_batch(double *u, const double2 *nh, const double2 *nv,
                                             const double *src, double *slots, int rows,
                                             int cols, int nb) {
    const int hw = (cols - 1) / 2;  // candidate columns per row and parity
    const int dmax = rows + cols - 4;
    const int s_end = 2 * (nb - 1) + dmax;
    for (int k = threadIdx.y * blockDim.x + threadIdx.x; k < nb; k += blockDim.x * blockDim.y)
        slots[k] = 0.0;
    __syncthreads();
    for (int s = 2; s <= s_end; ++s) {
        const int p = s & 1;
        for (int i = 1 + threadIdx.y; i <= rows - 2; i += blockDim.y) {
            const int off = (i + 1 + p) & 1;
            for (int jj = threadIdx.x; jj < hw; jj += blockDim.x) {
                const int j = 1 + 2 * jj + off;
                if (j > cols - 2) continue;
                const int tt = (s - i - j) >> 1;  // sweep (relative to the batch) this node is in
                if (tt < 0 || tt >= nb || (s - i - j) < 0) continue;
                const int c = i * cols + j;
                double2 ln = {1.0, 1.0}, ls = ln, lw = ln, le = ln;
                double sv = 0.0;
                if (HAS_NOISE) {
                    ln = nv[c - cols];
                    ls = nv[c];
                    lw = nh[i * (cols - 1) + j - 1];
                    le = nh[i * (cols - 1) + j];
                }
                if (HAS_SRC) sv = src[c];
                const double old = u[c];
                const double nw = relax_node<HAS_NOISE, HAS_SRC>(u[c - cols], u[c + cols], u[c - 1],
                                                                 u[c + 1], ln, ls, lw, le, sv);
                const double ch = fabs(__dsub_rn(nw, old));
                u[c] = nw;
                // strict '>' like solver.py:277 (NaN never enters); skip the atomic when it cannot win
                if (ch > slots[tt]) atomic_max_nonneg(&slots[tt], ch);
            }
        }
        __syncthreads();
    }
}

template <bool HAS_NOISE, bool HAS_SRC>
__global__ void __launch_bounds__(1024, 1) exact_small_kernel(K1Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int rows = p.rows, cols = p.cols;
    const int n = rows * cols;
    double *u = reinterpret_cast<double *>(smem_raw);
    double *slots = u + n + (n & 1);
    double *cur = slots + K1_MAX_BATCH;  // (n + (n&1) + K1_MAX_BATCH) is even: cur stays 16-byte aligned
    double2 *nh = nullptr, *nv = nullptr;
    double *src = nullptr;
    if (HAS_NOISE) {
        nh = reinterpret_cast<double2 *>(cur);
        cur += 2 * rows * (cols - 1);
        nv = reinterpret_cast<double2 *>(cur);
        cur += 2 * (rows - 1) * cols;
    }
    if (HAS_SRC) src = cur;
    __shared__ int sh_stop;  // -1: no stop in this batch, else index of first converged sweep

    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthr = blockDim.x * blockDim.y;
    for (int k = tid; k < n; k += nthr) u[k] = p.grid[k];
    if (HAS_NOISE) {
        for (int k = tid; k < rows * (cols - 1); k += nthr) nh[k] = p.nh[k];
        for (int k = tid; k < (rows - 1) * cols; k += nthr) nv[k] = p.nv[k];
    }
    if (HAS_SRC)
        for (int k = tid; k < n; k += nthr) src[k] = p.src[k];
    __syncthreads();

    const bool may_stop = p.tol > 0.0;  // max_change >= 0 is never < tol otherwise (NaN tol: never)
    long long t0 = 0;
    long long iters = p.max_iter;  // for...else of solver.py:286-287
    while (t0 < p.max_iter) {
        const long long left = p.max_iter - t0;
        const int nb = left < p.batch ? (int)left : p.batch;
        if (may_stop) {
            for (int k = tid; k < n; k += nthr) p.snapshot[k] = u[k];
        }
        k1_run_batch<HAS_NOISE, HAS_SRC>(u, nh, nv, src, slots, rows, cols, nb);
        if (tid == 0) {
            int stop = -1;
            if (may_stop)
                for (int k = 0; k < nb; ++k)
                    if (slots[k] < p.tol) {  // solver.py:283
                        stop = k;
                        break;
                    }
            sh_stop = stop;
        }
        __syncthreads();
        const int stop = sh_stop;
        const int keep = stop < 0 ? nb : stop + 1;
        for (int k = tid; k < keep; k += nthr) p.hist[t0 + k] = slots[k];  // solver.py:281
        if (stop >= 0) {
            if (stop != nb - 1) {
                // sweeps after the stopping one already touched the grid: roll back and replay
                // exactly stop+1 sweeps (deterministic, so the replay is bit-identical).
                __syncthreads();
                for (int k = tid; k < n; k += nthr) u[k] = p.snapshot[k];
                __syncthreads();
                k1_run_batch<HAS_NOISE, HAS_SRC>(u, nh, nv, src, slots, rows, cols, stop + 1);
            }
            iters = t0 + stop + 1;  // solver.py:284
            break;
        }
        t0 += nb;
        __syncthreads();
    }
    __syncthreads();
    for (int k = tid; k < n; k += nthr) p.grid[k] = u[k];
    if (tid == 0) *p.iters_out = iters;
}

// ==========================================================================================
// K2: ring of CTAs over an anti-diagonal-major layout
//   A[d][i] (pitch P) holds the value of node (i, j = d - i); for a node on diagonal d:
//     north (i-1,j) = U[d-1][i-1]   west (i,j-1) = U[d-1][i]
//     south (i+1,j) = U[d+1][i+1]   east (i,j+1) = U[d+1][i]
//     link arrays in node layout: LV[d][i] = nv[i,j] (link to south), LH[d][i] = nh[i,j] (link to east)
//     north link = LV[d-1][i-1], west link = LH[d-1][i], south link = LV[d][i], east link = LH[d][i]
//   every access is unit-stride in i.
// ==========================================================================================
// links: row-major {b} -> diagonal-major {b, 1/b}
__global__ void to_diag_links_kernel(const double *__restrict__ rm, double2 *__restrict__ dm, int rows,
                                     int cols, int P) {
    const long long n = (long long)rows * cols;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(k / cols), j = (int)(k - (long long)i * cols);
        const double v = rm[k];
        dm[(long long)(i + j) * P + i] = make_double2(v, link_reciprocal(v));
    }
}
__global__ void to_diag_kernel(const double *__restrict__ rm, double *__restrict__ dm, int rows,
                               int cols, int P) {
    // rm is rows x cols row-major (its own dims); dm[(i+j)*P + i]
    const long long n = (long long)rows * cols;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(k / cols), j = (int)(k - (long long)i * cols);
        dm[(long long)(i + j) * P + i] = rm[k];
    }
}
__global__ void from_diag_kernel(const double *__restrict__ dm, double *__restrict__ rm, int rows,
                                 int cols, int P) {
    const long long n = (long long)rows * cols;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(k / cols), j = (int)(k - (long long)i * cols);
        rm[k] = dm[(long long)(i + j) * P + i];
    }
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ void prefetch_l2_sector(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct K2Params {
    double *U;
    const double2 *LH, *LV;
    const double *S;
    double *hist;                   // this launch's first sweep
    unsigned long long *progress;   // one counter per CTA (stride 16 to keep them on separate lines)
    int rows, cols, P;
    long long n_sweeps;             // sweeps in this launch
    int flag_tid;                   // K2 / K2d: the thread that polls and publishes the progress counters
    int l2_ahead;                   // K2 / K2d: diagonals the round's first CTA prefetches ahead into L2 (0: off)
    int dbg;                        // TIMING EXPERIMENTS ONLY (wrong results): 1 no result store, 2 no release, 4 no block barrier
};


// The CTA of a round's first sweep (c == 0) is the first to touch a diagonal's links / source / values since the previous
// round went by (4096^2: 0.67 GB per pass, the L2 holds 126 MB), so each of its node iterations waits for HBM instead of
// L2 -- and every other sweep in flight follows it flag by flag at its pace (round 2, session 3: a sweep running ALONE
// takes the same 26 ms as 148 in flight; the followers' flag threads spin ~37 polls per diagonal).  It therefore pulls the
// operand rows of diagonal d + ahead into L2 at the start of step d: one prefetch per 32-byte sector, a handful per
// thread and step, in a loop of its own so that the node loop's registers are untouched.
template <bool HAS_NOISE, bool HAS_SRC>
__device__ __forceinline__ void k2_leader_prefetch(const K2Params &p, int d, int tid, int nthr) {
    const int da = d + p.l2_ahead;
    const int dmax = p.rows + p.cols - 4;
    if (da > dmax) return;
    const int ilo = max(1, da - (p.cols - 2)), ihi = min(p.rows - 2, da - 1);
    const long long oa = (long long)da * p.P;
    if (HAS_NOISE)
        for (int i = (ilo & ~1) + 2 * tid; i <= ihi; i += 2 * nthr) {  // 16-byte links: two per sector
            prefetch_l2_sector(p.LV + oa + i);
            prefetch_l2_sector(p.LH + oa + i);
        }
    for (int i = (ilo & ~3) + 4 * tid; i <= ihi + 1; i += 4 * nthr) {  // 8-byte values: four per sector
        prefetch_l2_sector(p.U + oa + p.P + i);                         // south / east operands live on diagonal da + 1
        if (HAS_SRC) prefetch_l2_sector(p.S + oa + i);
    }
}

// MAXT is THE block size (the launch must use exactly MAXT threads: the node stride of a thread is a compile-time
// constant so that a thread's further nodes are reached with immediate offsets); MINB shapes the register budget:
// <1024,1> and <512,2> get 64 registers, <256,2> 128.
// NB = nodes whose operands a thread requests before it relaxes any of them (nodes of one diagonal are independent):
// their L2 round trips overlap.  A node has 5 (+1 source, +8 for the four {b, 1/b} links) operand doubles, so NB is
// what the register budget of the launch shape allows: 4 without links, 1 with links at 64 registers, 4 at 128.
template <bool HAS_NOISE, bool HAS_SRC, int MAXT, int MINB, int NB>
__global__ void __launch_bounds__(MAXT, MINB) exact_ring_kernel(K2Params p) {
    const int rows = p.rows, cols = p.cols;
    const long long P = p.P;
    constexpr int nthr = MAXT;
    const int G = gridDim.x, c = blockIdx.x, tid = threadIdx.x;
    const int dmax = rows + cols - 4;
    const long long ND = dmax - 1;  // diagonals 2..dmax
    const int pred = (c + G - 1) % G;
    __shared__ double red[32];
    double *__restrict__ U = p.U;

    long long m = 0;
    for (long long t = c; t < p.n_sweeps; t += G, ++m) {
        const long long mpred = (c == 0) ? m - 1 : m;  // index of sweep t-1 in pred's own sequence
        double mx = 0.0;
        // sweep t-1 must have finished diagonal d+1 before this sweep touches diagonal d (RAW on
        // south/east, WAR on the diagonal itself).  Thread 0 polls for the NEXT diagonal while the rest
        // of the block is still finishing the current one, so one barrier per diagonal suffices.
        auto wait_pred = [&](int d) {
            const unsigned long long need = (unsigned long long)(mpred * ND + (long long)(min(d + 1, dmax) - 1));
            const unsigned long long *flag = p.progress + (size_t)pred * 16;
            while (ld_acquire_u64(flag) < need) {
            }
        };
        const bool flagger = tid == p.flag_tid;
        if (t > 0 && flagger) wait_pred(2);
        __syncthreads();
        // Element offset of this thread's first node of diagonal d, (d*P + ilo + tid), advanced incrementally: the
        // seven operand rows of a node (u on diagonals d-1, d, d+1; links and source on d-1 and d) then sit at fixed
        // offsets from it and the thread's further nodes nthr elements on -- no per-node 64-bit address arithmetic
        // (which, at 64 registers, used to be recomputed and spilled for every node).
        long long off = 2 * P + max(1, 2 - (cols - 2)) + tid;
        for (int d = 2; d <= dmax; ++d) {
            if (c == 0 && p.l2_ahead > 0) k2_leader_prefetch<HAS_NOISE, HAS_SRC>(p, d, tid, nthr);
            const int ilo = max(1, d - (cols - 2)), ihi = min(rows - 2, d - 1);
            const double *me0 = U + off;            // (d, i)
            const double *up0 = me0 - P;            // (d-1, i): west; [-1]: north
            const double *dn0 = me0 + P;            // (d+1, i): east; [+1]: south
            const double2 *lvd0 = p.LV + off, *lhd0 = p.LH + off;          // links of diagonal d   : south, east
            const double2 *lvu0 = lvd0 - P, *lhu0 = lhd0 - P;              // links of diagonal d-1 : [-1] north, west
            const double *s0 = p.S + off;
            for (int ib = ilo + tid; ib <= ihi; ib += NB * nthr) {
                const long long cb = ib - (ilo + tid);  // element distance of this chunk from the thread's first node
                const double *up = up0 + cb, *dn = dn0 + cb, *sp = s0 + cb;
                double *me = const_cast<double *>(me0) + cb;
                const double2 *lvu = lvu0 + cb, *lhu = lhu0 + cb, *lvd = lvd0 + cb, *lhd = lhd0 + cb;
                // request the operands of all NB nodes first; the stores follow all loads (for the compiler they alias)
                double un[NB], uw[NB], us[NB], ue[NB], old[NB], sv[NB];
                double2 ln[NB], lw[NB], ls[NB], le[NB];
#pragma unroll
                for (int e = 0; e < NB; ++e) {
                    const int k = e * nthr;  // compile-time: immediate address offsets
                    ln[e] = lw[e] = ls[e] = le[e] = make_double2(1.0, 1.0);
                    sv[e] = 0.0;
                    if (ib + k <= ihi) {
                        un[e] = __ldcg(up + k - 1);
                        uw[e] = __ldcg(up + k);
                        us[e] = __ldcg(dn + k + 1);
                        ue[e] = __ldcg(dn + k);
                        old[e] = __ldcg(me + k);
                        if (HAS_NOISE) {
                            ln[e] = __ldg(lvu + k - 1);
                            lw[e] = __ldg(lhu + k);
                            ls[e] = __ldg(lvd + k);
                            le[e] = __ldg(lhd + k);
                        }
                        if (HAS_SRC) sv[e] = __ldg(sp + k);
                    }
                }
#pragma unroll
                for (int e = 0; e < NB; ++e) {
                    const int k = e * nthr;
                    if (ib + k <= ihi) {
                        const double nw = relax_node<HAS_NOISE, HAS_SRC>(un[e], us[e], uw[e], ue[e], ln[e], ls[e], lw[e], le[e], sv[e]);
                        mx = max_ignore_nan(mx, fabs(__dsub_rn(nw, old[e])));
                        if (!(p.dbg & 1)) __stcg(me + k, nw);
                    }
                }
            }
            // next diagonal: one pitch further, and one node further once the diagonal starts at the east column
            off += P + (max(1, d + 1 - (cols - 2)) - ilo);
            if (t > 0 && flagger && d < dmax) wait_pred(d + 1);
            if (!(p.dbg & 4)) __syncthreads();
            if (flagger && !(p.dbg & 2))
                st_release_u64(p.progress + (size_t)c * 16, (unsigned long long)(m * ND + (d - 1)));
        }
        // block max of this sweep -> residual_history[t]   (solver.py:281)
        mx = warp_max(mx);
        if ((tid & 31) == 0) red[tid >> 5] = mx;
        __syncthreads();
        if (tid < 32) {
            double v = (tid < (nthr + 31) / 32) ? red[tid] : 0.0;
            v = warp_max(v);
            if (tid == 0) p.hist[t] = v;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// K2b: the same ring, restructured for latency.
//   * the sweep's own previous diagonal (north/west operands) lives in shared memory, double
//     buffered -- it never round-trips through L2;
//   * south/east/old operands of diagonal d+1 (written by the predecessor CTA, read from L2) are
//     requested right after diagonal d's node consumed its own, so the L2 latency overlaps the
//     remaining nodes, the barrier and the next diagonal's shared-memory reads;
//   * link/source rows (immutable) are prefetched into L1 two diagonals ahead;
//   * progress flags are exchanged once per KB diagonals instead of every diagonal (the CTA of
//     sweep t trails the CTA of sweep t-1 by KB+1 diagonals), so the acquire-poll and the
//     release fence leave the per-diagonal critical path.
// NPT = nodes per thread per diagonal (compile time), blockDim.x * NPT >= longest diagonal.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_l1_line(const void *p) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

template <bool HAS_NOISE, bool HAS_SRC, int NPT>
__global__ void __launch_bounds__(1024, 1) exact_ring2_kernel(K2Params p, int kb) {
    extern __shared__ __align__(16) double sdiag[];  // 2 x SP doubles: this sweep's diagonal d-1 / d
    const int rows = p.rows, cols = p.cols;
    const long long P = p.P;
    const int SP = rows + 2;
    const int G = gridDim.x, c = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    const int dmax = rows + cols - 4;
    const long long ND = dmax - 1;
    const int pred = (c + G - 1) % G;
    __shared__ double red[32];
    double *__restrict__ U = p.U;

    long long m = 0;
    for (long long t = c; t < p.n_sweeps; t += G, ++m) {
        const long long mpred = (c == 0) ? m - 1 : m;
        double mx = 0.0;
        int cur = 0;  // sdiag[cur] holds diagonal d-1 of this sweep
        // diagonal 1 = nodes (0,1) and (1,0): both on the fixed ring
        if (tid == 0) sdiag[0] = __ldcg(U + P + 0);
        if (tid == 1) sdiag[1] = __ldcg(U + P + 1);
        double us[NPT], ue[NPT], old[NPT];
        for (int d0 = 2; d0 <= dmax; d0 += kb) {
            const int d1 = min(d0 + kb - 1, dmax);
            if (t > 0 && tid == 0) {
                // the predecessor sweep must be past diagonal d1+1 for the whole batch
                const unsigned long long need = (unsigned long long)(mpred * ND + (long long)(min(d1 + 1, dmax) - 1));
                const unsigned long long *flag = p.progress + (size_t)pred * 16;
                while (ld_acquire_u64(flag) < need) {
                }
            }
            __syncthreads();
            {   // operands of the batch's first diagonal (latency exposed once per batch)
                const int ilo = max(1, d0 - (cols - 2)), ihi = min(rows - 2, d0 - 1);
                const double *dn = U + (long long)(d0 + 1) * P, *me = U + (long long)d0 * P;
#pragma unroll
                for (int e = 0; e < NPT; ++e) {
                    const int i = ilo + tid + e * nthr;
                    if (i <= ihi) {
                        us[e] = __ldcg(dn + i + 1);
                        ue[e] = __ldcg(dn + i);
                        old[e] = __ldcg(me + i);
                    }
                }
            }
            for (int d = d0; d <= d1; ++d) {
                const int ilo = max(1, d - (cols - 2)), ihi = min(rows - 2, d - 1);
                const int lo_all = max(0, d - (cols - 1)), hi_all = min(rows - 1, d);
                const double *sp = sdiag + cur * SP;
                double *sn = sdiag + (cur ^ 1) * SP;
                double *me = U + (long long)d * P;
                const bool more = d < d1;
                const int nlo = max(1, d + 1 - (cols - 2)), nhi = min(rows - 2, d);
                const double *ndn = U + (long long)(d + 2) * P, *nme = U + (long long)(d + 1) * P;
#pragma unroll
                for (int e = 0; e < NPT; ++e) {
                    const int i = ilo + tid + e * nthr;
                    if (i <= ihi) {
                        double2 ln = {1.0, 1.0}, ls = ln, lw = ln, le = ln;
                        double sv = 0.0;
                        if (HAS_NOISE) {
                            ln = __ldg(p.LV + (long long)(d - 1) * P + i - 1);
                            lw = __ldg(p.LH + (long long)(d - 1) * P + i);
                            ls = __ldg(p.LV + (long long)d * P + i);
                            le = __ldg(p.LH + (long long)d * P + i);
                            if (d + 2 <= dmax) {
                                prefetch_l1_line(p.LV + (long long)(d + 2) * P + i);
                                prefetch_l1_line(p.LH + (long long)(d + 2) * P + i);
                            }
                        }
                        if (HAS_SRC) {
                            sv = __ldg(p.S + (long long)d * P + i);
                            if (d + 2 <= dmax) prefetch_l1_line(p.S + (long long)(d + 2) * P + i);
                        }
                        const double nw = relax_node<HAS_NOISE, HAS_SRC>(sp[i - 1], us[e], sp[i], ue[e], ln, ls, lw, le, sv);
                        mx = max_ignore_nan(mx, fabs(__dsub_rn(nw, old[e])));
                        sn[i] = nw;
                        __stcg(me + i, nw);
                    }
                    // request the next diagonal's operands for this thread's slot e
                    const int ni = nlo + tid + e * nthr;
                    if (more && ni <= nhi) {
                        us[e] = __ldcg(ndn + ni + 1);
                        ue[e] = __ldcg(ndn + ni);
                        old[e] = __ldcg(nme + ni);
                    }
                }
                // ring nodes at both ends of diagonal d are operands of diagonal d+1
                if (tid == nthr - 1 && lo_all < ilo) sn[lo_all] = __ldcg(me + lo_all);
                if (tid == nthr - 2 && hi_all > ihi) sn[hi_all] = __ldcg(me + hi_all);
                __syncthreads();
                cur ^= 1;
            }
            if (tid == 0) st_release_u64(p.progress + (size_t)c * 16, (unsigned long long)(m * ND + (d1 - 1)));
        }
        mx = warp_max(mx);
        if ((tid & 31) == 0) red[tid >> 5] = mx;
        __syncthreads();
        if (tid < 32) {
            double v = (tid < (nthr + 31) / 32) ? red[tid] : 0.0;
            v = warp_max(v);
            if (tid == 0) p.hist[t] = v;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// K2d: the K2 ring with the sweep's own operands resident in shared memory (mismatched meshes).
//
// K2 moves 112 B per update through L2 (5 u loads, four 16-byte {b, 1/b} links, 1 store) and at 4096^2 sits at
// ~10 TB/s of L2 traffic with and without links, close to what the L2 delivers -- the hypothesis this kernel was built
// to test.  Of those operands only south / east
// (values of sweep t-1) and the node's own south / east links are new to the CTA; everything else it has touched one
// diagonal earlier.  Two kinds of in-place shared-memory arrays carry them from diagonal to diagonal:
//   row-indexed    Ui[i]  = u_t(i, j-1)        west value     written by node (i, j-1) as its result
//                  Oi[i]  = u_{t-1}(i, j)      old value      written by node (i, j-1): its east operand
//                  LHi[i] = nh[i, j-1]         west link      written by node (i, j-1): its east link
//   column-indexed Uj[j]  = u_t(i-1, j)        north value    written by node (i-1, j) as its result
//                  LVj[j] = nv[i-1, j]         north link     written by node (i-1, j): its south link
// A diagonal holds at most one node per row and per column, so every cell is read and then rewritten by exactly one
// thread per diagonal and the barrier K2 already has between diagonals orders the hand-over: no double buffering.
// Per update: 16 B of u + 32 B of links loaded, 8 B stored (56 B through L2 instead of 112).  32 B per row + 24 B per
// column of shared memory: 4096^2 needs 224 KB, so one 1024-thread CTA per SM.  The row's / column's first cells come
// from the fixed ring (boundary values, links of row 0 / column 0), loaded at the start of every sweep; the old value of
// a row's first interior node is the one operand nobody loaded before (one extra 8-byte load per diagonal).
// NB = nodes whose L2 operands a thread requests before it relaxes any of them (as in K2; only 1 is instantiated:
// 2 spills at the 64 registers a 1024-thread CTA leaves).
// MEASURED (B200, 4096^2, 1 % mismatch, profiles/r2s3_exact_experiments.txt): bit-exact, 90 GLUP/s at K=1000 against
// K2's 94-97, 65 against 64 at K=100 -- half the L2 bytes buy nothing, so L2 traffic is not what bounds the ring.  Opt-in
// (APDE_RING_KERNEL=k2d), kept as the tested cross-check of that statement.
// ------------------------------------------------------------------------------------------
#ifndef K2D_DEFAULT
#define K2D_DEFAULT false
#endif
static inline size_t k2d_smem_bytes(int64_t rows, int64_t cols) { return (size_t)(32 * rows + 24 * cols); }

template <bool HAS_SRC, int NB>
__global__ void __launch_bounds__(1024, 1) exact_ring_smem_kernel(K2Params p) {
    extern __shared__ __align__(16) unsigned char k2d_smem[];
    const int rows = p.rows, cols = p.cols;
    const long long P = p.P;
    constexpr int nthr = 1024;
    double2 *LHi = reinterpret_cast<double2 *>(k2d_smem);  // rows
    double2 *LVj = LHi + rows;                              // cols
    double *Ui = reinterpret_cast<double *>(LVj + cols);    // rows
    double *Oi = Ui + rows;                                 // rows
    double *Uj = Oi + rows;                                 // cols
    const int G = gridDim.x, c = blockIdx.x, tid = threadIdx.x;
    const int dmax = rows + cols - 4;
    const long long ND = dmax - 1;  // diagonals 2..dmax
    const int pred = (c + G - 1) % G;
    const bool flagger = tid == p.flag_tid;
    __shared__ double red[32];
    double *__restrict__ U = p.U;

    long long m = 0;
    for (long long t = c; t < p.n_sweeps; t += G, ++m) {
        const long long mpred = (c == 0) ? m - 1 : m;
        double mx = 0.0;
        auto wait_pred = [&](int d) {  // as K2: sweep t-1 past diagonal d+1 before this sweep touches diagonal d
            const unsigned long long need = (unsigned long long)(mpred * ND + (long long)(min(d + 1, dmax) - 1));
            const unsigned long long *flag = p.progress + (size_t)pred * 16;
            while (ld_acquire_u64(flag) < need) {
            }
        };
        // the fixed ring: west value / west link of every row's first node, north value / north link of every column's
        for (int i = tid; i < rows; i += nthr) {
            Ui[i] = __ldcg(U + (long long)i * P + i);    // u(i, 0)
            LHi[i] = __ldg(p.LH + (long long)i * P + i);  // nh[i, 0]
        }
        for (int j = tid; j < cols; j += nthr) {
            Uj[j] = __ldcg(U + (long long)j * P);         // u(0, j)
            LVj[j] = __ldg(p.LV + (long long)j * P);      // nv[0, j]
        }
        if (t > 0 && flagger) wait_pred(2);
        __syncthreads();
        long long off = 2 * P + max(1, 2 - (cols - 2)) + tid;  // (d*P + ilo + tid), advanced incrementally as in K2
        for (int d = 2; d <= dmax; ++d) {
            if (c == 0 && p.l2_ahead > 0) k2_leader_prefetch<true, HAS_SRC>(p, d, tid, nthr);
            const int ilo = max(1, d - (cols - 2)), ihi = min(rows - 2, d - 1);
            double *me0 = U + off;                               // (d, i); +P: east, +P+1: south
            const double2 *lv0 = p.LV + off, *lh0 = p.LH + off;  // south / east link of the node
            const double *s0 = p.S + off;
            for (int ib = ilo + tid; ib <= ihi; ib += NB * nthr) {
                const long long cb = ib - (ilo + tid);
                double *me = me0 + cb;
                const double2 *lv = lv0 + cb, *lh = lh0 + cb;
                const double *sp = s0 + cb;
                // the operands that come through L2, for all NB nodes before any is relaxed (their round trips overlap)
                double us[NB], ue[NB], sv[NB], og[NB];
                double2 ls[NB], le[NB];
#pragma unroll
                for (int e = 0; e < NB; ++e) {
                    const int k = e * nthr;
                    sv[e] = og[e] = 0.0;
                    if (ib + k <= ihi) {
                        us[e] = __ldcg(me + k + P + 1);
                        ue[e] = __ldcg(me + k + P);
                        ls[e] = __ldg(lv + k);
                        le[e] = __ldg(lh + k);
                        if (HAS_SRC) sv[e] = __ldg(sp + k);
                        if (d - (ib + k) == 1) og[e] = __ldcg(me + k);  // the row's first interior node
                    }
                }
#pragma unroll
                for (int e = 0; e < NB; ++e) {
                    const int k = e * nthr;
                    const int i = ib + k, j = d - i;
                    if (i <= ihi) {
                        const double old = (j == 1) ? og[e] : Oi[i];
                        const double nw = relax_node<true, HAS_SRC>(Uj[j], us[e], Ui[i], ue[e], LVj[j], ls[e], LHi[i], le[e], sv[e]);
                        mx = max_ignore_nan(mx, fabs(__dsub_rn(nw, old)));
                        if (!(p.dbg & 1)) __stcg(me + k, nw);
                        Ui[i] = nw;
                        Uj[j] = nw;
                        Oi[i] = ue[e];
                        LHi[i] = le[e];
                        LVj[j] = ls[e];
                    }
                }
            }
            off += P + (max(1, d + 1 - (cols - 2)) - ilo);
            if (t > 0 && flagger && d < dmax) wait_pred(d + 1);
            if (!(p.dbg & 4)) __syncthreads();
            if (flagger && !(p.dbg & 2)) st_release_u64(p.progress + (size_t)c * 16, (unsigned long long)(m * ND + (d - 1)));
        }
        mx = warp_max(mx);
        if ((tid & 31) == 0) red[tid >> 5] = mx;
        __syncthreads();
        if (tid < 32) {
            double v = red[tid];
            v = warp_max(v);
            if (tid == 0) p.hist[t] = v;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// K2c: row bands, every thread walks along ONE grid row (round 2; measured slower than K2, opt-in: APDE_RING_KERNEL=band).
//
// The interior rows are cut into bands of NT rows; a CTA owns (sweep slot, band) and thread r of it
// owns row i = 1 + band*NT + r for the whole launch.  At CTA step S thread r handles column
// j = (S - r) mod cols - 1 of task (S - r) div cols (task n of slot s = sweep s + n*Gs), i.e. all
// threads of a CTA sit on one anti-diagonal (unit-stride accesses in the diagonal-major layout) and a
// thread that finishes its row starts the same row of the CTA's next sweep in the next step -- no
// pipeline drain between the sweeps of a CTA.  Columns -1 and 0 are two bookkeeping steps per row
// that pull the west boundary value and the west link through the same register path as every
// other column (so no step ever waits for a special-case load).
// What a node needs and where it comes from:
//   west  u_t(i,j-1)      the thread's own previous result              register
//   west link nh[i,j-1]   the thread's own previous east link            register
//   old   u_{t-1}(i,j)    the thread's own previous east operand         register
//   north u_t(i-1,j)      thread r-1's previous result                   shuffle (lane 0: shared memory;
//                                                                        thread 0: the band above, L2)
//   south u_{t-1}(i+1,j)  thread r+1's east operand of this step         shuffle (lane 31 / last row: L2)
//   east  u_{t-1}(i,j+1)  written by the CTA of sweep t-1                ONE 8-byte L2 load, requested a step ahead
//   links nv[i-1,j], nv[i,j], nh[i,j], source                            __ldg (north link: an L1 hit, it was
//                                                                        the neighbour's south link a step ago)
// => 8 B of u + 32 B of links loaded and 8 B stored per update through L2, against 40 + 64 + 8 B for K2
// (which re-reads all five u operands and all four links per node).
// Ordering between CTAs: every CTA publishes the number of steps it has completed (release, every kb
// steps); before a batch of kb steps three lanes acquire-poll
//   A  (slot-1, band)    >= S+3      east/south operands (+1: they are requested one step ahead)
//   B  (slot-1, band+1)  >= S-NT+3   south operand of the band's last row
//   C  (slot,   band-1)  >= S+NT+1   north operand of the band's first row (also the WAR on that row)
// with S the last step of the batch; the predecessor of slot 0 is slot Gs-1 one task earlier (-cols).
// A thread runs its next task while later rows still finish the current one, so the ring of sweep
// slots must be short enough for the slot-0 CTA never to wait for itself: the host keeps
// Gs * (3*kb + 4) <= cols (tools/band_protocol_sim.py replays the protocol on the CPU).  One __syncthreads per step
// (north hand-off between warps), one more where progress is published.
// ------------------------------------------------------------------------------------------
// LPF: the links and the source value of the NEXT step are requested into registers together with its east operand
// (12-14 more registers per walker), so no load of the current step is ever waited for; without it they are read in
// the step itself behind an L1 prefetch.  Either way every thread pulls its operand lines of the step L2AHEAD steps
// ahead from HBM into L2: the first sweep to reach a diagonal (its links were evicted since the previous round) would
// otherwise pay a DRAM round trip in every step, and every other sweep in flight is rate-locked to it.
#ifndef APDE_BAND_L2AHEAD
#define APDE_BAND_L2AHEAD 12
#endif
__device__ __forceinline__ void prefetch_l2_line(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Operands of one step that come from global memory, requested one step ahead.  The kernel keeps TWO of these and
// alternates them between consecutive steps (the loop body is instantiated twice), so a request never has to be
// moved to another register: it stays in flight across the step's barrier and is first waited for where the next step
// uses it.  (A single set rotated with moves at the end of the step made every step wait for its own requests -- a
// full L2 / DRAM round trip per step, 2.5 us per step even on an otherwise idle SM.)
template <int R, bool LPF>
struct BandOperands {
    double ue[R];               // east operand u_{t-1}(i, j+1)
    double us[R];               // south operand from memory: lane 31 and the band's last row
    double un0;                 // north operand of the band's first row
    double2 le[LPF ? R : 1], ls[LPF ? R : 1], ln[LPF ? R : 1];  // east / south / north link of the node
    double sv[LPF ? R : 1];     // source value
};

template <bool HAS_NOISE, bool HAS_SRC, int NT, int R, int MINB, bool LPF>
__global__ void __launch_bounds__(NT, MINB) exact_band_kernel(K2Params p, int B, int Gs, int kb) {
    constexpr int H = R * NT, NWARP = NT / 32;
    const int rows = p.rows, cols = p.cols;
    const unsigned P = (unsigned)p.P;  // element offsets fit 32 bits (checked by the host): one IMAD.WIDE per address
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = blockIdx.x / B, band = blockIdx.x - slot * B;
    const int CP = cols;  // steps per task: columns -1 .. cols-2
    const int ntasks = (int)((p.n_sweeps - slot + Gs - 1) / Gs);  // the launch guarantees slot < n_sweeps
    const int pslot = slot == 0 ? Gs - 1 : slot - 1;
    const int ntasks_p = (int)((p.n_sweeps - pslot + Gs - 1) / Gs);
    const int i0 = 1 + band * H;
    const int rows_b = min(H, rows - 1 - i0);  // rows of this band (>= 1)
    const long long Stot = (long long)ntasks * CP + H - 1, totP = (long long)ntasks_p * CP + H - 1;
    // north hand-off between neighbouring virtual threads that are not neighbouring lanes: [parity][walker][warp]
    __shared__ double nbuf[2][R][NWARP];  // lane 31's result of the step
    double *__restrict__ U = p.U;
    const double2 *__restrict__ LH = p.LH, *__restrict__ LV = p.LV;
    const double *__restrict__ SRC = p.S;

    // what this thread polls before a batch of kb steps (threads 0..2), as  flag >= min(S + kb + f_add, f_cap)
    const unsigned long long *f_ptr = nullptr;
    long long f_add = 0, f_cap = 0;
    {
        const long long adj = slot == 0 ? CP : 0;
        if (tid == 0 && Gs > 1) {
            f_ptr = p.progress + ((size_t)pslot * B + band) * 16;
            f_add = 2 - adj;
            f_cap = totP;
        }
        if (tid == 1 && band + 1 < B) {
            f_ptr = p.progress + ((size_t)pslot * B + band + 1) * 16;
            f_add = 2 - H - adj;
            f_cap = totP;
        }
        if (tid == 2 && band > 0) {
            f_ptr = p.progress + ((size_t)slot * B + band - 1) * 16;
            f_add = H;
            f_cap = Stot;
        }
    }
    unsigned long long *fme = p.progress + (size_t)blockIdx.x * 16;

    // the west link: carried in registers from the previous step (R <= 2) or re-read (an L1 hit: it was the east link
    // one step ago) where four walkers leave no registers for it
    constexpr bool LWREG = R <= 2;
    const unsigned off00 = (unsigned)(i0 + tid - 1) * P + (unsigned)(i0 + tid);  // node (i, -1) of walker 0
    unsigned off[R];
    int jx[R], n[R];
    double uw[R], old[R], mx[R];
    double2 lw[LWREG ? R : 1];
#define OFF0(k) (off00 + (unsigned)((k) * NT) * (P + 1u))
#pragma unroll
    for (int k = 0; k < R; ++k) {
        off[k] = OFF0(k);
        jx[k] = -(k * NT + tid);  // column jx - 1; negative: the row has not started yet
        n[k] = 0;
        uw[k] = old[k] = mx[k] = 0.0;
        if (LWREG) lw[k] = make_double2(1.0, 1.0);
    }
    const int v_last = rows_b - 1;  // the band's last row: its south operand comes from memory
    BandOperands<R, LPF> opA, opB;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        opA.ue[k] = opA.us[k] = opB.ue[k] = opB.us[k] = 0.0;
        if (LPF) {
            opA.le[k] = opA.ls[k] = opA.ln[k] = opB.le[k] = opB.ls[k] = opB.ln[k] = make_double2(1.0, 1.0);
            opA.sv[k] = opB.sv[k] = 0.0;
        }
    }
    opA.un0 = opB.un0 = 0.0;
    if (tid == 0 && ntasks > 0) opA.ue[0] = __ldcg(U + off00 + P);  // u(i,0), a ring value

    int kbc = 0;
    unsigned par = 0;  // S & 1
    long long S = 0;
    // one step: consumes `cur` (requested during the previous step), requests `nxt`
    const bool nosync = kb < 0;  // timing experiment only (APDE_BAND_NOSYNC): no flag exchange, results are garbage
    if (nosync) kb = -kb;
    auto step = [&](BandOperands<R, LPF> &cur, BandOperands<R, LPF> &nxt) {
        if (kbc == 0 && !nosync) {
            // publish BEFORE polling: a CTA that is waiting for its predecessors must not withhold the steps it has
            // already completed from its successors (withholding adds kb steps of lag per hop around the ring of
            // slots -- measured as a deadlock at 15 slots x kb 8 on 300 columns)
            if (S > 0) {
                __syncthreads();  // every thread's stores of step S-1 are ordered before the release
                if (tid == NT - 1) st_release_u64(fme, (unsigned long long)S);
            }
            if (f_ptr != nullptr) {
                const long long need = min(S + kb + f_add, f_cap);
                if (need > 0)
                    while (ld_acquire_u64(f_ptr) < (unsigned long long)need) {
                    }
            }
        }
        __syncthreads();
        if (++kbc == kb) kbc = 0;

        // ---- every walker's next step, and the operands it will need
        int jxn[R], nn[R];
        unsigned offn[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int v = k * NT + tid;
            jxn[k] = jx[k] + 1;
            nn[k] = n[k];
            offn[k] = jx[k] >= 0 ? off[k] + P : OFF0(k);
            if (jxn[k] == CP) {
                jxn[k] = 0;
                offn[k] = OFF0(k);
                nn[k] = n[k] + 1;
            }
            if (v < rows_b && jxn[k] >= 0 && nn[k] < ntasks) {
                nxt.ue[k] = __ldcg(U + (offn[k] + P));
                if (LPF && HAS_NOISE && jxn[k] >= 1) nxt.le[k] = __ldg(LH + offn[k]);
                if (jxn[k] >= 2) {
                    if (lane == 31 || v == v_last) nxt.us[k] = __ldcg(U + (offn[k] + P + 1u));
                    if (k == 0 && tid == 0) nxt.un0 = __ldcg(U + (offn[k] - P - 1u));
                    if (LPF) {
                        if (HAS_NOISE) {
                            nxt.ls[k] = __ldg(LV + offn[k]);
                            nxt.ln[k] = __ldg(LV + (offn[k] - P - 1u));
                        }
                        if (HAS_SRC) nxt.sv[k] = __ldg(SRC + offn[k]);
                    } else {
                        if (HAS_NOISE) {
                            prefetch_l1_line(LV + offn[k]);
                            prefetch_l1_line(LH + offn[k]);
                        }
                        if (HAS_SRC) prefetch_l1_line(SRC + offn[k]);
                    }
                }
                if (APDE_BAND_L2AHEAD > 0 && (HAS_NOISE || HAS_SRC)) {
                    // same row, L2AHEAD columns on (past the row's end: padding or the first diagonals below, harmless)
                    const unsigned oa = offn[k] + (unsigned)APDE_BAND_L2AHEAD * P;
                    if (HAS_NOISE) {
                        prefetch_l2_line(LV + oa);
                        prefetch_l2_line(LH + oa);
                    }
                    if (HAS_SRC) prefetch_l2_line(SRC + oa);
                }
            }
        }
        // ---- this step
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int v = k * NT + tid;
            const bool act = v < rows_b && jx[k] >= 0 && n[k] < ntasks;
            double us = __shfl_down_sync(0xffffffffu, cur.ue[k], 1);
            double un = __shfl_up_sync(0xffffffffu, uw[k], 1);
            if (lane == 31 || v == v_last) us = cur.us[k];
            if (lane == 0) {
                if (warp > 0) un = nbuf[par ^ 1u][k][warp > 0 ? warp - 1 : 0];
                else if (k > 0) un = nbuf[par ^ 1u][k > 0 ? k - 1 : 0][NWARP - 1];
                else un = cur.un0;
            }
            double val = old[k];  // the boundary column passes through
            double2 le = make_double2(1.0, 1.0);
            if (act && jx[k] >= 1) {
                const unsigned o = off[k];
                if (LPF) le = cur.le[LPF ? k : 0];
                else if (HAS_NOISE && (LWREG || jx[k] >= 2)) le = __ldg(LH + o);
                if (jx[k] >= 2) {
                    double sv = 0.0;
                    if (LPF) sv = cur.sv[LPF ? k : 0];
                    else if (HAS_SRC) sv = __ldg(SRC + o);
                    if (HAS_NOISE) {
                        const double2 ln = LPF ? cur.ln[LPF ? k : 0] : __ldg(LV + (o - P - 1u));
                        const double2 ls = LPF ? cur.ls[LPF ? k : 0] : __ldg(LV + o);
                        const double2 lwk = LWREG ? lw[LWREG ? k : 0] : __ldg(LH + (o - P));
                        bool slow = false;
                        const double qn = band_div(un, ln, slow), qs = band_div(us, ls, slow);
                        const double qw = band_div(uw[k], lwk, slow), qe = band_div(cur.ue[k], le, slow);
                        double acc = __dadd_rn(__dadd_rn(__dadd_rn(qn, qs), qw), qe);  // solver.py:275, left-associative
                        if (HAS_SRC) acc = __dsub_rn(acc, sv);
                        val = __dmul_rn(acc, 0.25);
                        if (slow) val = relax_node_general<HAS_SRC>(un, us, uw[k], cur.ue[k], ln.x, ls.x, lwk.x, le.x, sv);
                    } else {
                        val = relax_node<false, HAS_SRC>(un, us, uw[k], cur.ue[k], le, le, le, le, sv);
                    }
                    mx[k] = max_ignore_nan(mx[k], fabs(__dsub_rn(val, old[k])));
                    __stcg(U + o, val);
                    if (jx[k] == CP - 1) {  // end of the row: this sweep's residual (solver.py:277-281)
                        atomic_max_nonneg(p.hist + (slot + (long long)n[k] * Gs), mx[k]);
                        mx[k] = 0.0;
                    }
                }
            }
            if (act) {
                uw[k] = val;
                old[k] = cur.ue[k];
                if (LWREG) lw[LWREG ? k : 0] = le;
            }
            if (lane == 31) nbuf[par][k][warp] = uw[k];
            jx[k] = jxn[k];
            off[k] = offn[k];
            n[k] = nn[k];
        }
        par ^= 1u;
        ++S;
    };
    while (S < Stot) {
        step(opA, opB);
        if (S < Stot) step(opB, opA);
    }
    __syncthreads();
    if (tid == NT - 1) st_release_u64(fme, (unsigned long long)Stot);
#undef OFF0
}

// ------------------------------------------------------------------------------------------
// host drivers
// ------------------------------------------------------------------------------------------
template <typename K>
static int max_dyn_smem(K kernel, size_t bytes) {
    APDE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return APDE_OK;
}

static size_t k1_smem_bytes(int64_t rows, int64_t cols, bool noise, bool src) {
    size_t n = (size_t)rows * cols;
    size_t d = n + (n & 1) + K1_MAX_BATCH;  // keep the link arrays 16-byte aligned
    if (noise) d += 2 * ((size_t)rows * (cols - 1) + (size_t)(rows - 1) * cols);
    if (src) d += n;
    return d * sizeof(double);
}

static int solve_small(double *grid, const double *nh, const double *nv, const double *src,
                       int64_t rows, int64_t cols, double tol, int64_t max_iter, double *hist,
                       int64_t *iters, double *ksec) {
    const bool noise = nh != nullptr, has_src = src != nullptr;
    const size_t n = (size_t)rows * cols;
    DevBuf d_grid, d_nh, d_nv, d_src, d_snap, d_hist, d_iters;
    APDE_TRY(d_grid.alloc(n * 8));
    APDE_TRY(d_snap.alloc(n * 8));
    APDE_TRY(d_hist.alloc((size_t)max_iter * 8));
    APDE_TRY(d_iters.alloc(8));
    APDE_CUDA(cudaMemcpy(d_grid.p, grid, n * 8, cudaMemcpyHostToDevice));
    if (noise) {
        const size_t nhn = (size_t)rows * (cols - 1), nvn = (size_t)(rows - 1) * cols;
        DevBuf tmp;
        APDE_TRY(tmp.alloc(std::max(nhn, nvn) * 8));
        APDE_TRY(d_nh.alloc(nhn * 16));
        APDE_TRY(d_nv.alloc(nvn * 16));
        APDE_CUDA(cudaMemcpy(tmp.p, nh, nhn * 8, cudaMemcpyHostToDevice));
        make_links_kernel<<<64, 256>>>(tmp.as<double>(), d_nh.as<double2>(), (long long)nhn);
        APDE_CUDA(cudaMemcpy(tmp.p, nv, nvn * 8, cudaMemcpyHostToDevice));
        make_links_kernel<<<64, 256>>>(tmp.as<double>(), d_nv.as<double2>(), (long long)nvn);
        APDE_CUDA(cudaGetLastError());
        APDE_CUDA(cudaDeviceSynchronize());
    }
    if (has_src) {
        APDE_TRY(d_src.alloc(n * 8));
        APDE_CUDA(cudaMemcpy(d_src.p, src, n * 8, cudaMemcpyHostToDevice));
    }
    K1Params p;
    p.grid = d_grid.as<double>();
    p.nh = d_nh.as<double2>();
    p.nv = d_nv.as<double2>();
    p.src = d_src.as<double>();
    p.snapshot = d_snap.as<double>();
    p.hist = d_hist.as<double>();
    p.iters_out = d_iters.as<long long>();
    p.rows = (int)rows;
    p.cols = (int)cols;
    p.max_iter = max_iter;
    p.tol = tol;
    const int nd = (int)(rows + cols - 5);
    p.batch = std::min(K1_MAX_BATCH, std::max(16, 2 * nd));

    const int hw = (int)(cols - 1) / 2;
    int bx = 1;
    while (bx < hw && bx < 32) bx <<= 1;
    int by = (int)std::min<int64_t>(rows - 2, 1024 / bx);
    if (by < 1) by = 1;
    dim3 block(bx, by);
    const size_t smem = k1_smem_bytes(rows, cols, noise, has_src);

    EventPair ev;
    APDE_TRY(ev.init());
#define K1_LAUNCH(N, S)                                                                 \
    do {                                                                                \
        APDE_TRY(max_dyn_smem(exact_small_kernel<N, S>, smem));                         \
        APDE_CUDA(cudaEventRecord(ev.a));                                               \
        exact_small_kernel<N, S><<<1, block, smem>>>(p);                                \
        APDE_CUDA(cudaEventRecord(ev.b));                                               \
    } while (0)
    if (noise && has_src) K1_LAUNCH(true, true);
    else if (noise) K1_LAUNCH(true, false);
    else if (has_src) K1_LAUNCH(false, true);
    else K1_LAUNCH(false, false);
#undef K1_LAUNCH
    APDE_CUDA(cudaGetLastError());
    APDE_CUDA(cudaEventSynchronize(ev.b));
    float ms = 0.f;
    APDE_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
    if (ksec) *ksec = ms * 1e-3;
    long long it = 0;
    APDE_CUDA(cudaMemcpy(&it, d_iters.p, 8, cudaMemcpyDeviceToHost));
    *iters = it;
    if (it > 0) APDE_CUDA(cudaMemcpy(hist, d_hist.p, (size_t)it * 8, cudaMemcpyDeviceToHost));
    APDE_CUDA(cudaMemcpy(grid, d_grid.p, n * 8, cudaMemcpyDeviceToHost));
    return APDE_OK;
}

struct RingCfg {
    int block = 1024;
    int npt = 0;       // 0 => K2 (generic kernel), else K2b with this many nodes per thread
    int kb = 8;        // diagonals per progress exchange (K2b) / steps per progress exchange (K2c)
    size_t smem = 0;
    int grid_max = 0;  // co-resident CTAs
    int band_nt = 0;   // > 0: K2c (row bands) with this many threads per CTA ...
    int band_r = 1;    // ... and rows per thread (band height = band_nt * band_r)
    int bands = 1;     // K2c: CTAs per sweep
    int slots_max = 0; // sweeps in flight at most (K2/K2b: grid_max)
};

template <bool N, bool S>
static const void *ring_fn(int npt) {
    // <= 0: generic kernel K2, the code is minus the block size.  Measured at 4096^2 with links (B200, round 2): the
    // 64-register shapes with one node in flight per thread (512 x 2 CTAs: 87 GLUP/s at K=1000; 1024 x 1: 58 at K=100)
    // beat the 128-register shapes with four nodes in flight (256 x 2: 71; 512 x 1: 62 / 43) -- more warps hide the
    // L2 latency better than more loads per warp; without links four nodes in flight fit 64 registers: 165 -> 205.
    switch (npt) {
        // K2c: 100000 * (links requested a step ahead) + 10000 * rows per thread + threads
        case 10064: return (const void *)exact_band_kernel<N, S, 64, 1, 8, false>;   // 64-thread shapes: the tests' way to
        case 20064: return (const void *)exact_band_kernel<N, S, 64, 2, 4, false>;   // many bands and slots on small grids
        case 30064: return (const void *)exact_band_kernel<N, S, 64, 3, 4, false>;
        case 110064: return (const void *)exact_band_kernel<N, S, 64, 1, 4, true>;
        case 120064: return (const void *)exact_band_kernel<N, S, 64, 2, 4, true>;
        case 20256: return (const void *)exact_band_kernel<N, S, 256, 2, 2, false>;
        case 40256: return (const void *)exact_band_kernel<N, S, 256, 4, 2, false>;
        case 10512: return (const void *)exact_band_kernel<N, S, 512, 1, 2, false>;
        case 110512: return (const void *)exact_band_kernel<N, S, 512, 1, 1, true>;
        case -2048:  // K2d (mismatched meshes only; ring_configure never asks for it otherwise)
            if constexpr (N) return (const void *)exact_ring_smem_kernel<S, 1>;
            else return (const void *)exact_ring_kernel<N, S, 1024, 1, 4>;
        case -1024: return (const void *)exact_ring_kernel<N, S, 1024, 1, N ? 1 : 4>;
        case -512: return (const void *)exact_ring_kernel<N, S, 512, 2, N ? 1 : 4>;
        case -256: return (const void *)exact_ring_kernel<N, S, 256, 2, 4>;
        case 1: return (const void *)exact_ring2_kernel<N, S, 1>;
        case 2: return (const void *)exact_ring2_kernel<N, S, 2>;
        case 4: return (const void *)exact_ring2_kernel<N, S, 4>;
        default: return (const void *)exact_ring2_kernel<N, S, 8>;
    }
}
static const void *ring_fn(bool noise, bool src, int npt) {
    if (noise && src) return ring_fn<true, true>(npt);
    if (noise) return ring_fn<true, false>(npt);
    if (src) return ring_fn<false, true>(npt);
    return ring_fn<false, false>(npt);
}

// `slots` = sweeps in flight (CTAs for K2/K2b; K2c launches slots * bands CTAs)
static int ring_dispatch(bool noise, bool src, const K2Params &p, const RingCfg &cfg, int slots, cudaStream_t st) {
    const void *fn = ring_fn(noise, src, cfg.npt);
    int kb = cfg.kb, bands = cfg.bands;
    if (cfg.band_nt > 0) {
        if (getenv("APDE_BAND_NOSYNC")) kb = -kb;
        void *args[] = {(void *)&p, (void *)&bands, (void *)&slots, (void *)&kb};
        APDE_CUDA(cudaLaunchCooperativeKernel(fn, dim3(slots * bands), dim3(cfg.block), args, 0, st));
        return APDE_OK;
    }
    void *args2[] = {(void *)&p, (void *)&kb};
    void *args1[] = {(void *)&p};
    APDE_CUDA(cudaLaunchCooperativeKernel(fn, dim3(slots), dim3(cfg.block), cfg.npt > 0 ? args2 : args1, cfg.smem, st));
    return APDE_OK;
}

// picks the kernel + launch shape; K2b whenever the longest diagonal fits 8 nodes/thread and the two
// diagonal buffers fit shared memory, else the generic K2
static int ring_configure(bool noise, bool src, int64_t rows, int64_t cols, int64_t sweeps_hint,
                          const cudaDeviceProp &prop, RingCfg *cfg) {
    const int64_t maxdiag = std::min(rows, cols) - 2;
    const char *force = getenv("APDE_RING_KERNEL");  // "k2" forces the generic kernel (cross-check)
    const size_t smem2 = (size_t)2 * (rows + 2) * sizeof(double);
    const bool can2 = maxdiag <= 8 * 1024 && smem2 + 1024 <= (size_t)prop.sharedMemPerBlockOptin;
    // Measured on B200 at 4096^2 (tools/exact_bench.py):
    //   many sweeps in flight (> 1 per SM): K2 with 512-thread CTAs, two per SM -- one CTA's L2 round
    //     trips overlap the other's arithmetic: 96 GLUP/s with mismatch, 178 without;
    //   at most one sweep per SM: 1024-thread CTAs; K2b (116) beats K2 (85) without mismatch, K2 (54)
    //     beats K2b (50) with it (the four divisions per node dominate there).
    const bool many = sweeps_hint > prop.multiProcessorCount;
    bool want2 = !noise && !many;
    if (force && !strcmp(force, "k2")) want2 = false;
    if (force && !strcmp(force, "k2b")) want2 = true;
    // K2d (operands of the sweep's own previous diagonal in shared memory): mismatched meshes whose 32 B per row +
    // 24 B per column fit one CTA's shared memory (4096^2: 224 KB)
    const size_t smem4 = k2d_smem_bytes(rows, cols);
    const bool can4 = noise && smem4 + 1024 <= (size_t)prop.sharedMemPerBlockOptin;
    bool want4 = K2D_DEFAULT;
    if (force) want4 = !strcmp(force, "k2d");
    // K2c (row bands): possible wherever its ring of sweep slots can be made short enough for the width
    // (slots * (3*kb + 4) <= cols, see the kernel) -- i.e. everything but very narrow grids
    int band_kb = 4;
    if (const char *v = getenv("APDE_BAND_KB")) band_kb = std::min(64, std::max(1, atoi(v)));
    while (band_kb > 1 && cols < 3 * band_kb + 4) band_kb /= 2;
    const bool can3 = cols >= 3 * band_kb + 4 && rows >= 3 && (rows + cols) * ((rows + 3) & ~(int64_t)3) < (1ll << 32);  // 32-bit element offsets
    // measured on B200 (4096^2, 1 % mismatch, K=1000): 48 GLUP/s against K2's 86 -- one node per thread per step cannot
    // amortise the ~100 instructions a step costs besides the node itself; opt-in only
    bool want3 = false;
    if (force) want3 = !strcmp(force, "band");
    if (can3 && want3) {
        // launch shapes the kernel is instantiated for: (threads, rows per thread, links requested a step ahead)
        static const int shapes[][3] = {{64, 1, 0}, {64, 2, 0}, {64, 3, 0}, {64, 1, 1}, {64, 2, 1},
                                        {256, 2, 0}, {256, 4, 0}, {512, 1, 0}, {512, 1, 1}};
        int nt = 512, r = 1, lpf = 0;  // the fastest measured shape (all shapes land within 15 %)
        if (const char *v = getenv("APDE_BAND_NT")) nt = atoi(v);
        if (const char *v = getenv("APDE_BAND_R")) r = atoi(v);
        if (const char *v = getenv("APDE_BAND_LPF")) lpf = atoi(v) != 0;
        bool known = false;
        for (const auto &sh : shapes) known = known || (sh[0] == nt && sh[1] == r && sh[2] == lpf);
        if (!known) {
            set_error("exact band kernel: no instantiation for APDE_BAND_NT=%d APDE_BAND_R=%d APDE_BAND_LPF=%d", nt, r, lpf);
            return APDE_EINVAL;
        }
        cfg->band_nt = nt;
        cfg->band_r = r;
        cfg->block = nt;
        cfg->npt = 100000 * lpf + 10000 * r + nt;
        cfg->kb = band_kb;
        cfg->smem = 0;
        cfg->bands = (int)((rows - 2 + nt * r - 1) / (nt * r));
    } else if (can4 && want4) {
        cfg->block = 1024;
        cfg->npt = -2048;
        cfg->smem = smem4;
    } else if (can2 && want2) {
        int block = 1024;
        while (block > 128 && block / 2 >= maxdiag) block /= 2;
        int npt = 1;
        while ((int64_t)block * npt < maxdiag) npt *= 2;
        cfg->block = block;
        cfg->npt = npt;
        cfg->smem = smem2;
        if (const char *v = getenv("APDE_RING_BATCH")) cfg->kb = std::max(1, atoi(v));
    } else {
        cfg->block = maxdiag > 2048 ? (many ? 512 : 1024) : (maxdiag > 512 ? 512 : 256);
        if (const char *v = getenv("APDE_RING_BLOCK")) {
            const int b = atoi(v);
            cfg->block = b >= 1024 ? 1024 : (b >= 512 ? 512 : 256);  // the block sizes the kernel is instantiated for
        }
        cfg->npt = -cfg->block;
        cfg->smem = 0;
    }
    const void *fn = ring_fn(noise, src, cfg->npt);
    if (cfg->smem > 48 * 1024)
        APDE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg->smem));
    int per_sm = 0;
    APDE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, cfg->block, cfg->smem));
    if (per_sm < 1) {
        set_error("exact ring kernel cannot be resident (block=%d smem=%zu)", cfg->block, cfg->smem);
        return APDE_ECUDA;
    }
    cfg->grid_max = per_sm * prop.multiProcessorCount;
    cfg->slots_max = cfg->grid_max;
    if (cfg->band_nt > 0) {
        cfg->slots_max = (int)std::min<int64_t>(cfg->grid_max / cfg->bands, cols / (3 * cfg->kb + 4));
        if (cfg->slots_max < 1) {
            set_error("exact band kernel: %d bands of %d rows exceed the %d co-resident CTAs", cfg->bands,
                      cfg->band_nt * cfg->band_r, cfg->grid_max);
            return APDE_EINVAL;
        }
    }
    return APDE_OK;
}

static int solve_ring(double *grid, const double *nh, const double *nv, const double *src,
                      int64_t rows, int64_t cols, double tol, int64_t max_iter, double *hist,
                      int64_t *iters, double *ksec, const cudaDeviceProp &prop) {
    const bool noise = nh != nullptr, has_src = src != nullptr;
    const size_t n = (size_t)rows * cols;
    const int P = (int)((rows + 3) & ~(int64_t)3);
    const size_t nd_total = (size_t)(rows + cols - 1);
    const size_t dbytes = nd_total * (size_t)P * 8;

    DevBuf stage, U, LH, LV, S, snap, d_hist, d_prog;
    APDE_TRY(stage.alloc(n * 8));
    APDE_TRY(U.alloc(dbytes));
    APDE_CUDA(cudaMemset(U.p, 0, dbytes));
    const int tb = 256, tg = prop.multiProcessorCount * 8;
    APDE_CUDA(cudaMemcpy(stage.p, grid, n * 8, cudaMemcpyHostToDevice));
    to_diag_kernel<<<tg, tb>>>(stage.as<double>(), U.as<double>(), (int)rows, (int)cols, P);
    if (noise) {
        APDE_TRY(LH.alloc(2 * dbytes));
        APDE_TRY(LV.alloc(2 * dbytes));
        // only valid links are ever read; the rest stays zero
        APDE_CUDA(cudaMemset(LH.p, 0, 2 * dbytes));
        APDE_CUDA(cudaMemset(LV.p, 0, 2 * dbytes));
        APDE_CUDA(cudaMemcpy(stage.p, nh, (size_t)rows * (cols - 1) * 8, cudaMemcpyHostToDevice));
        to_diag_links_kernel<<<tg, tb>>>(stage.as<double>(), LH.as<double2>(), (int)rows, (int)cols - 1, P);
        APDE_CUDA(cudaMemcpy(stage.p, nv, (size_t)(rows - 1) * cols * 8, cudaMemcpyHostToDevice));
        to_diag_links_kernel<<<tg, tb>>>(stage.as<double>(), LV.as<double2>(), (int)rows - 1, (int)cols, P);
    }
    if (has_src) {
        APDE_TRY(S.alloc(dbytes));
        APDE_CUDA(cudaMemset(S.p, 0, dbytes));
        APDE_CUDA(cudaMemcpy(stage.p, src, n * 8, cudaMemcpyHostToDevice));
        to_diag_kernel<<<tg, tb>>>(stage.as<double>(), S.as<double>(), (int)rows, (int)cols, P);
    }
    APDE_CUDA(cudaGetLastError());
    APDE_TRY(d_hist.alloc((size_t)max_iter * 8));

    RingCfg cfg;
    APDE_TRY(ring_configure(noise, has_src, rows, cols, max_iter, prop, &cfg));
    const int gmax = cfg.slots_max;  // sweeps in flight
    APDE_TRY(d_prog.alloc((size_t)cfg.grid_max * 16 * 8));

    const bool may_stop = tol > 0.0;
    if (may_stop) APDE_TRY(snap.alloc(dbytes));
    const int64_t nd = rows + cols - 5;
    // sweeps per launch when convergence has to be checked: enough to amortise the pipeline
    // fill/drain (2 diagonals per in-flight sweep) against nd diagonals per sweep
    int64_t chunk_cap = gmax;
    while (chunk_cap < 8 * (int64_t)gmax && (int64_t)(cfg.kb + 1) * chunk_cap > nd) chunk_cap *= 2;
    // K2c: the fill is one band stagger per band plus a few steps per slot, a task is `cols` steps
    if (cfg.band_nt > 0) chunk_cap = (int64_t)gmax * (cfg.bands > 2 ? 4 : 2);

    EventPair ev;
    APDE_TRY(ev.init());
    float total_ms = 0.f;
    K2Params p;
    p.U = U.as<double>();
    p.LH = LH.as<double2>();
    p.LV = LV.as<double2>();
    p.S = S.as<double>();
    p.progress = d_prog.as<unsigned long long>();
    p.rows = (int)rows;
    p.cols = (int)cols;
    p.P = P;
    // The flag exchange (acquire-poll of the predecessor, release of the own counter) is the LAST thread's job: it has
    // one node less than thread 0 on all but 1 in 32 diagonals, so the release fence sits in that slack instead of on
    // the block's critical path.  APDE_RING_FLAGTID=0 restores thread 0 (measurement aid).
    p.flag_tid = cfg.block - 1;
    if (const char *v = getenv("APDE_RING_FLAGTID")) p.flag_tid = std::min(cfg.block - 1, std::max(0, atoi(v)));
    p.dbg = 0;
    if (const char *v = getenv("APDE_RING_DEBUG")) p.dbg = atoi(v);  // timing experiments only: results are wrong
    p.l2_ahead = 4;
    if (const char *v = getenv("APDE_RING_L2AHEAD")) p.l2_ahead = std::min(64, std::max(0, atoi(v)));

    std::vector<double> hbuf;
    int64_t done = 0, result_iters = max_iter;
    while (done < max_iter) {
        const int64_t chunk = may_stop ? std::min<int64_t>(max_iter - done, chunk_cap) : (max_iter - done);
        // CTAs: as many as fit, but balanced over the rounds (sweeps c, c+G, c+2G ... per CTA): 300 sweeps on 296 slots
        // would run one full round plus a nearly serial round of 4; 2 rounds of 150 finish sooner
        const int64_t rounds = (chunk + gmax - 1) / gmax;
        const int g = (int)((chunk + rounds - 1) / rounds);
        if (may_stop) APDE_CUDA(cudaMemcpyAsync(snap.p, U.p, dbytes, cudaMemcpyDeviceToDevice, 0));
        APDE_CUDA(cudaMemsetAsync(d_prog.p, 0, d_prog.bytes, 0));
        p.hist = d_hist.as<double>() + done;
        APDE_CUDA(cudaMemsetAsync(p.hist, 0, (size_t)chunk * 8, 0));  // K2c folds the residual with atomicMax
        p.n_sweeps = chunk;
        APDE_CUDA(cudaEventRecord(ev.a));
        APDE_TRY(ring_dispatch(noise, has_src, p, cfg, g, 0));
        APDE_CUDA(cudaEventRecord(ev.b));
        APDE_CUDA(cudaEventSynchronize(ev.b));
        float ms = 0.f;
        APDE_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
        total_ms += ms;
        if (may_stop) {
            hbuf.resize((size_t)chunk);
            APDE_CUDA(cudaMemcpy(hbuf.data(), p.hist, (size_t)chunk * 8, cudaMemcpyDeviceToHost));
            int64_t stop = -1;
            for (int64_t k = 0; k < chunk; ++k)
                if (hbuf[(size_t)k] < tol) {  // solver.py:283
                    stop = k;
                    break;
                }
            if (stop >= 0) {
                if (stop != chunk - 1) {
                    // later sweeps of the chunk already ran: roll back, replay exactly stop+1 sweeps
                    APDE_CUDA(cudaMemcpyAsync(U.p, snap.p, dbytes, cudaMemcpyDeviceToDevice, 0));
                    APDE_CUDA(cudaMemsetAsync(d_prog.p, 0, d_prog.bytes, 0));
                    APDE_CUDA(cudaMemsetAsync(p.hist, 0, (size_t)(stop + 1) * 8, 0));
                    p.n_sweeps = stop + 1;
                    APDE_CUDA(cudaEventRecord(ev.a));
                    APDE_TRY(ring_dispatch(noise, has_src, p, cfg, (int)std::min<int64_t>(stop + 1, gmax), 0));
                    APDE_CUDA(cudaEventRecord(ev.b));
                    APDE_CUDA(cudaEventSynchronize(ev.b));
                    APDE_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
                    total_ms += ms;
                }
                result_iters = done + stop + 1;
                break;
            }
        }
        done += chunk;
    }
    *iters = result_iters;
    if (ksec) *ksec = total_ms * 1e-3;
    if (result_iters > 0)
        APDE_CUDA(cudaMemcpy(hist, d_hist.p, (size_t)result_iters * 8, cudaMemcpyDeviceToHost));
    from_diag_kernel<<<tg, tb>>>(U.as<double>(), stage.as<double>(), (int)rows, (int)cols, P);
    APDE_CUDA(cudaGetLastError());
    APDE_CUDA(cudaMemcpy(grid, stage.p, n * 8, cudaMemcpyDeviceToHost));
    return APDE_OK;
}

int selftest_division(int64_t n, uint64_t seed, int64_t *mismatches) {
    cudaDeviceProp prop;
    APDE_TRY(require_sm100(&prop));
    DevBuf cnt;
    APDE_TRY(cnt.alloc(8));
    APDE_CUDA(cudaMemset(cnt.p, 0, 8));
    division_selftest_kernel<<<prop.multiProcessorCount * 8, 256>>>((long long)n, (unsigned long long)seed,
                                                                  cnt.as<unsigned long long>());
    APDE_CUDA(cudaGetLastError());
    unsigned long long h = 0;
    APDE_CUDA(cudaMemcpy(&h, cnt.p, 8, cudaMemcpyDeviceToHost));
    *mismatches = (int64_t)h;
    return APDE_OK;
}

// entry used by the C-ABI (apde_api.cu)
int solve_exact_f64(double *grid, const double *nh, const double *nv, const double *src,
                    int64_t rows, int64_t cols, double tol, int64_t max_iter, double *hist,
                    int64_t *iters, double *ksec, int force_path) {
    if (ksec) *ksec = 0.0;
    if (max_iter <= 0) {  // range(max_iter) is empty: for...else sets iterations_run = max_iter
        *iters = max_iter;
        return APDE_OK;
    }
    if (rows < 3 || cols < 3) {
        // no interior node: every sweep has max_change == 0.0 (solver.py:264, :281-287)
        const int64_t n = (0.0 < tol) ? 1 : max_iter;
        for (int64_t k = 0; k < n; ++k) hist[k] = 0.0;
        *iters = n;
        return APDE_OK;
    }
    if (rows > (1 << 30) || cols > (1 << 30) || rows + cols > (1ll << 30)) {
        set_error("exact path: grid %lld x %lld too large", (long long)rows, (long long)cols);
        return APDE_EINVAL;
    }
    cudaDeviceProp prop;
    APDE_TRY(require_sm100(&prop));
    const size_t smem = k1_smem_bytes(rows, cols, nh != nullptr, src != nullptr);
    const bool fits = smem + 1024 <= (size_t)prop.sharedMemPerBlockOptin;
    bool small = fits;
    if (force_path == 1) {
        if (!fits) {
            set_error("exact path: %lld x %lld does not fit the one-CTA kernel", (long long)rows, (long long)cols);
            return APDE_EINVAL;
        }
        small = true;
    } else if (force_path == 2) {
        small = false;
    }
    if (small) return solve_small(grid, nh, nv, src, rows, cols, tol, max_iter, hist, iters, ksec);
    return solve_ring(grid, nh, nv, src, rows, cols, tol, max_iter, hist, iters, ksec, prop);
}

}  // namespace apde

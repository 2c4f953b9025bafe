// Exact two-sided two-sample Kolmogorov-Smirnov p-values from the integer statistics
// (n0, n1, h) the position kernel emits.
//
// Replaces scipy.stats._stats_py._attempt_exact_2kssamp (scipy/stats/_stats_py.py:7825-7870),
// which nanocompore reaches through scipy.stats.mstats.ks_twosamp (comparisons.py:222,
// 236-245):
//   n0 == n1: _compute_prob_outside_square (7713-7748), an O(n) Horner product
//   n0 != n1: _compute_outer_prob_inside_method (compiled in SciPy): the proportion of lattice
//             paths (0,0)->(m,n) that leave the band |i/m - j/n| < h/lcm,
//                 C(i,j) = (C(i-1,j)*i + C(i,j-1)*j) / (i+j),   C = 1 outside the band,
//             C(0,j) = 0 inside the band; the answer is C(m,n).
// Every cell is computed with the IEEE operations SciPy performs (two products and a sum
// without contraction, and a correctly rounded quotient -- obtained from a correctly rounded
// reciprocal plus one exact-remainder correction, see div_by_recip), so the result does not
// depend on how the cells are scheduled and is bitwise SciPy's.
//
// The recurrence is sequential along a row, so problems are binned by kind and cost:
//   SMALL  window <= 64 cells: one thread per problem, its window in shared memory
//          (interleaved by thread: conflict free); a warp holds problems of one cost bucket
//   EQUAL  one thread per problem
//   LARGE  one warp per problem: 64 rows (two per lane) advance together as a skewed wavefront (lane l is two
//          column behind lane l-1 and takes the cell above by shuffle); the row above the
//          64-row chunk lives in a shared-memory ring.  Warps fetch problems from a counter,
//          most expensive bucket first.
//   HUGE   (a band too wide for the ring, plus the LARGE problems that alone would outlast the
//          batch) one CTA per problem: 512 rows advance in lockstep, one barrier per step.
#include "ncb_internal.h"
#include "ncb_math.cuh"

namespace ncb {

__device__ __forceinline__ long long floordiv_ll(long long a, long long b) {  // b > 0
    long long q = a / b;
    if ((a % b) != 0 && a < 0) --q;
    return q;
}
__device__ __forceinline__ long long ceildiv_ll(long long a, long long b) {   // a >= 0, b > 0
    return (a + b - 1) / b;
}

#define KS_RECIP_TABLE 1024    /* reciprocals of i + j: SMALL problems have n0 + n1 < this */
/* A LARGE problem gets a whole CTA instead of a warp when it alone would outlast the batch: its
 * band cells exceed 1/KS_CTA_SHARE of all wavefront cells of the batch (about the number of warps
 * the wavefront kernel keeps resident), and at least KS_CTA_MIN_CELLS. */
#ifndef KS_CTA_SHARE
#define KS_CTA_SHARE 3072.0
#endif
#define KS_CTA_MIN_CELLS 131072.0
#ifndef KS_USE_STRIP
#define KS_USE_STRIP 1    /* 0: narrow bands stay on the chunked warp form (A/B builds) */
#endif
#ifndef KS_SMALL_TWO_ROWS
#define KS_SMALL_TWO_ROWS 1   /* thread-per-problem kernel: two rows per trip (0: one, for A/B builds) */
#endif
#define KS_SMALL_MAX_CELLS 16384.0   /* band cells a single thread may walk (~0.4 ms) */

struct KsProblem {
    int m, n;            // m >= n
    long long mg, ng, h;
};

__device__ __forceinline__ KsProblem make_problem(int n0, int n1, long long h) {
    KsProblem q;
    q.m = max(n0, n1);
    q.n = min(n0, n1);
    unsigned ga = (unsigned)q.m, gb = (unsigned)q.n;   // sample sizes fit 32 bits: no 64-bit remainders
    while (gb) { const unsigned t = ga % gb; ga = gb; gb = t; }
    const int g = (int)max(ga, 1u);
    q.mg = q.m / g;
    q.ng = q.n / g;
    q.h = h;
    return q;
}

__device__ __forceinline__ int cost_bucket(double cost) {
    // two buckets per octave, bucket 0 = most expensive
    int e;
    double f = frexp(fmax(cost, 1.0), &e);   // cost = f * 2^e, f in [0.5, 1)
    int b = 2 * e + (f >= 0.70710678118654752 ? 1 : 0);
    b = NCB_KS_BUCKETS - 1 - b;
    return b < 0 ? 0 : b;
}

__global__ void ks_classify_kernel(int64_t nprob, const int32_t *n0, const int32_t *n1,
                                   const int64_t *h, double *p_out, NcbKsWork w) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nprob;
         q += (int64_t)gridDim.x * blockDim.x) {
        int a = n0[q], b = n1[q];
        long long hh = h[q];
        int bin = 255;
        if (a <= 0 || b <= 0) {
            p_out[q] = ncb_nan();
        } else if (hh == 0) {
            p_out[q] = 1.0;
        } else if (a == b && hh > 0) {
            bin = NCB_KS_KIND_EQUAL * NCB_KS_BUCKETS + cost_bucket((double)a + (double)hh);
        } else if (min(a, b) > NCB_MAX_READS / 2 || hh < 0) {
            // outside what ncb_compare can emit (n0 + n1 <= NCB_MAX_READS): the wavefront
            // kernel's row buffer would not hold the smaller sample
            p_out[q] = ncb_nan();
        } else {
            KsProblem k = make_problem(a, b, hh);
            const double lcm = (double)k.mg * (double)k.n;
            const double d = (double)hh / lcm;
            // P(D >= d) is of the order of exp(-2 d^2 mn/(m+n)); beyond exp(-921) = 1e-400 the
            // lattice recurrence underflows to exactly 0 (smallest denormal 4.9e-324, 76 orders
            // of magnitude of margin), which is also what SciPy returns: skip the walk.
            if (2.0 * d * d * (double)k.m * (double)k.n / (double)(k.m + k.n) > 921.0) {
                p_out[q] = 0.0;
            } else {
                long long maxj0 = min(ceildiv_ll(k.h, k.mg), (long long)k.n + 1);
                long long lenA = min(2 * maxj0 + 2, (long long)k.n + 1);
                double width = fmin(2.0 * (double)k.h / (double)k.mg + 1.0, (double)k.n + 1.0);
                // columns a 64-row chunk of the wavefront touches, plus those entering next
                double span = width + 2.0 * (64.0 * (double)k.ng / (double)k.mg) + 8.0;
                int kind;
                // lenA < window: a row has at most lenA - 1 cells and the band moves by at most one
                // column per row, so the thread kernel never reads past its window
                // (and a thread walks its problem alone: the long ones go to a warp)
                if (lenA < NCB_KS_SMALL_WINDOW && a + b < KS_RECIP_TABLE && (double)k.m * width <= KS_SMALL_MAX_CELLS)
                    kind = NCB_KS_KIND_SMALL;
                else if (k.n + 1 <= NCB_KS_RING_SMALL || span <= (double)NCB_KS_RING_SMALL) kind = NCB_KS_KIND_LARGE;
                else kind = NCB_KS_KIND_HUGE;
                // thread-per-problem kind: ordered by window (widest first), so that the problems of a
                // block fit the same number of slots and the threads of a warp walk rows of like length
                bin = kind * NCB_KS_BUCKETS + (kind == NCB_KS_KIND_SMALL ? NCB_KS_SMALL_WINDOW - 1 - (int)lenA
                                                                         : cost_bucket((double)k.m * width));
            }
        }
        w.bin[q] = (uint8_t)bin;
        // one atomic per (warp, bin): the problems of a batch fall into a handful of bins
        const unsigned peers = __match_any_sync(__activemask(), bin);
        if (bin != 255 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&w.hist[bin], __popc(peers));
    }
}

// representative cost (band cells) of a bucket: the inverse of cost_bucket
__device__ __forceinline__ double bucket_cost(int b) {
    return exp2(0.5 * (double)(NCB_KS_BUCKETS - 1 - b) - 0.75);   // middle of [2^(b'/2 - 1), 2^(b'/2 - 1/2))
}

// exclusive scan of the NCB_KS_NBINS counters (one block of NCB_KS_NBINS threads), and the split of
// the LARGE problems between the CTA kernel (the dearest ones, a prefix of their cost-sorted
// list) and the warp kernel: w.fetch[NCB_KS_KINDS] = length of that prefix
__global__ void ks_scan_kernel(NcbKsWork w) {
    __shared__ int s[NCB_KS_NBINS];
    __shared__ int raw[NCB_KS_NBINS];
    int t = threadIdx.x;
    int v = w.hist[t];
    s[t] = v;
    raw[t] = v;
    __syncthreads();
    for (int d = 1; d < NCB_KS_NBINS; d <<= 1) {
        int y = t >= d ? s[t - d] : 0;
        __syncthreads();
        s[t] += y;
        __syncthreads();
    }
    int excl = s[t] - v;
    w.hist[t] = excl;
    w.cursor[t] = excl;
    if (t == NCB_KS_NBINS - 1) w.hist[NCB_KS_NBINS] = s[t];
    if (t < NCB_KS_KINDS) w.fetch[t] = 0;
    if (t == 0) {
        double total = 0.0;
        for (int b = 0; b < NCB_KS_BUCKETS; ++b)
            total += bucket_cost(b) * (double)(raw[NCB_KS_KIND_LARGE * NCB_KS_BUCKETS + b] +
                                               raw[NCB_KS_KIND_HUGE * NCB_KS_BUCKETS + b]);
        const double cut = fmax(total / KS_CTA_SHARE, KS_CTA_MIN_CELLS);
        int n_cut = 0;
        for (int b = 0; b < NCB_KS_BUCKETS && bucket_cost(b) >= cut; ++b)
            n_cut += raw[NCB_KS_KIND_LARGE * NCB_KS_BUCKETS + b];
        w.fetch[NCB_KS_KINDS] = n_cut;
    }
}

__global__ void ks_scatter_kernel(int64_t nprob, NcbKsWork w) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nprob;
         q += (int64_t)gridDim.x * blockDim.x) {
        int bin = w.bin[q];
        const unsigned peers = __match_any_sync(__activemask(), bin);
        if (bin != 255) {
            const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&w.cursor[bin], __popc(peers));
            base = __shfl_sync(peers, base, leader);
            w.list[base + __popc(peers & ((1u << lane) - 1u))] = (int32_t)q;
        }
    }
}

// a / b for an integer-valued b with r = RN(1/b): q = RN(a r), then one correction through the
// exact remainder.  With a correctly rounded reciprocal this IS the correctly rounded quotient
// (Markstein), i.e. the same bits as SciPy's true division, for 3 fp64 operations instead of
// the ~12 of a division.  (Not valid in the denormal range, p < 1e-290.)
__device__ __forceinline__ double div_by_recip(double a, double b, double r) {
    const double q = __dmul_rn(a, r);
    const double rem = __fma_rn(-q, b, a);
    return __fma_rn(rem, r, q);
}

// ---- thread per problem: unequal sizes, small window ------------------------------------
#define KS_SMALL_THREADS 128
// WINDOW slots of shared memory per problem.  Most problems of a batch at ordinary coverage have a
// band of fewer than 32 columns (under the null D n ~ 15 at 100 reads per condition): a block whose
// problems all fit 32 slots runs in the 32-slot instance, of which an SM holds twice as many (the
// kernel waits on its dependent chain of five fp64 operations per cell: residency is throughput).
// Both instances are launched over all blocks; a block leaves the one that is not its own.
template <int WINDOW>
__global__ void __launch_bounds__(KS_SMALL_THREADS)
ks_small_kernel(const int32_t *n0, const int32_t *n1, const int64_t *h, double *p_out,
                NcbKsWork w) {
    extern __shared__ double A_all[];   // [WINDOW][KS_SMALL_THREADS]
    const double *__restrict__ rtab = w.rtab;    // reciprocals of the integers (per handle, L1 resident)
    const int lo = w.hist[NCB_KS_KIND_SMALL * NCB_KS_BUCKETS];
    const int hi = w.hist[(NCB_KS_KIND_SMALL + 1) * NCB_KS_BUCKETS];
    if (lo + (int)blockIdx.x * KS_SMALL_THREADS >= hi) return;   // whole block idle
    const int idx = lo + blockIdx.x * KS_SMALL_THREADS + threadIdx.x;
    const bool live = idx < hi;
    const int q = live ? w.list[idx] : 0;
    KsProblem k = make_problem(1, 2, 1);
    bool wide = false;
    if (live) {
        k = make_problem(n0[q], n1[q], h[q]);
        const long long maxj0 = min(ceildiv_ll(k.h, k.mg), (long long)k.n + 1);
        wide = min(2 * maxj0 + 2, (long long)k.n + 1) >= 32;      // lenA of the classification
    }
    const bool block_wide = __syncthreads_or(wide) != 0;
    if (block_wide != (WINDOW == NCB_KS_SMALL_WINDOW) || !live) return;
    double *A = A_all + threadIdx.x;
#define AT(j) A[(j) * KS_SMALL_THREADS]
    const int n = k.n, m = k.m;
    // every quantity fits in 32 bits (n0 + n1 < KS_RECIP_TABLE here)
    const int mg = (int)k.mg, ng = (int)k.ng, hh = (int)k.h;
    int minj = 0, maxj = min((hh + mg - 1) / mg, n + 1);
    int curlen = maxj - minj;
    for (int j = 0; j < WINDOW; ++j) AT(j) = (j < maxj) ? 0.0 : 1.0;
    // band of row i: minj = floor((ng i - h) / mg) + 1 and maxj = ceil((ng i + h) / mg), clamped.
    // Both numerators grow by ng <= mg per row, so the quotients advance by 0 or 1: they are
    // carried as (quotient, remainder) pairs -- no division in the row loop.
    int q_lo = -((hh + mg - 1) / mg), r_lo = q_lo * -mg - hh;          // floor(-h / mg), remainder in [0, mg)
    int q_hi = (hh + mg - 1) / mg, r_hi = (hh + mg - 1) - q_hi * mg;   // floor((h + mg - 1) / mg)
    double result = -1.0;
    int i = 1;
#if KS_SMALL_TWO_ROWS
    // Two rows per trip.  Row b = i + 1 runs one column behind row a = i: the cell above b's is a's
    // cell of the step before (a register), so the two cells of a step are independent -- two
    // dependent chains of five fp64 operations in flight per thread instead of one, which is what this
    // latency-bound kernel lacks (its residency is fixed by the windows in shared memory) -- and they
    // share i + j, its reciprocal and the loop.  Only row b is stored.  Same cells, same operations
    // per cell: the bits of the one-row form.
    for (; i + 1 <= m; i += 2) {
        const int minj_p = minj, len_p = curlen;
        r_lo += ng;
        if (r_lo >= mg) { r_lo -= mg; ++q_lo; }
        r_hi += ng;
        if (r_hi >= mg) { r_hi -= mg; ++q_hi; }
        const int minj_a = min(max(q_lo + 1, 0), n), maxj_a = min(q_hi, n + 1);
        if (maxj_a <= minj_a) { result = 1.0; break; }
        r_lo += ng;
        if (r_lo >= mg) { r_lo -= mg; ++q_lo; }
        r_hi += ng;
        if (r_hi >= mg) { r_hi -= mg; ++q_hi; }
        const int minj_b = min(max(q_lo + 1, 0), n), maxj_b = min(q_hi, n + 1);
        minj = minj_b; maxj = maxj_b;
        if (maxj_b <= minj_b) { result = 1.0; break; }
        const int len_a = maxj_a - minj_a, len_b = maxj_b - minj_b, da = minj_b - minj_a;
        const double dia = (double)i, dib = dia + 1.0;
        double va = (minj_a == 0) ? 0.0 : 1.0, vb = (minj_b == 0) ? 0.0 : 1.0;
        double dja = (double)minj_a, djb = (double)minj_b;
        double a_col = 1.0;                  // row a at the column row b computes next (1 right of a's band)
        const double *rt = rtab + i + minj_a;
        const int shift = minj_a - minj_p;
        double up = AT(shift);
        double r = __ldg(rt);
        int t = 0;
        // row a alone: the columns row b does not start at
        const int t1 = min(da + 1, len_a);
        for (; t < t1; ++t) {
            const double up_next = AT(shift + t + 1);
            const double r_next = __ldg(rt + t + 1);
            va = div_by_recip(NCB_DADD(NCB_DMUL(up, dia), NCB_DMUL(va, dja)), dia + dja, r);
            if (t == da) a_col = va;
            dja += 1.0;
            up = up_next;
            r = r_next;
        }
        // both rows: step t computes a's column minj_a + t and b's column minj_a + t - 1 (same i + j)
        for (; t < len_a; ++t) {
            const double up_next = AT(shift + t + 1);
            const double r_next = __ldg(rt + t + 1);
            const double sum = dia + dja;
            const double na = div_by_recip(NCB_DADD(NCB_DMUL(up, dia), NCB_DMUL(va, dja)), sum, r);
            vb = div_by_recip(NCB_DADD(NCB_DMUL(a_col, dib), NCB_DMUL(vb, djb)), sum, r);
            AT(t - da - 1) = vb;
            va = na;
            a_col = na;
            dja += 1.0;
            djb += 1.0;
            up = up_next;
            r = r_next;
        }
        // row b alone: its last one or two columns (the cell above the very last may lie right of a's band)
        {
            int u = max(len_a - da - 1, 0);
            const double *rb = rtab + (i + 1) + minj_b;
            if (len_a <= da) a_col = 1.0;
            for (; u < len_b; ++u) {
                vb = div_by_recip(NCB_DADD(NCB_DMUL(a_col, dib), NCB_DMUL(vb, djb)), dib + djb, __ldg(rb + u));
                AT(u) = vb;
                djb += 1.0;
                a_col = 1.0;
            }
        }
        curlen = len_b;
        for (int jj = curlen; jj < len_p; ++jj) AT(jj) = 1.0;
    }
#endif
    for (; result < 0.0 && i <= m; ++i) {
        const int lastminj = minj, lastlen = curlen;
        r_lo += ng;
        if (r_lo >= mg) { r_lo -= mg; ++q_lo; }
        r_hi += ng;
        if (r_hi >= mg) { r_hi -= mg; ++q_hi; }
        minj = min(max(q_lo + 1, 0), n);
        maxj = min(q_hi, n + 1);
        if (maxj <= minj) { result = 1.0; break; }
        double val = (minj == 0) ? 0.0 : 1.0;
        const double di = (double)i;
        const int shift = minj - lastminj;
        const int len = maxj - minj;
        double dj = (double)minj;
        const double *rt = rtab + i + minj;
        // the cell above and the reciprocal of the next column are fetched before this column's
        // store (they never alias it: src > jj), so memory latency stays off the
        // loop-carried chain val -> val
        // (src <= len + 1 < window: see the classification; slots right of the band hold 1.0)
        double up = AT(shift);
        double r = __ldg(rt);
        for (int jj = 0; jj < len; ++jj) {
            const int src = jj + 1 + shift;
            const double up_next = AT(src);
            const double r_next = __ldg(rt + jj + 1);
            const double num = NCB_DADD(NCB_DMUL(up, di), NCB_DMUL(val, dj));
            val = div_by_recip(num, di + dj, r);
            AT(jj) = val;
            dj += 1.0;
            up = up_next;
            r = r_next;
        }
        curlen = len;
        for (int jj = curlen; jj < lastlen; ++jj) AT(jj) = 1.0;
    }
    if (result < 0.0) result = AT(maxj - minj - 1);
#undef AT
    p_out[q] = clip01(result);
}

// correctly rounded reciprocals of the integers, once per handle: the wavefront kernel takes
// 1 / (i + j) from here (L1-resident, loaded a step ahead) instead of computing it on the
// critical path of every step
__global__ void ks_rtab_kernel(double *rtab) {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < NCB_KS_RTAB_SIZE; b += gridDim.x * blockDim.x)
        rtab[b] = b ? __drcp_rn((double)b) : 0.0;
}

// ---- thread per problem: equal sizes -----------------------------------------------------
// The Horner product of ks_prob_outside_square (ncb_math.cuh) with the bits of its n divisions
// but not their cost: numerators and denominators run as doubles (exact integers), the quotient
// by the integer n + k h + j + 1 comes from the handle's table of correctly rounded reciprocals
// (div_by_recip: the correctly rounded quotient).  The quotient trick needs normal numbers, so a
// factor that has fallen below 1e-270 finishes its block with true divisions; a factor that has
// reached zero stays zero.  At 5000 reads per condition this is 40 instead of ~1300 cycles per
// term (64-bit integer conversions and a division in the chain).
__device__ double ks_prob_outside_square_fast(int n, int h, const double *__restrict__ rtab) {
    double P = 0.0;
    for (int k = n / h; k >= 0; --k) {
        double p1 = 1.0;
        double a = (double)(n - k * h);            // numerator n - k h - j
        int b = n + k * h + 1;                     // denominator n + k h + j + 1
        double bd = (double)b;
        int j = 0;
        for (; j < h && p1 >= 1e-270; ++j) {
            p1 = div_by_recip(NCB_DMUL(a, p1), bd, __ldg(rtab + b));
            a -= 1.0; bd += 1.0; ++b;
        }
        for (; j < h && p1 != 0.0; ++j) {
            p1 = NCB_DDIV(NCB_DMUL(a, p1), bd);
            a -= 1.0; bd += 1.0;
        }
        if (j < h) p1 = 0.0;                       // (+-0 either way: only 1 - P is taken from it)
        P = NCB_DMUL(p1, NCB_DADD(1.0, -P));
    }
    return 2.0 * P;
}

__global__ void ks_equal_kernel(const int32_t *n0, const int64_t *h, double *p_out, NcbKsWork w) {
    const int lo = w.hist[NCB_KS_KIND_EQUAL * NCB_KS_BUCKETS];
    const int hi = w.hist[(NCB_KS_KIND_EQUAL + 1) * NCB_KS_BUCKETS];
    const int idx = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= hi) return;
    const int q = w.list[idx];
    const long long n = n0[q], hh = h[q];
    // (the unit entry point accepts samples beyond the table: those take the plain form)
    if (2 * n + 1 < NCB_KS_RTAB_SIZE && hh <= n)
        p_out[q] = ks_equal_finish(ks_prob_outside_square_fast((int)n, (int)hh, w.rtab), n, hh);
    else
        p_out[q] = ks_equal_p(n, hh);
}

// ---- wavefront kernels ------------------------------------------------------------------------
// One step of a wavefront computes one cell per row: row i = i0 + r is at column j = jlo - r + t
// at step t, so i + j = i0 + jlo + t is the same for every row (one reciprocal per step, read from
// the handle's table a step ahead) and the cell above is what row r - 1 computed in the step
// before.  Everything a step needs besides the two neighbours is carried incrementally: the
// column as an int (ring index) and as a double, the divisor, the band coordinate
// v = ng i - mg j (inside the band iff |v| < h).  The step is branch free: every lane computes,
// the selects decide what is kept (outside the band a cell is 1, outside the chunk's column range
// nothing changes).
// The rings are addressed with 32-bit shared-window addresses computed once per problem: through a
// generic pointer the compiler rebuilt the window base (S2R SR_CgaCtaId, shifts, adds) next to every
// predicated access, 14 of the 51 instructions of a step.
// (opaque: the compiler otherwise treats the address as rematerialisable and rebuilds it at every use)
__device__ __forceinline__ unsigned smem_addr(const void *p) {
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));
    return a;
}
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(unsigned addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

// Two rows per thread: thread r of a chunk holds rows a = 2 r and b = 2 r + 1.  Row rho is at
// column jlo - rho + t at step t, so b runs one column behind a and takes the cell above from
// a's value of the step before -- the two cells of a step are independent of each other (two
// chains to interleave), share the divisor i + j and its reciprocal, the loop, the exchange with
// the neighbouring thread and the table / ring reads: 24-26 instead of 37 instructions per 32 cells.
struct WavePair {
    double di_a, di_b, dj_a, dj_b, den, a, b;   // a, b: the rows' latest cells (also their `left`)
    int j, jr, va, vb;      // a's column, a's column - jlo, band coordinates ng i - mg j + h - 1
                            // (a cell is inside the band iff 0 <= v < 2 h - 1); b is at column j - 1
    unsigned span_a, span_b; // jhi - jlo for a live row, 0 for a row past the last one
    unsigned keep_a, keep_b; // the same for the chunk's last row (its cells go to the row buffer), else 0
};
// (the compiler otherwise re-derives "is this row live / the last one" from the lane index in
// every step instead of keeping the folded bound in a register)
__device__ __forceinline__ unsigned opaque(unsigned x) {
    asm volatile("" : "+r"(x));
    return x;
}

__device__ __forceinline__ void wave_pair_init(WavePair &r, int r_idx, int rows, int i0, int jlo, int jhi,
                                               int ng, int mg, int hh) {
    const int ia = i0 + 2 * r_idx;
    r.di_a = (double)ia;
    r.di_b = (double)(ia + 1);
    r.j = jlo - 2 * r_idx;
    r.jr = -2 * r_idx;
    r.dj_a = (double)r.j;
    r.dj_b = (double)(r.j - 1);
    r.den = (double)(i0 + jlo);
    r.va = ng * ia - mg * r.j + hh - 1;
    r.vb = ng * (ia + 1) - mg * (r.j - 1) + hh - 1;
    r.span_a = opaque(2 * r_idx < rows ? (unsigned)(jhi - jlo) : 0u);
    r.span_b = opaque(2 * r_idx + 1 < rows ? (unsigned)(jhi - jlo) : 0u);
    r.keep_a = opaque(2 * r_idx == rows - 1 ? (unsigned)(jhi - jlo) : 0u);
    r.keep_b = opaque(2 * r_idx + 1 == rows - 1 ? (unsigned)(jhi - jlo) : 0u);
    r.a = 1.0;
    r.b = 1.0;
}

// the two cells of this step; `up` = the cell above row a (row b of the thread before, one step
// ago).  Returns through act_a / act_b whether the rows are inside the chunk's column range.
__device__ __forceinline__ void wave_pair_cells(WavePair &r, double up, double rden, int hh, bool &act_a, bool &act_b) {
    act_a = (unsigned)r.jr < r.span_a;
    act_b = (unsigned)(r.jr - 1) < r.span_b;
    const unsigned lim = (unsigned)(2 * hh - 1);
    const bool in_a = (unsigned)r.va < lim, in_b = (unsigned)r.vb < lim;
    const double val_b = div_by_recip(NCB_DADD(NCB_DMUL(r.a, r.di_b), NCB_DMUL(r.b, r.dj_b)), r.den, rden);
    const double val_a = div_by_recip(NCB_DADD(NCB_DMUL(up, r.di_a), NCB_DMUL(r.a, r.dj_a)), r.den, rden);
    const double xb = in_b ? val_b : 1.0, xa = in_a ? val_a : 1.0;
    r.b = act_b ? xb : r.b;
    r.a = act_a ? xa : r.a;
}
__device__ __forceinline__ void wave_pair_advance(WavePair &r, int mg) {
    ++r.j; ++r.jr; r.va -= mg; r.vb -= mg;
    r.dj_a += 1.0; r.dj_b += 1.0; r.den += 1.0;
}

// reciprocal of k: from the table, or (unit entry point, samples beyond the table) computed
__device__ __forceinline__ double wave_recip(const double *__restrict__ rtab, int k) {
    return k < NCB_KS_RTAB_SIZE ? __ldg(rtab + k) : __drcp_rn((double)k);
}

// ---- warp per problem: skewed wavefront over 64 rows --------------------------------------
// The row above the current 64-row chunk lives in a power-of-two ring of RING doubles indexed
// by j & (RING - 1).  The band moves to the right, so a slot is reused only by columns RING
// apart; classification guarantees RING >= the widest span a chunk touches (or RING > n, in
// which case the ring never wraps).  Slots that enter the band for the first time must read 1
// ("outside the band"): they are reset at the end of the chunk before.  Lane 0 reads the ring a
// step ahead (the slot was written in the chunk before; the last lane of this chunk writes
// `rows - 1` columns behind, never the slot being read).
template <int RING, bool IN_TABLE>
__device__ __forceinline__ void warp_chunk(double *prev, const double *__restrict__ rtab, int lane, int rows,
                                           int i0, int jlo, int jhi, int ng, int mg, int hh) {
    constexpr int MASK = RING - 1;
    WavePair r;
    wave_pair_init(r, lane, rows, i0, jlo, jhi, ng, mg, hh);
    const int nsteps = (jhi - jlo) + rows - 1;
    const int k0 = i0 + jlo;
    const double *rp = rtab + k0;
    double rden = IN_TABLE ? __ldg(rp) : wave_recip(rtab, k0);
    const bool first = lane == 0;
    const unsigned base = smem_addr(prev);
    double up0 = first ? lds_f64(base + 8u * (unsigned)(r.j & MASK)) : 1.0;
#pragma unroll 2
    for (int t = 0; t < nsteps; ++t) {
        const double rden_next = IN_TABLE ? __ldg(rp + t + 1) : wave_recip(rtab, k0 + t + 1);
        double up0_next = 1.0;
        if (first) up0_next = lds_f64(base + 8u * (unsigned)((r.j + 1) & MASK));
        const double up_sh = __shfl_up_sync(0xffffffffu, r.b, 1);
        bool act_a, act_b;
        wave_pair_cells(r, first ? up0 : up_sh, rden, hh, act_a, act_b);
        if ((unsigned)r.jr < r.keep_a) sts_f64(base + 8u * (unsigned)(r.j & MASK), r.a);
        if ((unsigned)(r.jr - 1) < r.keep_b) sts_f64(base + 8u * (unsigned)((r.j - 1) & MASK), r.b);
        wave_pair_advance(r, mg);
        rden = rden_next;
        up0 = up0_next;
    }
}

template <int RING>
__device__ void warp_problem(double *prev, const double *__restrict__ rtab, int lane, const KsProblem &k,
                             double *p_out_q) {
    constexpr int MASK = RING - 1;
    static_assert((RING & (RING - 1)) == 0, "ring size must be a power of two");
    // every quantity below fits in 32 bits: n0 + n1 <= NCB_MAX_READS, so ng*i, mg*j and
    // h <= lcm are at most 5120 * 5120
    const int n = k.n, m = k.m;
    const int mg = (int)k.mg, ng = (int)k.ng, hh = (int)k.h;
    const int maxj0 = min((hh + mg - 1) / mg, n + 1);
    const int fill = min(n + 1, RING);
    for (int j = lane; j < fill; j += 32) prev[j] = (j < maxj0) ? 0.0 : 1.0;
    __syncwarp();
    int jhi_done = fill;                      // columns below this hold valid ring slots
    const bool in_table = m + n + 2 < NCB_KS_RTAB_SIZE;
    for (int i0 = 1; i0 <= m; i0 += 64) {
        const int rows = min(64, m - i0 + 1);
        const int ilast = i0 + rows - 1;
        const int a0 = ng * i0 - hh;            // floor(a0 / mg) + 1, clamped to [0, n]
        const int jlo = min(max((a0 >= 0 ? a0 / mg : -((-a0 + mg - 1) / mg)) + 1, 0), n);
        const int jhi = min((ng * ilast + hh + mg - 1) / mg, n + 1);
        // slots entering the band in this chunk that were last used RING columns ago
        for (int j = jhi_done + lane; j < jhi; j += 32) prev[j & MASK] = 1.0;
        jhi_done = max(jhi_done, jhi);
        __syncwarp();
        if (in_table) warp_chunk<RING, true>(prev, rtab, lane, rows, i0, jlo, jhi, ng, mg, hh);
        else warp_chunk<RING, false>(prev, rtab, lane, rows, i0, jlo, jhi, ng, mg, hh);
        __syncwarp();
    }
    if (lane == 0) *p_out_q = clip01(prev[n & MASK]);
    __syncwarp();
}

// ---- warp per problem, narrow bands: lanes own columns ------------------------------------------
// A chunked wavefront keeps its lanes busy w / (w + rows + drift) of the time: 55 % for the median
// band of a 5000-read position (155 columns), because every 64-row chunk starts and drains at the
// band's edges.  Here lane l owns the C columns [l C, l C + C) of the band RELATIVE to its left
// edge, for every row, and runs l rows behind lane l - 1: in period p it computes its strip of row
// i = p - l + 1, one cell per step, so apart from the first and last 31 periods of a problem every
// lane computes a cell in every step.  The band's left edge moves by s = 0 or 1 columns per row,
// so the cell above relative column r is relative column r + s of the row before: the lane's own
// previous strip (registers), and for the last column of a strip with s = 1 the FIRST column of
// lane l + 1's strip of that row -- which lane l + 1 computes in the first step of the same
// period (shuffle after that step).  The cell to the left of a strip's first column is the last
// column of lane l - 1's strip of the same row, finished in the step before the period starts.
// Rows carry their exact band limits ((quotient, remainder) pairs, no division), so there is no
// in-band test, and outside the band a cell above reads as 1.  Same cell arithmetic as the other
// forms, hence the same bits (schedule checked against the oracle by a lane-level emulation).
template <int C>
__device__ void strip_problem(const double *__restrict__ rtab, int lane, const KsProblem &k, double *p_out_q) {
    const int n = k.n, m = k.m;
    const int mg = (int)k.mg, ng = (int)k.ng, hh = (int)k.h;
    double U[C], V[C];          // the strip of the row before / of the current row
#pragma unroll
    for (int q = 0; q < C; ++q) { U[q] = 0.0; V[q] = 0.0; }      // row 0: 0 inside the band
    int prev_min = 0, prev_w = min((hh + mg - 1) / mg, n + 1);    // row 0: [0, maxj0)
    // band of row i: minj = floor((ng i - h) / mg) + 1, maxj = floor((ng i + h + mg - 1) / mg), clamped
    int q_lo = -((hh + mg - 1) / mg), r_lo = q_lo * -mg - hh;    // floor(-h / mg), remainder in [0, mg)
    int q_hi = (hh + mg - 1) / mg, r_hi = (hh + mg - 1) - q_hi * mg;
    const int rel0 = lane * C;
    int i = -lane;
    const int periods = m + 31;
#pragma unroll 1
    for (int p = 0; p < periods; ++p) {
        ++i;
        const bool act = i >= 1 && i <= m;
        int minj = prev_min, w = 0;
        if (act) {
            r_lo += ng;
            if (r_lo >= mg) { r_lo -= mg; ++q_lo; }
            r_hi += ng;
            if (r_hi >= mg) { r_hi -= mg; ++q_hi; }
            minj = min(max(q_lo + 1, 0), n);
            w = max(min(q_hi, n + 1) - minj, 0);
        }
        const bool shifted = minj != prev_min;                    // s = 1
        const int up_lim = prev_w - (shifted ? 1 : 0);            // the cell above rel exists iff rel < up_lim
        const double di = (double)i;
        double dj = (double)(minj + rel0);
        const double *rp = rtab + (act ? i + minj + rel0 : 0);
        double left = __shfl_up_sync(0xffffffffu, V[C - 1], 1);
        if (lane == 0) left = 1.0;
        double v0_next = 1.0;
#pragma unroll
        for (int q = 0; q < C; ++q) {
            double up = shifted ? (q + 1 < C ? U[q + 1 < C ? q + 1 : q] : v0_next) : U[q];
            up = (rel0 + q < up_lim) ? up : 1.0;
            const double rden = __ldg(rp + q);
            const double val = div_by_recip(NCB_DADD(NCB_DMUL(up, di), NCB_DMUL(left, dj)), di + dj, rden);
            V[q] = (rel0 + q < w) ? val : V[q];
            left = V[q];
            dj += 1.0;
            if (q == 0) v0_next = __shfl_down_sync(0xffffffffu, V[0], 1);   // lane l + 1's strip starts here
        }
        if (act) {
            if (i == m) {                                         // C(m, n)
                const int reln = n - minj - rel0;
#pragma unroll
                for (int q = 0; q < C; ++q)
                    if (q == reln && rel0 + q < w) *p_out_q = clip01(V[q]);
            }
#pragma unroll
            for (int q = 0; q < C; ++q) U[q] = V[q];
            prev_min = minj;
            prev_w = w;
        }
    }
    __syncwarp();
}

// widest strip the strip form is compiled for (columns of band per lane)
#define KS_STRIP_MAX 8
__device__ __forceinline__ bool strip_dispatch(const double *__restrict__ rtab, int lane, const KsProblem &k,
                                               double *p_out_q) {
    // a row of the band has at most floor(2 h / mg) + 2 cells
    const long long wmax = min(2 * k.h / k.mg + 2, (long long)k.n + 1);
    if (wmax > 32 * KS_STRIP_MAX || k.m + k.n + 2 + 32 * KS_STRIP_MAX >= NCB_KS_RTAB_SIZE) return false;
    const int c = (int)((wmax + 31) / 32);
    switch (c) {
        case 0: case 1: case 2: strip_problem<2>(rtab, lane, k, p_out_q); break;
        case 3: strip_problem<3>(rtab, lane, k, p_out_q); break;
        case 4: strip_problem<4>(rtab, lane, k, p_out_q); break;
        case 5: strip_problem<5>(rtab, lane, k, p_out_q); break;
        case 6: strip_problem<6>(rtab, lane, k, p_out_q); break;
        default: strip_problem<8>(rtab, lane, k, p_out_q); break;
    }
    return true;
}

// ---- CTA per problem: lockstep wavefront over 2 x THREADS rows ------------------------------
// The problems with the most band cells (a high-coverage position with a moderate shift: up to
// 2 x 10^7 cells) would keep one warp busy for tens of milliseconds after everything else is
// done; ks_scan_kernel hands those to the CTA form.  Here 2 x THREADS rows advance together, thread r
// one column behind thread r - 1; the cell above travels through a double-buffered shared-memory
// row (one barrier per step), the row above the THREADS-row chunk is the full row `prev` (no wrap,
// so it never needs resetting: the slots right of the band still hold the 1.0 of the initial
// fill).  Same cell arithmetic as the other kernels, hence the same bits.
template <int THREADS, bool IN_TABLE>
__device__ __forceinline__ void cta_chunk(double *prev, double *xbuf, const double *__restrict__ rtab, int tid,
                                          int rows, int i0, int jlo, int jhi, int ng, int mg, int hh) {
    WavePair r;
    wave_pair_init(r, tid, rows, i0, jlo, jhi, ng, mg, hh);
    const int nsteps = (jhi - jlo) + rows - 1;
    const int k0 = i0 + jlo;
    const double *rp = rtab + k0;
    double rden = IN_TABLE ? __ldg(rp) : wave_recip(rtab, k0);
    const bool first = tid == 0;
    const unsigned base = smem_addr(prev);
    double up0 = first ? lds_f64(base + 8u * (unsigned)r.j) : 1.0;
    const unsigned xw = smem_addr(xbuf + tid), xr = smem_addr(xbuf + (first ? 0 : tid - 1));
    // one step; `off` selects the half of the double buffer (a compile-time constant per call site)
    auto step = [&](const unsigned off, const int t) {
        const double rden_next = IN_TABLE ? __ldg(rp + t + 1) : wave_recip(rtab, k0 + t + 1);
        sts_f64(xw + off, r.b);         // the cell row b finished in the step before
        double up0_next = 1.0;
        if (first) up0_next = lds_f64(base + 8u * (unsigned)min(r.j + 1, NCB_KS_RING_FULL - 1));
        __syncthreads();
        const double up_x = lds_f64(xr + off);
        bool act_a, act_b;
        wave_pair_cells(r, first ? up0 : up_x, rden, hh, act_a, act_b);
        if ((unsigned)r.jr < r.keep_a) sts_f64(base + 8u * (unsigned)r.j, r.a);
        if ((unsigned)(r.jr - 1) < r.keep_b) sts_f64(base + 8u * (unsigned)(r.j - 1), r.b);
        wave_pair_advance(r, mg);
        rden = rden_next;
        up0 = up0_next;
    };
    int t = 0;
#pragma unroll 1
    for (; t + 1 < nsteps; t += 2) {
        step(0u, t);
        step(8u * THREADS, t + 1);
    }
    if (t < nsteps) step(0u, t);
}

template <int THREADS>
__device__ void cta_problem(double *prev, double *xbuf, const double *__restrict__ rtab, int tid,
                            const KsProblem &k, double *p_out_q) {
    const int n = k.n, m = k.m;
    const int mg = (int)k.mg, ng = (int)k.ng, hh = (int)k.h;
    const int maxj0 = min((hh + mg - 1) / mg, n + 1);
    for (int j = tid; j <= n; j += THREADS) prev[j] = (j < maxj0) ? 0.0 : 1.0;
    __syncthreads();
    const bool in_table = m + n + 2 < NCB_KS_RTAB_SIZE;
    for (int i0 = 1; i0 <= m; i0 += 2 * THREADS) {
        const int rows = min(2 * THREADS, m - i0 + 1);
        const int ilast = i0 + rows - 1;
        const int a0 = ng * i0 - hh;            // floor(a0 / mg) + 1, clamped to [0, n]
        const int jlo = min(max((a0 >= 0 ? a0 / mg : -((-a0 + mg - 1) / mg)) + 1, 0), n);
        const int jhi = min((ng * ilast + hh + mg - 1) / mg, n + 1);
        if (in_table) cta_chunk<THREADS, true>(prev, xbuf, rtab, tid, rows, i0, jlo, jhi, ng, mg, hh);
        else cta_chunk<THREADS, false>(prev, xbuf, rtab, tid, rows, i0, jlo, jhi, ng, mg, hh);
        __syncthreads();
    }
    if (tid == 0) *p_out_q = clip01(prev[n]);
    __syncthreads();   // prev is rewritten by the next problem
}

// ---- one launch for both forms ---------------------------------------------------------------
// Every CTA first takes problems from the CTA queue (the HUGE bins, then the dearest LARGE ones,
// a prefix of their cost-sorted list); when that queue is empty its warps turn to the warp queue.
// The tail of the CTA phase on one SM overlaps the warp phase on the others -- as two launches on
// one stream the two forms ran back to back, each with its own tail (24 % of the warp slots in
// use on a high-coverage batch).
#ifndef KS_WAVE_MINB
#define KS_WAVE_MINB 3   /* 74 registers, no spills (64 with 4): -2 % */
#endif
template <int WARPS, int RING>
__global__ void __launch_bounds__(WARPS * 32, KS_WAVE_MINB)
ks_wave_kernel(const int32_t *n0, const int32_t *n1, const int64_t *h, double *p_out, NcbKsWork w) {
    extern __shared__ double wave_smem[];   // CTA form: prev[NCB_KS_RING_FULL], xbuf[2][THREADS]
                                            // warp form: [WARPS][RING]
    constexpr int THREADS = WARPS * 32;
    __shared__ int s_idx;
    const int tid = threadIdx.x;
    {
        double *prev = wave_smem;
        double *xbuf = wave_smem + NCB_KS_RING_FULL;
        const int lo = w.hist[NCB_KS_KIND_HUGE * NCB_KS_BUCKETS];
        const int n_own = w.hist[(NCB_KS_KIND_HUGE + 1) * NCB_KS_BUCKETS] - lo;
        const int lo_large = w.hist[NCB_KS_KIND_LARGE * NCB_KS_BUCKETS];
        const int n_all = n_own + w.fetch[NCB_KS_KINDS];
        for (;;) {
            if (tid == 0) s_idx = atomicAdd(&w.fetch[NCB_KS_KIND_HUGE], 1);
            __syncthreads();
            const int idx = s_idx;
            __syncthreads();               // s_idx is rewritten by the next fetch
            if (idx >= n_all) break;
            const int q = w.list[idx < n_own ? lo + idx : lo_large + (idx - n_own)];
            cta_problem<THREADS>(prev, xbuf, w.rtab, tid, make_problem(n0[q], n1[q], h[q]), p_out + q);
        }
    }
    {
        const int lane = tid & 31;
        double *prev = wave_smem + (size_t)(tid >> 5) * RING;
        const int lo = w.hist[NCB_KS_KIND_LARGE * NCB_KS_BUCKETS] + w.fetch[NCB_KS_KINDS];   // the dearest went to the CTA form
        const int hi = w.hist[(NCB_KS_KIND_LARGE + 1) * NCB_KS_BUCKETS];
        for (;;) {
            int idx = 0;
            if (lane == 0) idx = lo + atomicAdd(&w.fetch[NCB_KS_KIND_LARGE], 1);
            idx = __shfl_sync(0xffffffffu, idx, 0);
            if (idx >= hi) break;
            const int q = w.list[idx];
            const KsProblem k = make_problem(n0[q], n1[q], h[q]);
#if KS_USE_STRIP
            if (strip_dispatch(w.rtab, lane, k, p_out + q)) continue;
#endif
            warp_problem<RING>(prev, w.rtab, lane, k, p_out + q);
        }
    }
}

}  // namespace ncb

using namespace ncb;

static const int KS_WAVE_WARPS = 8;
static constexpr size_t KS_WAVE_SMEM_CTA = (NCB_KS_RING_FULL + 2 * KS_WAVE_WARPS * 32) * sizeof(double);
static constexpr size_t KS_WAVE_SMEM_WARP = (size_t)KS_WAVE_WARPS * NCB_KS_RING_SMALL * sizeof(double);
static constexpr size_t KS_WAVE_SMEM = KS_WAVE_SMEM_CTA > KS_WAVE_SMEM_WARP ? KS_WAVE_SMEM_CTA : KS_WAVE_SMEM_WARP;
static const size_t KS_SMALL_SMEM = (size_t)NCB_KS_SMALL_WINDOW * KS_SMALL_THREADS * sizeof(double);

int ncb_configure_ks_kernels(void) {
    cudaError_t e;
    e = cudaFuncSetAttribute(ks_small_kernel<NCB_KS_SMALL_WINDOW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KS_SMALL_SMEM);
    if (e != cudaSuccess) return -1;
    e = cudaFuncSetAttribute(ks_wave_kernel<KS_WAVE_WARPS, NCB_KS_RING_SMALL>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KS_WAVE_SMEM);
    if (e != cudaSuccess) return -1;
    return 0;
}

int ncb_launch_ks_rtab(double *rtab, cudaStream_t s) {
    ks_rtab_kernel<<<(NCB_KS_RTAB_SIZE + 255) / 256, 256, 0, s>>>(rtab);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int ncb_launch_ks_pvalues(int64_t nprob, const int32_t *n0, const int32_t *n1, const int64_t *h,
                          double *p_out, const NcbKsWork &w, int sm_count, cudaStream_t s) {
    if (nprob <= 0) return 0;
    if (cudaMemsetAsync(w.hist, 0, (NCB_KS_NBINS + 1) * sizeof(int32_t), s) != cudaSuccess) return -1;
    int blocks = (int)((nprob + 255) / 256);
    if (blocks > sm_count * 8) blocks = sm_count * 8;
    ks_classify_kernel<<<blocks, 256, 0, s>>>(nprob, n0, n1, h, p_out, w);
    ks_scan_kernel<<<1, NCB_KS_NBINS, 0, s>>>(w);
    ks_scatter_kernel<<<blocks, 256, 0, s>>>(nprob, w);
    // the long problems first: their warps are busy while the short kernels fill the rest
    ks_wave_kernel<KS_WAVE_WARPS, NCB_KS_RING_SMALL>
        <<<sm_count * 3, KS_WAVE_WARPS * 32, KS_WAVE_SMEM, s>>>(n0, n1, h, p_out, w);
    int tblocks = (int)((nprob + KS_SMALL_THREADS - 1) / KS_SMALL_THREADS);
    ks_small_kernel<NCB_KS_SMALL_WINDOW><<<tblocks, KS_SMALL_THREADS, KS_SMALL_SMEM, s>>>(n0, n1, h, p_out, w);
    ks_small_kernel<32><<<tblocks, KS_SMALL_THREADS, KS_SMALL_SMEM / (NCB_KS_SMALL_WINDOW / 32), s>>>(n0, n1, h, p_out, w);
    ks_equal_kernel<<<tblocks, KS_SMALL_THREADS, 0, s>>>(n0, h, p_out, w);
    return cudaGetLastError() == cudaSuccess ? 7 : -1;
}

// The 2-variable mixture fit of the warp size classes (positions of up to 512 reads) as three
// kernels, SUB lanes per position (G = 32 / SUB positions per warp):
//
//   gmm_init_kernel   K = 1 closed form (BIC1), farthest-pair + Lloyd 2-means, the sufficient
//                     statistics of the 2-means assignment and the first M-step
//   gmm_em_kernel     the EM passes (E-step + M-step) until the stop rule fires
//   gmm_final_kernel  the final E-step: log-likelihood (BIC2), hard assignment, contingency
//                     table, per-sample counts, BIC gate, LOR, chi-squared, logit
//
// Replaces comparisons.py:338-392, 477-605 + gmm_gpu.GMM of the reference (definition:
// oracle/gmm.py), like position_gmm in position_kernel.cu, which keeps the CTA size classes.
//
// Why.  A position of config 3 has ~170 reads in its fit: with one warp per position a pass over
// the reads is 6 trips of the per-read code, and what follows every pass -- the reduction of the
// sufficient statistics, the M-step (a few hundred dependent fp64 operations, the same on every
// lane), the stop test -- costs as many instructions as the pass itself.  With 8 lanes per
// position that fixed part is executed once for 4 positions, and the last trip of a pass is
// 95 % instead of 87 % full.  The positions of a warp do not run in lockstep through the fit
// (the number of EM passes varies 20-fold), so the code is arranged to be the SAME for every
// group whatever the state of its position: the EM kernel is one loop "E-step pass, reduce,
// stop test, M-step" that every group executes; a group whose position has converged stores the
// parameters and takes the next position from the work counter (the warp loads it cooperatively)
// without leaving the loop.  The parts whose code differs per stage (initialisation, final pass
// and tail) are kernels of their own; the initialisation runs its groups in lockstep (a group
// whose Lloyd iterations have converged idles until the slowest of its warp has).  Three small
// kernels also keep the instruction working set of an SM small (position_kernel.cu, "Why two
// kernels").
//
// Hand-over between the three: a record of NCB_GMM_REC doubles per position (handle scratch).
// Values live in shared memory as the float32 numbers they are (8 B/read) and are widened when
// used.
#include "gmm_common.cuh"

#ifndef NCB_GS_U_INIT
#define NCB_GS_U_INIT 4
#endif
#ifndef NCB_GS_INTSEL
#define NCB_GS_INTSEL 1    /* EM kernel: comparisons of doubles on the integer pipe (e_step_one_i) */
#endif

namespace ncb {

// record of a position: [0, 12) the parameters of both components as the E-step uses them
// (CompE x 2), [12, 18) totals of the position (n, sum x, y, xx, xy, yy), 18 BIC of the K = 1 fit,
// 19 which component the next pass sums (0 / 1), 20 EM passes done
enum { REC_CE = 0, REC_TOT = 12, REC_BIC1 = 18, REC_SECOND = 19, REC_ITERS = 20 };

// The positions a launch works on: the worklists of one or two size classes, taken as one sequence
// (one launch for the classes of up to 128 and up to 256 reads: one tail instead of two).
struct WorkLists {
    const int32_t *list_a, *count_a, *list_b, *count_b;   // b may be absent (NULL)
    __device__ __forceinline__ int total() const { return *count_a + (count_b ? *count_b : 0); }
    __device__ __forceinline__ int32_t at(int i) const {
        const int na = *count_a;
        return i < na ? list_a[i] : list_b[i - na];
    }
};

template <int SUB, int NV>
__device__ __forceinline__ void sub_reduce_add(double (&v)[NV]) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
        for (int m = SUB / 2; m > 0; m >>= 1) v[i] += __shfl_xor_sync(NCB_FULL, v[i], m);
    }
}
template <int SUB>
__device__ __forceinline__ u64 sub_reduce_add_u64(u64 v) {
#pragma unroll
    for (int m = SUB / 2; m > 0; m >>= 1) v += __shfl_xor_sync(NCB_FULL, v, m);
    return v;
}

// The reads of position p that take part in the fit (x is a number), in read order, compacted
// into val[]; all 32 lanes of the warp load.  Returns their number.
__device__ __forceinline__ int load_fit_reads(const NcbDeviceArgs &a, int64_t p, float2 *val, int lane) {
    const float2 *zin;
    int n_in;
    if (a.pos_offsets) {
        const int64_t o = a.pos_offsets[p];
        n_in = (int)(a.pos_offsets[p + 1] - o);
        zin = a.zbuf + o;
    } else {
        n_in = a.n_reads;
        zin = a.zbuf + p * (int64_t)n_in;
    }
    // four loads per lane are in flight before the first is looked at: one round trip to HBM per
    // 128 entries instead of one per 32
    int n = 0;
#pragma unroll 1
    for (int base = 0; base < n_in; base += 128) {
        float2 z[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = base + j * 32 + lane;
            z[j] = make_float2(__int_as_float(0x7fc00000), 0.f);
            if (e < n_in) z[j] = zin[e];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool ok = z[j].x == z[j].x;
            const unsigned m = __ballot_sync(NCB_FULL, ok);
            if (ok) val[n + __popc(m & ((1u << lane) - 1u))] = z[j];
            n += __popc(m);
        }
    }
    return n;
}

__device__ __forceinline__ void load_ce(const double *r, CompE (&ce)[2]) {
    ce[0].mx = r[0]; ce[0].my = r[1]; ce[0].ha = r[2]; ce[0].hb = r[3]; ce[0].hc = r[4]; ce[0].cst = r[5];
    ce[1].mx = r[6]; ce[1].my = r[7]; ce[1].ha = r[8]; ce[1].hb = r[9]; ce[1].hc = r[10]; ce[1].cst = r[11];
}
__device__ __forceinline__ void store_ce(double *r, const CompE (&ce)[2]) {
    r[0] = ce[0].mx; r[1] = ce[0].my; r[2] = ce[0].ha; r[3] = ce[0].hb; r[4] = ce[0].hc; r[5] = ce[0].cst;
    r[6] = ce[1].mx; r[7] = ce[1].my; r[8] = ce[1].ha; r[9] = ce[1].hb; r[10] = ce[1].hc; r[11] = ce[1].cst;
}

// sufficient statistics of both components from those of the one that was summed (r) and the
// totals of the position (position_kernel.cu, NCB_GMM_DERIVE)
__device__ __forceinline__ void derive12(const double (&r)[7], const double *tot, bool second, double (&s)[12]) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double rest = tot[i] - r[i];
        if (i == 0) rest = fmax(rest, 0.0);
        s[i] = second ? rest : r[i];
        s[6 + i] = second ? r[i] : rest;
    }
}

// ---- initialisation ------------------------------------------------------------------------
template <int SUB, int CAP>
struct InitSmem {
    // + 8: the groups' rows start 64 bytes apart modulo the 128 bytes of the banks, so that the four
    // 64-byte reads of a step fall on two wavefronts (the least possible), not four
    float2 val[32 / SUB][CAP + 8];
    uint8_t k0[32 / SUB][CAP + 8];     // 2-means assignment of a read: 1 = cluster 0
};

template <int SUB, int CAP, int WPC, int MINB>
__global__ void __launch_bounds__(WPC * 32, MINB)
gmm_init_kernel(NcbDeviceArgs a, WorkLists wl, int32_t *fetch, double *rec) {
    constexpr int G = 32 / SUB;
    constexpr int UI = NCB_GS_U_INIT;   // trips of a pass over the reads unrolled together
    extern __shared__ __align__(16) unsigned char smem[];
    typedef InitSmem<SUB, CAP> Smem;
    Smem *sm = reinterpret_cast<Smem *>(smem) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int sub = lane & (SUB - 1), grp = lane / SUB;
    const float2 *val = sm->val[grp];
    uint8_t *k0 = sm->k0[grp];
    const int total = wl.total();
    const double reg = a.reg_covar;
    for (;;) {
        int i0 = 0;
        if (lane == 0) i0 = atomicAdd(fetch, G);
        i0 = __shfl_sync(NCB_FULL, i0, 0);
        if (i0 >= total) break;
        int64_t p = -1;
        int n = 0;
#pragma unroll 1
        for (int gi = 0; gi < G; ++gi) {
            const int idx = i0 + gi;
            const int64_t q = idx < total ? (int64_t)wl.at(idx) : -1;
            int nq = 0;
            if (q >= 0 && (unsigned)a.gmm_n[q] >= 2u) nq = load_fit_reads(a, q, sm->val[gi], lane);
            if (grp == gi) { p = q; n = nq; }
        }
        __syncwarp();
        const bool valid = n >= 2;      // fewer: the fit is undefined, BIC and p-value stay NaN, counts 0
        if (!valid) n = 0;
        const double dn = (double)n;

        // K = 1: closed form
        double T[5] = {0, 0, 0, 0, 0};   // sum x, y, xx, xy, yy
#pragma unroll UI
        for (int e = sub; e < n; e += SUB) {
            const float2 zf = val[e];
            const double x = (double)zf.x, y = (double)zf.y;
            T[0] += x; T[1] += y;
            T[2] = fma(x, x, T[2]); T[3] = fma(x, y, T[3]); T[4] = fma(y, y, T[4]);
        }
        __syncwarp();
        sub_reduce_add<SUB, 5>(T);
        double bic1;
        {
            double s[6] = {dn, T[0], T[1], T[2], T[3], T[4]};
            Comp c1;
            const double nk = m_step(s, reg, c1);
            finish_comp(c1, 1.0);
            // sum of Mahalanobis distances = ia*Axx + 2 ib*Axy + ic*Ayy with A = nk*(cov - reg)
            const double axx = (c1.cxx - reg) * nk, ayy = (c1.cyy - reg) * nk, axy = c1.cxy * nk;
            const double maha = c1.ia * axx + 2.0 * c1.ib * axy + c1.ic * ayy;
            const double L1 = c1.cst - 0.5 * maha / dn;
            bic1 = -2.0 * dn * L1 + 5.0 * log(dn);
        }

        // K = 2: deterministic 2-means initialisation (oracle/gmm.py).  Centre 0 = the read farthest
        // from the mean, centre 1 = the read farthest from centre 0 (ties -> lowest entry); then
        // Lloyd iterations until no assignment changes (at most 20).
        double cen[4];   // c0x, c0y, c1x, c1y
        {
            double refx = T[0] / dn, refy = T[1] / dn;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                double best = -1.0;
                int beste = 0x7fffffff;
#pragma unroll UI
                for (int e = sub; e < n; e += SUB) {
                    const float2 zf = val[e];
                    const double dx = (double)zf.x - refx, dy = (double)zf.y - refy;
                    const double d = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                    if (d > best) { best = d; beste = e; }
                }
                __syncwarp();
                double gbest = best;
#pragma unroll
                for (int m = SUB / 2; m > 0; m >>= 1) gbest = fmax(gbest, __shfl_xor_sync(NCB_FULL, gbest, m));
                int ge = best == gbest ? beste : 0x7fffffff;
#pragma unroll
                for (int m = SUB / 2; m > 0; m >>= 1) ge = min(ge, __shfl_xor_sync(NCB_FULL, ge, m));
                if (!valid) ge = 0;
                const float2 zf = val[ge];
                refx = (double)zf.x; refy = (double)zf.y;
                cen[pass * 2] = refx; cen[pass * 2 + 1] = refy;
            }
        }
        // Cluster 0 (seeded by the read farthest from the mean: as a rule the smaller one) is summed,
        // cluster 1 is what is left of the totals T.  The groups of a warp iterate in lockstep; one
        // that has converged waits for the others.
        bool done = !valid;
#pragma unroll 1
        for (int it = 1; it <= 20; ++it) {
            if (__all_sync(NCB_FULL, done)) break;
            double k[4] = {0, 0, 0, 0};   // n0, sx0, sy0, changed
            if (!done) {
#pragma unroll UI
                for (int e = sub; e < n; e += SUB) {
                    const float2 zf = val[e];
                    const double x = (double)zf.x, y = (double)zf.y;
                    const double ax = x - cen[0], ay = y - cen[1], bx = x - cen[2], by = y - cen[3];
                    const double d0 = __dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay));
                    const double d1 = __dadd_rn(__dmul_rn(bx, bx), __dmul_rn(by, by));
                    const bool a0 = d0 <= d1;                        // tie -> cluster 0
                    const double w0 = a0 ? 1.0 : 0.0;
                    k[0] += w0; k[1] = fma(w0, x, k[1]); k[2] = fma(w0, y, k[2]);
                    const bool changed = it == 1 || a0 != (k0[e] != 0);
                    k[3] += changed ? 1.0 : 0.0;
                    k0[e] = a0 ? 1 : 0;
                }
            }
            __syncwarp();
            sub_reduce_add<SUB, 4>(k);
            if (!done) {
                if (it > 1 && k[3] == 0.0) {
                    done = true;
                } else {
                    if (k[0] > 0.0) {   // an emptied cluster keeps its centre
                        const double r = __drcp_rn(k[0]);
                        cen[0] = quot_rn(k[1], k[0], r); cen[1] = quot_rn(k[2], k[0], r);
                    }
                    const double n1 = dn - k[0];
                    if (n1 > 0.0) {
                        const double r = __drcp_rn(n1);
                        cen[2] = quot_rn(T[0] - k[1], n1, r); cen[3] = quot_rn(T[1] - k[2], n1, r);
                    }
                }
            }
        }
        __syncwarp();

        // sufficient statistics of the 2-means assignment (cluster 0 summed) and the first M-step
        double r7[7] = {0, 0, 0, 0, 0, 0, 0};
        {
            double acc[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll UI
            for (int e = sub; e < n; e += SUB) {
                const float2 zf = val[e];
                accumulate6(acc, k0[e] ? 1.0 : 0.0, (double)zf.x, (double)zf.y);
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) r7[i] = acc[i];
        }
        __syncwarp();
        sub_reduce_add<SUB, 7>(r7);
        const double tot[6] = {dn, T[0], T[1], T[2], T[3], T[4]};
        double s[12];
        derive12(r7, tot, false, s);
        CompE ce[2];
        m_step_pair(s, reg, ce, lane);
        if (p >= 0 && sub == 0) {
            double *r = rec + p * NCB_GMM_REC;
            if (valid) {
                store_ce(r + REC_CE, ce);
#pragma unroll
                for (int i = 0; i < 6; ++i) r[REC_TOT + i] = tot[i];
                r[REC_BIC1] = bic1;
                r[REC_SECOND] = s[6] < s[0] ? 1.0 : 0.0;
                r[REC_ITERS] = 0.0;
            } else {
                r[REC_TOT] = 0.0;    // no fit on this position: the other two kernels pass it by
            }
        }
        __syncwarp();   // the groups' shared memory is reused by the next positions
    }
}

// ---- EM passes --------------------------------------------------------------------------------
template <int SUB, int CAP>
struct EmSmem {
    float2 val[32 / SUB][CAP + 8];     // + 8: see InitSmem
    double tot[32 / SUB][6];
};

// float32 -> float64 of a number in [0, 1] on the integer pipe (the conversion instruction shares the
// quarter-rate pipe of ex2 / rcp, which bounds the float32 E-step); denormals count as 0
__device__ __forceinline__ double widen01(float r) {
    const uint32_t b = __float_as_uint(r);
    const uint32_t hi = (b >> 3) + 0x38000000u, lo = b << 29;
    const bool tiny = b < 0x00800000u;
    return __hiloint2double(tiny ? 0 : (int)hi, tiny ? 0 : (int)lo);
}

template <int SUB, int CAP, int WPC, int MINB, bool F32, int U>
__global__ void __launch_bounds__(WPC * 32, MINB)
gmm_em_kernel(NcbDeviceArgs a, WorkLists wl, int32_t *fetch, double *rec) {
    constexpr int G = 32 / SUB;
    extern __shared__ __align__(16) unsigned char smem[];
    typedef EmSmem<SUB, CAP> Smem;
    Smem *sm = reinterpret_cast<Smem *>(smem) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int sub = lane & (SUB - 1), grp = lane / SUB;
    const float2 *val = sm->val[grp];
    const double *tot = sm->tot[grp];
    const int total = wl.total();
    const int64_t P = a.n_positions;
    const double reg = a.reg_covar;
    // state of this group's position
    bool have = false, second = false;
    int64_t p = -1;
    int n = 0, it = 0;
    double prev = 0.0, dn = 1.0, rdn = 1.0;   // rdn = 1 / dn, correctly rounded
    CompE ce[2];
    ce[0].mx = ce[0].my = ce[0].ha = ce[0].hb = ce[0].hc = ce[0].cst = 0.0;   // a group without a position
    ce[1] = ce[0];                                                            // computes and discards
    bool exhausted = false;
    for (;;) {
        // ---- groups without a position take the next one from the work counter -----------------
        const unsigned need = __ballot_sync(NCB_FULL, !have);
        if (need && !exhausted) {
#pragma unroll 1
            for (int gi = 0; gi < G; ++gi) {
                if (!((need >> (gi * SUB)) & 1u)) continue;
                int64_t q = -1;
                while (!exhausted) {
                    int idx = 0;
                    if (lane == 0) idx = atomicAdd(fetch, 1);
                    idx = __shfl_sync(NCB_FULL, idx, 0);
                    if (idx >= total) { exhausted = true; break; }
                    const int64_t c = (int64_t)wl.at(idx);
                    if ((unsigned)a.gmm_n[c] >= 2u && rec[c * NCB_GMM_REC + REC_TOT] >= 2.0) { q = c; break; }
                }
                if (q < 0) break;
                const int nq = load_fit_reads(a, q, sm->val[gi], lane);
                if (grp == gi) {
                    const double *r = rec + q * NCB_GMM_REC;
                    p = q; n = nq; have = true; it = 0;
                    prev = __longlong_as_double(0xfff0000000000000LL);
                    load_ce(r + REC_CE, ce);
                    second = r[REC_SECOND] != 0.0;
                    dn = r[REC_TOT];
                    rdn = __drcp_rn(dn);
                    for (int i = sub; i < 6; i += SUB) sm->tot[gi][i] = r[REC_TOT + i];
                }
            }
            __syncwarp();
        }
        if (!__any_sync(NCB_FULL, have)) break;

        // ---- E-step pass: the sufficient statistics of ONE component (the smaller one of the pass
        // before) and the log-likelihood; max(l0, l1) + log(1 + e^-|l0 - l1|) per read, the logarithms
        // of a lane's reads taken together as the log of the product (each factor is in (1, 2])
        double r7[7];
        {
            double acc[6] = {0, 0, 0, 0, 0, 0};
            double ll = 0.0;
            const int nn = have ? n : 0;
            // U reads of a lane go through the chain at a time (independent instructions to fill the
            // latency of the dependent ones); the last trip may hold fewer: the missing ones are
            // computed on the lane's first read of the trip and enter every sum with weight zero
            if (F32) {
                const CompEf c0 = narrow(ce[0]), c1 = narrow(ce[1]);
                float lmax_sum = 0.f, lprod = 1.f;
#pragma unroll 1
                for (int e = sub; e < nn; e += U * SUB) {
                    float2 zf[U];
                    bool ok[U];
                    RespF r[U];
#pragma unroll
                    for (int j = 0; j < U; ++j) {
                        ok[j] = j == 0 || e + j * SUB < nn;
                        zf[j] = val[ok[j] ? e + j * SUB : e];
                    }
#pragma unroll
                    for (int j = 0; j < U; ++j) r[j] = e_step_f(c0, c1, zf[j].x, zf[j].y);
#pragma unroll
                    for (int j = 0; j < U; ++j) {
                        lmax_sum += ok[j] ? r[j].lmax : 0.f;
                        lprod *= ok[j] ? r[j].sum : 1.f;
                        const float w = second ? r[j].r1 : r[j].r0;
                        accumulate6(acc, widen01(ok[j] ? w : 0.f), (double)zf[j].x, (double)zf[j].y);
                    }
                }
                ll = (double)lmax_sum + (double)__logf(lprod);
            } else {
                const CompE c0 = ce[0], c1 = ce[1];
                double lprod = 1.0;
#pragma unroll 1
                for (int e = sub; e < nn; e += U * SUB) {
                    double x[U], y[U];
                    bool ok[U];
                    RespOne r[U];
#pragma unroll
                    for (int j = 0; j < U; ++j) {
                        ok[j] = j == 0 || e + j * SUB < nn;
                        const float2 zf = val[ok[j] ? e + j * SUB : e];
                        x[j] = (double)zf.x; y[j] = (double)zf.y;
                    }
#pragma unroll
                    for (int j = 0; j < U; ++j) {
#if NCB_GS_INTSEL
                        r[j] = e_step_one_i(c0, c1, x[j], y[j], second ? 0x80000000u : 0u);
#else
                        r[j] = e_step_one(c0, c1, x[j], y[j], second);
#endif
                    }
#pragma unroll
                    for (int j = 0; j < U; ++j) {
                        ll += ok[j] ? r[j].lmax : 0.0;
                        lprod *= ok[j] ? r[j].sum : 1.0;
                        accumulate6(acc, ok[j] ? r[j].r : 0.0, x[j], y[j]);
                    }
                }
                ll += log(lprod);
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) r7[i] = acc[i];
            r7[6] = ll;
        }
        __syncwarp();
        sub_reduce_add<SUB, 7>(r7);
        double s[12];
        derive12(r7, tot, second, s);
        ++it;
        const double Lm = r7[6] * rdn;    // mean log-likelihood (a division: 15 instructions on every lane)
        const bool last = fabs(Lm - prev) < a.em_tol || it >= a.em_max_iter;
        prev = Lm;
        if (have && last && a.diag_f64 && sub == 0) {
            // the parameters of the fit, from the sufficient statistics of the last pass
            Comp c0, c1;
            const double nk0 = m_step(s, reg, c0), nk1 = m_step(s + 6, reg, c1);
            double *d = a.diag_f64;
            d[NCB_DF_W0 * P + p] = nk0 / (nk0 + nk1);
            d[NCB_DF_MU0_X * P + p] = c0.mx; d[NCB_DF_MU0_Y * P + p] = c0.my;
            d[NCB_DF_C0_XX * P + p] = c0.cxx; d[NCB_DF_C0_XY * P + p] = c0.cxy;
            d[NCB_DF_C0_YY * P + p] = c0.cyy;
            d[NCB_DF_MU1_X * P + p] = c1.mx; d[NCB_DF_MU1_Y * P + p] = c1.my;
            d[NCB_DF_C1_XX * P + p] = c1.cxx; d[NCB_DF_C1_XY * P + p] = c1.cxy;
            d[NCB_DF_C1_YY * P + p] = c1.cyy;
        }
        __syncwarp();
        m_step_pair(s, reg, ce, lane);
        second = s[6] < s[0];
        if (have && last) {
            if (sub == 0) {
                double *r = rec + p * NCB_GMM_REC;
                store_ce(r + REC_CE, ce);
                r[REC_ITERS] = (double)it;
            }
            have = false;
        }
    }
}

// ---- final E-step and what follows the fit -----------------------------------------------------
template <int SUB, int WPC, int MINB, bool F32>
__global__ void __launch_bounds__(WPC * 32, MINB)
gmm_final_kernel(NcbDeviceArgs a, WorkLists wl, int32_t *fetch, const double *rec) {
    constexpr int G = 32 / SUB;
    extern __shared__ __align__(16) unsigned char smem[];
    PosShared *ps = reinterpret_cast<PosShared *>(smem) + (threadIdx.x >> 5) * G + (threadIdx.x & 31) / SUB;
    const int lane = threadIdx.x & 31;
    const int sub = lane & (SUB - 1), grp = lane / SUB;
    const int total = wl.total();
    const int64_t P = a.n_positions;
    const bool soft = a.cluster_counts == NCB_COUNTS_SOFT;
    const int S2 = a.n_samples * 2;
    for (;;) {
        int i0 = 0;
        if (lane == 0) i0 = atomicAdd(fetch, G);
        i0 = __shfl_sync(NCB_FULL, i0, 0);
        if (i0 >= total) break;
        const int idx = i0 + grp;
        int64_t p = idx < total ? (int64_t)wl.at(idx) : -1;
        const double *r = rec + (p < 0 ? 0 : p) * NCB_GMM_REC;
        if (p >= 0 && !((unsigned)a.gmm_n[p] >= 2u && r[REC_TOT] >= 2.0)) p = -1;
        for (int i = sub; i < S2; i += SUB) { ps->scnt[i] = 0u; ps->sacc[i] = 0.0; }
        __syncwarp();
        int n_in = 0;
        const uint8_t *lab = a.label;
        const float2 *zin = a.zbuf;
        CompE ce[2];
        load_ce(r + REC_CE, ce);
        if (p >= 0) {
            if (a.pos_offsets) {
                const int64_t o = a.pos_offsets[p];
                n_in = (int)(a.pos_offsets[p + 1] - o);
                lab = a.label + o;
                zin = a.zbuf + o;
            } else {
                n_in = a.n_reads;
                zin = a.zbuf + p * (int64_t)n_in;
            }
        }
        u64 cont = 0;                  // hard contingency, field cond*2 + cluster
        double sc[4] = {0, 0, 0, 0};   // soft contingency
        double ll = 0.0;
        {
            const CompE c0 = ce[0], c1 = ce[1];
            const CompEf c0f = narrow(ce[0]), c1f = narrow(ce[1]);
            double lprod = 1.0;
            // four reads per lane are loaded before the first is used (one round trip to HBM per 4 SUB reads)
#pragma unroll 1
            for (int base = 0; base < n_in; base += 4 * SUB) {
                float2 z4[4];
                uint32_t f4[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int e = base + j * SUB + sub;
                    z4[j] = make_float2(__int_as_float(0x7fc00000), 0.f);
                    f4[j] = 0u;
                    if (e < n_in) { z4[j] = zin[e]; f4[j] = (uint32_t)__ldg(lab + e) & 0x7fu; }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                const float2 zf = z4[j];
                if (!(zf.x == zf.x)) continue;
                const uint32_t f = f4[j];
                Resp rr;
                if (F32) {
                    const RespF rf = e_step_f(c0f, c1f, zf.x, zf.y);
                    rr.r0 = (double)rf.r0; rr.r1 = (double)rf.r1; rr.lmax = (double)rf.lmax;
                    rr.sum = (double)rf.sum; rr.d = (double)rf.d;
                } else {
                    rr = e_step(c0, c1, (double)zf.x, (double)zf.y);
                }
                ll += rr.lmax;
                lprod *= rr.sum;
                const int pred = rr.d < 0.0 ? 1 : 0;      // l1 > l0
                const int c = f & F_COND;
                const int smp = (int)(f >> 1);
                cont += 1ull << (16 * (c * 2 + pred));
                if (smp < a.n_samples) atomicAdd(&ps->scnt[smp * 2 + pred], 1u);
                if (soft) {
                    // the reference sums float32 posteriors
                    const double q0 = (double)(float)rr.r0, q1 = (double)(float)rr.r1;
                    if (c) { sc[2] += q0; sc[3] += q1; } else { sc[0] += q0; sc[1] += q1; }
                    if (smp < a.n_samples) {
                        atomicAdd(&ps->sacc[smp * 2], q0);
                        atomicAdd(&ps->sacc[smp * 2 + 1], q1);
                    }
                }
            }
            }
            ll += log(lprod);
        }
        __syncwarp();
        {
            double v[1] = {ll};
            sub_reduce_add<SUB, 1>(v);
            ll = v[0];
        }
        cont = sub_reduce_add_u64<SUB>(cont);
        if (soft) sub_reduce_add<SUB, 4>(sc);
        __syncwarp();
        if (p >= 0 && sub == 0) {
            const double bic1 = r[REC_BIC1];
            const double bic2 = gmm_tail(a, p, ps, cont, sc, soft, ll, r[REC_TOT], bic1, 11.0, (int)r[REC_ITERS]);
            if (a.diag_f64) {
                a.diag_f64[NCB_DF_BIC1 * P + p] = bic1;
                a.diag_f64[NCB_DF_BIC2 * P + p] = bic2;
            }
        }
        __syncwarp();   // the group's counters are reset for its next position
    }
}

}  // namespace ncb

using namespace ncb;

// lanes per position / warps per CTA / CTAs per SM of the three warp size classes
#ifndef NCB_GS_SUB1
#define NCB_GS_SUB1 8
#endif
#ifndef NCB_GS_SUB2
#define NCB_GS_SUB2 16
#endif
// warps per CTA and CTAs per SM of the three kernels; reads a lane carries through the E-step at a time
#ifndef NCB_GS_WPC
#define NCB_GS_WPC 8
#endif
#ifndef NCB_GS_MINB_INIT
#define NCB_GS_MINB_INIT 3
#endif
#ifndef NCB_GS_MINB_EM
#define NCB_GS_MINB_EM 2      /* 117 registers at U = 3 */
#endif
#ifndef NCB_GS_MINB_FINAL
#define NCB_GS_MINB_FINAL 3
#endif
#ifndef NCB_GS_SUB_INIT
#define NCB_GS_SUB_INIT 0     /* lanes per position of the initialisation kernel; 0: as the other two */
#endif
#ifndef NCB_GS_U_EM
#define NCB_GS_U_EM 3         /* measured per 100 k positions of config 3: 1.244 ms (U = 1, 3 CTAs) / 1.209 (2, 2) / 1.207 (3, 2) */
#endif

template <int SUB, int CAP, bool F32>
static int split_class(const NcbDeviceArgs &a, const WorkLists &wl, int32_t *f_init, int32_t *f_em, int32_t *f_final,
                       int sm_count, cudaStream_t s, bool configure, cudaEvent_t *marks) {
    constexpr int WPC = NCB_GS_WPC, G = 32 / SUB;
    constexpr int MB_I = NCB_GS_MINB_INIT, MB_E = NCB_GS_MINB_EM, MB_F = NCB_GS_MINB_FINAL;
    // the initialisation may use another number of lanes per position (the hand-over is per position)
    constexpr int SUB_I = NCB_GS_SUB_INIT > 0 ? (NCB_GS_SUB_INIT > SUB ? NCB_GS_SUB_INIT : SUB) : SUB, G_I = 32 / SUB_I;
    constexpr size_t sm_init = WPC * sizeof(InitSmem<SUB_I, CAP>), sm_em = WPC * sizeof(EmSmem<SUB, CAP>),
                     sm_final = WPC * G * sizeof(PosShared);
    static_assert(sm_init * MB_I + 1024 * MB_I <= 233472 && sm_em * MB_E + 1024 * MB_E <= 233472,
                  "split mixture kernels: the CTAs of an SM do not fit its shared memory");
    auto k_init = gmm_init_kernel<SUB_I, CAP, WPC, MB_I>;
    auto k_em = gmm_em_kernel<SUB, CAP, WPC, MB_E, F32, NCB_GS_U_EM>;
    auto k_final = gmm_final_kernel<SUB, WPC, MB_F, F32>;
    if (configure) {
        if (cudaFuncSetAttribute(k_init, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_init) != cudaSuccess) return -1;
        if (cudaFuncSetAttribute(k_em, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_em) != cudaSuccess) return -1;
        if (cudaFuncSetAttribute(k_final, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_final) != cudaSuccess) return -1;
        return 0;
    }
    const int64_t P = a.n_positions;
    auto grid = [&](int minb, int groups) {
        const int64_t need = (P + WPC * groups - 1) / (WPC * groups);
        const int64_t cap = (int64_t)sm_count * minb;
        return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
    };
    k_init<<<grid(MB_I, G_I), WPC * 32, sm_init, s>>>(a, wl, f_init, a.gmm_rec);
    if (marks) cudaEventRecord(marks[0], s);
    k_em<<<grid(MB_E, G), WPC * 32, sm_em, s>>>(a, wl, f_em, a.gmm_rec);
    if (marks) cudaEventRecord(marks[1], s);
    k_final<<<grid(MB_F, G), WPC * 32, sm_final, s>>>(a, wl, f_final, a.gmm_rec);
    return cudaGetLastError() == cudaSuccess ? 3 : -1;
}

// `wide`: the class of up to 512 reads (16 lanes per position); else the classes of up to 128 and up to
// 256 reads as one launch (8 lanes per position).  The work counters are those of the larger class.
template <bool F32>
static int split_all(const NcbDeviceArgs &a, const NcbWorklists &w, bool wide, int sm_count, cudaStream_t s, bool configure,
                     cudaEvent_t *marks = nullptr) {
    const int64_t P = a.n_positions;
    const int cls = wide ? 2 : 1;
    WorkLists wl{};
    int32_t *f_init = nullptr, *f_em = nullptr, *f_final = nullptr;
    if (!configure) {
        wl.list_a = w.class_list + (int64_t)cls * P;
        wl.count_a = w.class_count + cls;
        if (!wide) {
            wl.list_b = w.class_list;
            wl.count_b = w.class_count;
        }
        f_init = w.class_count + NCB_NUM_CLASSES * (1 + NCB_PHASE_GMM) + cls;
        f_em = w.class_count + NCB_NUM_CLASSES * (1 + NCB_PHASE_GMM_EM) + cls;
        f_final = w.class_count + NCB_NUM_CLASSES * (1 + NCB_PHASE_GMM_FINAL) + cls;
    }
    return wide ? split_class<NCB_GS_SUB2, 512, F32>(a, wl, f_init, f_em, f_final, sm_count, s, configure, marks)
                : split_class<NCB_GS_SUB1, 256, F32>(a, wl, f_init, f_em, f_final, sm_count, s, configure, marks);
}

int ncb_configure_gmm_split(void) {
    NcbDeviceArgs a{};
    NcbWorklists w{};
    for (int wide = 0; wide < 2; ++wide)
        if (split_all<false>(a, w, wide != 0, 0, 0, true) != 0 || split_all<true>(a, w, wide != 0, 0, 0, true) != 0) return -1;
    return 0;
}

int ncb_launch_gmm_split(const NcbDeviceArgs &a, const NcbWorklists &w, bool wide, int sm_count, cudaStream_t s,
                         cudaEvent_t *marks) {
    return a.em_fp32 ? split_all<true>(a, w, wide, sm_count, s, false, marks)
                     : split_all<false>(a, w, wide, sm_count, s, false, marks);
}

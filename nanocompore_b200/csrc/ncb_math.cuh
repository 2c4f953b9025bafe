// Scalar fp64 special functions and closed-form p-values used once per position.
// All functions are __host__ __device__ so that tests/csrc_host_check.cpp can run them
// on the CPU against SciPy without a GPU (they are not a CPU product path).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define NCB_HD __host__ __device__
#else
#define NCB_HD
#endif

namespace ncb {

NCB_HD inline double ncb_nan() { return nan(""); }

// a / b from rb = RN(1 / b): q = RN(a rb), r = a - b q (exact in an fma), RN(q + r rb).  With a
// correctly rounded reciprocal this is the correctly rounded quotient (Markstein), i.e. the bits
// of the division, for 3 operations; several quotients by one divisor share the reciprocal.
// Operands must be normal numbers whose quotient is normal (no range handling).
NCB_HD inline double quot_rn(double a, double b, double rb) {
#if defined(__CUDA_ARCH__)
    const double q = __dmul_rn(a, rb);
#else
    const double q = a * rb;
#endif
    return fma(fma(-b, q, a), rb, q);
}

NCB_HD inline double clip01(double p) {
    // np.clip keeps NaN
    if (p != p) return p;
    return p < 0.0 ? 0.0 : (p > 1.0 ? 1.0 : p);
}

// two-sided normal p-value 2*Phi(-|z|) = erfc(|z|/sqrt 2)
NCB_HD inline double normal_two_sided(double z) {
    if (z != z) return z;
    return clip01(erfc(fabs(z) * 0.70710678118654752440));
}

// chi2.sf(x, 1) = erfc(sqrt(x/2))
NCB_HD inline double chi2_sf_1dof(double x) {
    if (x != x) return x;
    if (x <= 0.0) return 1.0;
    return erfc(sqrt(0.5 * x));
}

// Continued fraction of the incomplete beta function (modified Lentz).
NCB_HD inline double betacf(double a, double b, double x) {
    const double TINY = 1e-300;
    const double EPS = 1e-16;
    double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0;
    double d = 1.0 - qab * x / qap;
    if (fabs(d) < TINY) d = TINY;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 1000; ++m) {
        double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < TINY) d = TINY;
        c = 1.0 + aa / c;
        if (fabs(c) < TINY) c = TINY;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < TINY) d = TINY;
        c = 1.0 + aa / c;
        if (fabs(c) < TINY) c = TINY;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < EPS) break;
    }
    return h;
}

// regularised incomplete beta I_x(a, b); xc = 1 - x is passed separately so that neither
// end of the interval loses digits
NCB_HD inline double incbet(double a, double b, double x, double xc) {
    if (!(x > 0.0)) return (x == 0.0) ? 0.0 : ncb_nan();
    if (!(xc > 0.0)) return (xc == 0.0) ? 1.0 : ncb_nan();
    double lbeta = lgamma(a + b) - lgamma(a) - lgamma(b);
    double front = exp(lbeta + a * log(x) + b * log(xc));
    if (x < (a + 1.0) / (a + b + 2.0)) return front * betacf(a, b, x) / a;
    return 1.0 - front * betacf(b, a, xc) / b;
}

// two-sided Student-t p-value 2*sf(|t|, df) = I_{df/(df+t^2)}(df/2, 1/2)
NCB_HD inline double student_t_two_sided(double t, double df) {
    if (t != t || df != df || !(df > 0.0)) return ncb_nan();
    if (isinf(t)) return 0.0;
    double t2 = t * t;
    double den = df + t2;
    return clip01(incbet(0.5 * df, 0.5, df / den, t2 / den));
}

// Regularised upper incomplete gamma function Q(a, x) = chi2.sf(2x, 2a) for a > 0: the power
// series of P for x < a + 1, the continued fraction of Q (modified Lentz) otherwise.  Used by
// the sequence-context combination (comparisons.py:700: chi2.sf(tau / c, f) with a fractional
// number of degrees of freedom).
NCB_HD inline double igamc(double a, double x) {
    if (a != a || x != x || !(a > 0.0)) return ncb_nan();
    if (!(x > 0.0)) return x == 0.0 ? 1.0 : ncb_nan();
    if (isinf(x)) return 0.0;
    const double lfront = a * log(x) - x - lgamma(a);   // log of x^a e^-x / Gamma(a)
    if (x < a + 1.0) {
        double term = 1.0 / a, sum = term, ap = a;
        for (int n = 0; n < 10000; ++n) {
            ap += 1.0;
            term *= x / ap;
            sum += term;
            if (term < sum * 1e-17) break;
        }
        return clip01(1.0 - sum * exp(lfront));
    }
    // SciPy's routine (cephes igam_fac) gives up once the factor x^a e^-x / Gamma(a) is below
    // exp(-MAXLOG) and returns 0 instead of a denormal; the caller replaces 0 by the smallest
    // normal number (comparisons.py:702-704), so the same cut is needed here
    if (lfront < -7.09782712893383996843e2) return 0.0;
    const double TINY = 1e-300;
    double b = x + 1.0 - a, c = 1.0 / TINY, d = 1.0 / b, h = d;
    for (int i = 1; i < 10000; ++i) {
        const double an = -(double)i * ((double)i - a);
        b += 2.0;
        d = an * d + b;
        if (fabs(d) < TINY) d = TINY;
        c = b + an / c;
        if (fabs(c) < TINY) c = TINY;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    return clip01(exp(lfront) * h);
}

NCB_HD inline long long gcd_ll(long long a, long long b) {
    while (b) { long long t = a % b; a = b; b = t; }
    return a;
}

// Mann-Whitney two-sided asymptotic p from exact integer statistics
// (scipy/stats/_mannwhitneyu.py:148-171, 520-522): u1x2 = 2*U1, tie = sum t^3 - t.
NCB_HD inline double mwu_asymptotic_p(long long u1x2, long long tie, long long n0, long long n1) {
    if (n0 <= 0 || n1 <= 0) return ncb_nan();
    double n = (double)(n0 + n1);
    double mu = (double)(n0 * n1) / 2.0;
    double U1 = (double)u1x2 / 2.0;
    double s = sqrt((double)(n0 * n1) / 12.0 * ((n + 1.0) - (double)tie / (n * (n - 1.0))));
    double num = U1 - mu;
    num -= 0.5 * ((num > 0.0) - (num < 0.0));
    double z = num / s;
    return normal_two_sided(z);
}

// Mann-Whitney exact two-sided p (no ties, a sample of at most 8 values): 2 * P(U <= umin)
// under the permutation distribution (scipy/stats/_mannwhitneyu.py:57-74, 497-515).  The
// number of arrangements with statistic u is the coefficient of q^u in the Gaussian binomial
// [a+b choose a]_q = prod_{i=1..a} (1 - q^(b+i)) / (1 - q^i), built in place in `c` and
// truncated at q^umin.  `c` must hold umin + 1 doubles; returns NaN if cap is too small.
NCB_HD inline double mwu_exact_p(long long umin, long long n0, long long n1, double *c, long long cap) {
    long long a = n0 < n1 ? n0 : n1, b = n0 < n1 ? n1 : n0;
    long long len = umin + 1;
    if (len > cap || len < 1) return ncb_nan();
    c[0] = 1.0;
    for (long long u = 1; u < len; ++u) c[u] = 0.0;
    for (long long i = 1; i <= a; ++i) {
        for (long long u = len - 1; u >= b + i; --u) c[u] -= c[u - (b + i)];
        for (long long u = i; u < len; ++u) c[u] += c[u - i];
    }
    double total = 1.0;
    for (long long i = 1; i <= a; ++i) total = total * (double)(b + i) / (double)i;
    double cdf = 0.0;
    for (long long u = 0; u < len; ++u) cdf += c[u];
    return clip01(2.0 * cdf / total);
}

// Welch two-sided p from means, unbiased variances and sizes
// (scipy/stats/_stats_py.py: _unequal_var_ttest_denom + _ttest_ind_from_stats).
NCB_HD inline double welch_p(double m0, double v0, double n0, double m1, double v1, double n1) {
    double vn0 = v0 / n0, vn1 = v1 / n1;
    double df = (vn0 + vn1) * (vn0 + vn1) / (vn0 * vn0 / (n0 - 1.0) + vn1 * vn1 / (n1 - 1.0));
    if (df != df) df = 1.0;
    double denom = sqrt(vn0 + vn1);
    double t = (m0 - m1) / denom;
    return student_t_two_sided(t, df);
}

// chi2_contingency(table + 0).pvalue for a 2x2 table of positive counts
// (scipy/stats/contingency.py:317-356: Yates, Pearson, dof 1).
NCB_HD inline double chi2_2x2_p(double o00, double o01, double o10, double o11) {
    double r0 = o00 + o01, r1 = o10 + o11, c0 = o00 + o10, c1 = o01 + o11;
    double total = r0 + r1;
    double e[4] = {r0 * c0 / total, r0 * c1 / total, r1 * c0 / total, r1 * c1 / total};
    double o[4] = {o00, o01, o10, o11};
    double stat = 0.0;
    for (int i = 0; i < 4; ++i) {
        double diff = e[i] - o[i];
        double mag = fmin(0.5, fabs(diff));
        double oc = o[i] + ((diff > 0.0) ? mag : ((diff < 0.0) ? -mag : 0.0));
        double r = oc - e[i];
        stat += r * r / e[i];
    }
    return chi2_sf_1dof(stat);
}

// Wald test of the slope of the logistic regression  condition ~ b0 + b1*cluster  on a
// 2x2 table of weights t[cond][cluster], Newton/IRLS from b = 0 (extension, no reference
// code: SURVEY 8a row L; mirrors oracle/gmm_test.py:logit_wald).
NCB_HD inline void logit_wald(double t00, double t01, double t10, double t11,
                              double *coef, double *pvalue) {
    double n0 = t00 + t10, n1 = t01 + t11;  // per-cluster totals
    double y0 = t10, y1 = t11;              // condition-1 counts per cluster
    double b0 = 0.0, b1 = 0.0;
    for (int it = 0; it < 25; ++it) {
        double mu0 = 1.0 / (1.0 + exp(-b0));
        double mu1 = 1.0 / (1.0 + exp(-(b0 + b1)));
        double w0 = n0 * mu0 * (1.0 - mu0), w1 = n1 * mu1 * (1.0 - mu1);
        double g0 = (y0 - n0 * mu0) + (y1 - n1 * mu1);
        double g1 = (y1 - n1 * mu1);
        double h00 = w0 + w1, h01 = w1, h11 = w1;
        double det = h00 * h11 - h01 * h01;
        double d0 = (h11 * g0 - h01 * g1) / det;
        double d1 = (-h01 * g0 + h00 * g1) / det;
        b0 += d0;
        b1 += d1;
        if (fabs(d0) + fabs(d1) < 1e-10) break;
    }
    double mu0 = 1.0 / (1.0 + exp(-b0));
    double mu1 = 1.0 / (1.0 + exp(-(b0 + b1)));
    double w0 = n0 * mu0 * (1.0 - mu0), w1 = n1 * mu1 * (1.0 - mu1);
    double det = (w0 + w1) * w1 - w1 * w1;
    double var_b1 = (w0 + w1) / det;
    double z = b1 / sqrt(var_b1);
    *coef = b1;
    *pvalue = normal_two_sided(z);
}

// ---- exact two-sample KS p-value, sequential forms (one thread per problem) --------
// scipy/stats/_stats_py.py:7713-7748 (_compute_prob_outside_square), equal sizes.
// Explicitly rounded operations: no FMA contraction, so every step is the IEEE
// operation SciPy performs.
#if defined(__CUDA_ARCH__)
#define NCB_DMUL(a, b) __dmul_rn((a), (b))
#define NCB_DADD(a, b) __dadd_rn((a), (b))
#define NCB_DDIV(a, b) __ddiv_rn((a), (b))
#else
#define NCB_DMUL(a, b) ((a) * (b))
#define NCB_DADD(a, b) ((a) + (b))
#define NCB_DDIV(a, b) ((a) / (b))
#endif

NCB_HD inline double ks_prob_outside_square(long long n, long long h) {
    double P = 0.0;
    long long k = n / h;
    while (k >= 0) {
        double p1 = 1.0;
        for (long long j = 0; j < h; ++j) {
            p1 = NCB_DDIV(NCB_DMUL((double)(n - k * h - j), p1), (double)(n + k * h + j + 1));
        }
        P = NCB_DMUL(p1, NCB_DADD(1.0, -P));
        --k;
    }
    return 2.0 * P;
}

// Equal sample sizes: the Horner product, and SciPy's behaviour when its rounding lands
// outside [0, 1] (only ever just above 1, when the true probability is 1 - O(1e-16)):
// _attempt_exact_2kssamp reports failure and ks_2samp falls back to the one-sample Kolmogorov
// distribution kstwo.sf(d = h/n, N = round-half-even(n/2)) (_stats_py.py:8156-8171).  In that
// regime only the first two branches of kolmogn matter:
//     d <= 1/(2N): 1          1/(2N) < d <= 1/N: 1 - N! (2d - 1/N)^N        else: 1
// (the third branch drops a term below 4e-13, checked against SciPy for every n that can
// reach it; tests/test_math_host.py).
// `p` = 2 * the Horner product (ks_prob_outside_square, or a faster evaluation with the same bits)
NCB_HD inline double ks_equal_finish(double p, long long n, long long h) {
    if (p >= 0.0 && p <= 1.0) return p;
    const double d = (double)h / (double)n;
    const double N = rint(0.5 * (double)n);
    if (N < 1.0 || d <= 1.0 / (2.0 * N)) return 1.0;
    if (d <= 1.0 / N) {
        const double base = 2.0 * d - 1.0 / N;
        if (!(base > 0.0)) return 1.0;
        return clip01(1.0 - exp(lgamma(N + 1.0) + N * log(base)));
    }
    return 1.0;
}
NCB_HD inline double ks_equal_p(long long n, long long h) {
    return ks_equal_finish(ks_prob_outside_square(n, h), n, h);
}

}  // namespace ncb

// Host build of nanocompore_b200/csrc/ncb_math.cuh for the CPU test-suite (tests/test_math_host.py).
// Test infrastructure only: it lets the scalar p-value routines the kernels use be checked
// against SciPy without a GPU.  Nothing in the product loads this.
#include "ncb_math.cuh"
#include <vector>
extern "C" {
double hc_student_t_two_sided(double t, double df) { return ncb::student_t_two_sided(t, df); }
double hc_welch_p(double m0, double v0, double n0, double m1, double v1, double n1) { return ncb::welch_p(m0, v0, n0, m1, v1, n1); }
double hc_mwu_asymptotic_p(long long u1x2, long long tie, long long n0, long long n1) { return ncb::mwu_asymptotic_p(u1x2, tie, n0, n1); }
double hc_mwu_exact_p(long long umin, long long n0, long long n1) {
    std::vector<double> c((size_t)umin + 2);
    return ncb::mwu_exact_p(umin, n0, n1, c.data(), (long long)c.size());
}
double hc_chi2_2x2_p(double a, double b, double c, double d) { return ncb::chi2_2x2_p(a, b, c, d); }
void hc_logit_wald(double a, double b, double c, double d, double *coef, double *p) { ncb::logit_wald(a, b, c, d, coef, p); }
double hc_ks_prob_outside_square(long long n, long long h) { return ncb::ks_prob_outside_square(n, h); }
double hc_ks_equal_p(long long n, long long h) { return ncb::ks_equal_p(n, h); }
void hc_quot_rn(const double *a, const double *b, double *q, long long n) {
    for (long long i = 0; i < n; ++i) q[i] = ncb::quot_rn(a[i], b[i], 1.0 / b[i]);
}
double hc_igamc(double a, double x) { return ncb::igamc(a, x); }
double hc_normal_two_sided(double z) { return ncb::normal_two_sided(z); }
}

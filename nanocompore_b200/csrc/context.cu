// Sequence-context combination of p-value columns (Hou's method for a weighted Fisher
// combination of dependent p-values).
//
// Replaces, for every transcript and p-value column of a batch at once,
//   TranscriptComparator._combine_context_pvalues   comparisons.py:436-474
//   cross_corr_matrix                               comparisons.py:627-651
//   combine_pvalues_hou                             comparisons.py:659-705
// The reference lays the p-values of a transcript on a dense track (1 where a position has no
// row, NaN -> 1), correlates the track with its own shifts by -context..context (np.corrcoef of
// every pair of shifted windows), and then walks a window of 2*context+1 positions over the
// track calling chi2.sf(tau / c, f) per row in a Python loop (np.vectorize).
//
// Here: one CTA per (transcript, column).  The track lives in global scratch (a few KB, L2
// resident); warps share the (2c+1)(2c+2)/2 correlations between them (two passes: means, then
// centred products; fixed lane order, so the result does not depend on scheduling); one thread
// per row evaluates the window and the regularised incomplete gamma function (ncb_math.cuh).
// Parity is by tolerance (1e-5 relative on log10 p): summation order differs from NumPy's.
#include "ncb_internal.h"
#include "ncb_math.cuh"

namespace ncb {

#define CTX_THREADS 256
#define CTX_K (2 * NCB_CTX_MAX + 1)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct CtxArgs {
    int64_t n_tx, n_rows, track_total;
    const int64_t *tx_offsets;     // [n_tx + 1]
    const int64_t *track_offsets;  // [n_tx + 1]
    const int32_t *idx;            // [n_rows] position - first position of the transcript
    const double *pvalues;         // [n_cols][n_rows]
    double *tracks;                // [n_cols][track_total]
    double *combined;              // [n_cols][n_rows]
    int32_t *status;               // [n_cols][n_tx]
    int32_t context;
    double weights[CTX_K];
};

__global__ void __launch_bounds__(CTX_THREADS)
context_kernel(CtxArgs a) {
    __shared__ double mean[CTX_K];
    __shared__ double cov[CTX_K][CTX_K];
    __shared__ double hou[2];   // c (scale of tau), f (degrees of freedom)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NWARP = CTX_THREADS / 32;
    const int64_t t = blockIdx.x;
    const int col = blockIdx.y;
    const int64_t r0 = a.tx_offsets[t], r1 = a.tx_offsets[t + 1];
    const int n = (int)(r1 - r0);
    if (n == 0) {
        if (tid == 0) a.status[col * a.n_tx + t] = NCB_CTX_OK;
        return;
    }
    const int c = a.context, k = 2 * c + 1;
    const int L = (int)(a.track_offsets[t + 1] - a.track_offsets[t]);
    double *x = a.tracks + (int64_t)col * a.track_total + a.track_offsets[t];
    const double *pv = a.pvalues + (int64_t)col * a.n_rows + r0;
    double *out = a.combined + (int64_t)col * a.n_rows + r0;
    const int32_t *idx = a.idx + r0;

    // the track: 1 where the transcript has no row (comparisons.py:456-459), NaN -> 1 (635)
    for (int u = tid; u < L; u += CTX_THREADS) x[u] = 1.0;
    __syncthreads();
    int invalid = 0, not_one = 0;
    for (int r = tid; r < n; r += CTX_THREADS) {
        const double p = pv[r];
        const double v = (p != p) ? 1.0 : p;
        invalid |= (v == 0.0 || isinf(v) || v > 1.0);   // comparisons.py:636
        not_one |= (v != 1.0);
        x[idx[r]] = v;
    }
    invalid = __syncthreads_or(invalid);
    not_one = __syncthreads_or(not_one);
    int st = NCB_CTX_OK;
    if (L < 3 * c + 3) st = NCB_CTX_TOO_SHORT;            // comparisons.py:632-633
    else if (invalid) st = NCB_CTX_INVALID_PVALUE;
    if (tid == 0) a.status[col * a.n_tx + t] = st;
    if (st != NCB_CTX_OK) {
        for (int r = tid; r < n; r += CTX_THREADS) out[r] = ncb_nan();
        return;
    }

    // correlation of the windows x[2c - i .. 2c - i + nw) , i = 0..2c (np.roll by i - c, then
    // [c : L - c]; comparisons.py:643-650)
    const int nw = L - 2 * c;
    if (!not_one) {                                       // comparisons.py:638-639
        for (int q = tid; q < k * k; q += CTX_THREADS) cov[q / k][q % k] = 1.0;
    } else {
        for (int i = warp; i < k; i += NWARP) {
            const double *xi = x + (2 * c - i);
            double s = 0.0;
            for (int u = lane; u < nw; u += 32) s += xi[u];
            s = warp_sum(s);
            if (lane == 0) mean[i] = s / (double)nw;
        }
        __syncthreads();
        const int npair = k * (k + 1) / 2;
        for (int q = warp; q < npair; q += NWARP) {
            int i = 0, rem = q;
            while (rem > i) { rem -= i + 1; ++i; }        // q = i (i + 1) / 2 + j, j <= i
            const int j = rem;
            const double *xi = x + (2 * c - i), *xj = x + (2 * c - j);
            const double mi = mean[i], mj = mean[j];
            double s = 0.0;
            for (int u = lane; u < nw; u += 32) s = fma(xi[u] - mi, xj[u] - mj, s);
            s = warp_sum(s);
            if (lane == 0) cov[i][j] = cov[j][i] = s / (double)(nw - 1);
        }
        __syncthreads();
        // np.corrcoef: c / stddev[:, None] / stddev[None, :], clipped to [-1, 1]
        double rr[(CTX_K * CTX_K + CTX_THREADS - 1) / CTX_THREADS];
        int m = 0;
        for (int q = tid; q < k * k; q += CTX_THREADS, ++m) {
            const int i = q / k, j = q % k;
            double r = cov[i][j] / sqrt(cov[i][i]) / sqrt(cov[j][j]);
            if (r > 1.0) r = 1.0;
            if (r < -1.0) r = -1.0;
            rr[m] = r;
        }
        __syncthreads();
        m = 0;
        for (int q = tid; q < k * k; q += CTX_THREADS, ++m) cov[q / k][q % k] = rr[m];
    }
    __syncthreads();
    if (tid == 0) {
        // comparisons.py:681-698 (Kost & McDermott eq. 8; same order of accumulation)
        double cov_sum = 0.0, sw_sum = 0.0, w_sum = 0.0;
        for (int i = 0; i < k; ++i) {
            for (int j = i + 1; j < k; ++j) {
                const double r = cov[i][j];
                cov_sum += a.weights[i] * a.weights[j] * ((3.263 * r) + (0.710 * (r * r)) + (0.027 * (r * r * r)));
            }
            sw_sum += a.weights[i] * a.weights[i];
            w_sum += a.weights[i];
        }
        hou[0] = (2.0 * sw_sum + cov_sum) / (2.0 * w_sum);
        hou[1] = (4.0 * (w_sum * w_sum)) / (2.0 * sw_sum + cov_sum);
    }
    __syncthreads();
    const double scale = hou[0], dof = hou[1];
    for (int r = tid; r < n; r += CTX_THREADS) {
        const double centre = pv[r];
        double res;
        if (centre != centre) {
            res = ncb_nan();                              // comparisons.py:464-465
        } else {
            const int u0 = idx[r] - c;
            double tau = 0.0;
            bool all_one = true;
            for (int i = 0; i < k; ++i) {
                const int u = u0 + i;
                const double v = (u < 0 || u >= L) ? 1.0 : x[u];
                all_one = all_one && (v == 1.0);
                tau += a.weights[i] * (-2.0 * log(v));
            }
            if (all_one) {
                res = 1.0;                                // comparisons.py:675-676
            } else {
                res = igamc(0.5 * dof, 0.5 * (tau / scale));   // chi2.sf(tau / c, f)
                if (res == 0.0) res = 2.2250738585072014e-308; // comparisons.py:702-704
            }
        }
        out[r] = res;
    }
}

}  // namespace ncb

using namespace ncb;

int ncb_launch_context(int64_t n_tx, int64_t n_rows, int64_t track_total, int32_t n_cols,
                       const int64_t *tx_offsets, const int64_t *track_offsets, const int32_t *idx,
                       const double *pvalues, double *tracks, double *combined, int32_t *status,
                       int32_t context, const double *weights, cudaStream_t s) {
    if (n_tx <= 0 || n_cols <= 0) return 0;
    CtxArgs a;
    a.n_tx = n_tx; a.n_rows = n_rows; a.track_total = track_total;
    a.tx_offsets = tx_offsets; a.track_offsets = track_offsets; a.idx = idx;
    a.pvalues = pvalues; a.tracks = tracks; a.combined = combined; a.status = status;
    a.context = context;
    for (int i = 0; i < CTX_K; ++i) a.weights[i] = i < 2 * context + 1 ? weights[i] : 0.0;
    dim3 grid((unsigned)n_tx, (unsigned)n_cols);
    context_kernel<<<grid, CTX_THREADS, 0, s>>>(a);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// Benjamini-Hochberg q-values over one p-value column, NaNs skipped, as a device sort-scan.
//
// Replaces Postprocessor._multipletests_filter_nan (nanocompore/postprocessing.py:255-285),
// i.e. statsmodels multipletests(method='fdr_bh') on the non-NaN entries:
//     sort ascending;  q_r = p_r / (r / m)  (r = 1..m);  running minimum from the largest p
//     downwards;  clip at 1;  un-sort.
// Here: keys sorted DESCENDING (NaN keyed lowest, so it sorts last), q_s = p_s / ((m - s)/m),
// forward inclusive min-scan, scatter through the carried indices.  The device-wide radix
// sort and scan are CUB's (library primitives, like cuBLAS for a GEMM); the key transform,
// the q formula and the scatter are the kernels below.
#include <cub/cub.cuh>
#include "ncb_internal.h"

namespace ncb {

typedef unsigned long long u64;

__device__ __forceinline__ u64 f64_orderable(double d) {
    u64 u = (u64)__double_as_longlong(d);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

__global__ void bh_keys_kernel(int64_t n, const double *p, u64 *keys, uint32_t *idx,
                               unsigned long long *n_valid) {
    unsigned long long local = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        double v = p[i];
        bool ok = (v == v);
        keys[i] = ok ? f64_orderable(v) : 0ull;   // NaN sorts after every number (descending)
        idx[i] = (uint32_t)i;
        local += ok ? 1ull : 0ull;
    }
    for (int m = 16; m > 0; m >>= 1) local += __shfl_xor_sync(0xffffffffu, local, m);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(n_valid, local);
}

// raw q in descending-p order: position s holds ascending rank M - above - s, where the column has M
// numbers in all and `above` of them are larger than every number of this segment (a whole column:
// M = m, above = 0)
__global__ void bh_raw_kernel(int64_t n, const double *p, const uint32_t *idx_sorted,
                              const unsigned long long *n_valid, double *qraw, int64_t above, int64_t m_total) {
    const int64_t m = (int64_t)*n_valid;
    const int64_t M = m_total < 0 ? m : m_total;
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < n;
         s += (int64_t)gridDim.x * blockDim.x) {
        double q = __longlong_as_double(0x7ff0000000000000LL);   // +inf: neutral for min
        if (s < m) {
            double pv = p[idx_sorted[s]];
            double ecdf = __ddiv_rn((double)(M - above - s), (double)M);
            q = __ddiv_rn(pv, ecdf);
        }
        qraw[s] = q;
    }
}

__global__ void bh_scatter_kernel(int64_t n, const uint32_t *idx_sorted,
                                  const unsigned long long *n_valid, const double *qmin,
                                  double scale, double *q_out) {
    const int64_t m = (int64_t)*n_valid;
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < n;
         s += (int64_t)gridDim.x * blockDim.x) {
        double q;
        if (s < m) {
            q = qmin[s];
            if (q > 1.0) q = 1.0;
            q = __dmul_rn(q, scale);
        } else {
            q = __longlong_as_double(0x7ff8000000000000LL);
        }
        q_out[idx_sorted[s]] = q;
    }
}

struct MinOp {
    __device__ __forceinline__ double operator()(double a, double b) const { return b < a ? b : a; }
};

struct BhLayout {
    size_t keys_in, keys_out, idx_in, idx_out, qraw, qmin, nvalid, cub_tmp, cub_bytes, total;
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static BhLayout bh_layout(int64_t n) {
    BhLayout L;
    size_t off = 0;
    L.keys_in = off; off += align256(n * sizeof(u64));
    L.keys_out = off; off += align256(n * sizeof(u64));
    L.idx_in = off; off += align256(n * sizeof(uint32_t));
    L.idx_out = off; off += align256(n * sizeof(uint32_t));
    L.qraw = off; off += align256(n * sizeof(double));
    L.qmin = off; off += align256(n * sizeof(double));
    L.nvalid = off; off += 256;
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, sort_bytes, (const u64 *)nullptr, (u64 *)nullptr,
                                              (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
    cub::DeviceScan::InclusiveScan(nullptr, scan_bytes, (const double *)nullptr, (double *)nullptr,
                                   MinOp(), (int)n);
    L.cub_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
    L.cub_tmp = off; off += align256(L.cub_bytes);
    L.total = off;
    return L;
}

}  // namespace ncb

using namespace ncb;

size_t ncb_bh_fdr_scratch_bytes(int64_t n) {
    if (n <= 0) return 256;
    return bh_layout(n).total;
}

int ncb_launch_bh_fdr(int64_t n, const double *p, double *q, double scale, int64_t above, int64_t m_total,
                      void *scratch, size_t scratch_bytes, cudaStream_t s) {
    if (n <= 0) return 0;
    if (n > 0x7fffffffLL) return -1;
    BhLayout L = bh_layout(n);
    if (scratch_bytes < L.total) return -1;
    char *base = (char *)scratch;
    u64 *keys_in = (u64 *)(base + L.keys_in), *keys_out = (u64 *)(base + L.keys_out);
    uint32_t *idx_in = (uint32_t *)(base + L.idx_in), *idx_out = (uint32_t *)(base + L.idx_out);
    double *qraw = (double *)(base + L.qraw), *qmin = (double *)(base + L.qmin);
    unsigned long long *nvalid = (unsigned long long *)(base + L.nvalid);
    if (cudaMemsetAsync(nvalid, 0, sizeof(unsigned long long), s) != cudaSuccess) return -1;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    bh_keys_kernel<<<blocks, 256, 0, s>>>(n, p, keys_in, idx_in, nvalid);
    size_t tmp = L.cub_bytes;
    if (cub::DeviceRadixSort::SortPairsDescending(base + L.cub_tmp, tmp, keys_in, keys_out, idx_in,
                                                  idx_out, (int)n, 0, 64, s) != cudaSuccess)
        return -1;
    bh_raw_kernel<<<blocks, 256, 0, s>>>(n, p, idx_out, nvalid, qraw, above, m_total);
    tmp = L.cub_bytes;
    if (cub::DeviceScan::InclusiveScan(base + L.cub_tmp, tmp, qraw, qmin, MinOp(), (int)n, s) !=
        cudaSuccess)
        return -1;
    bh_scatter_kernel<<<blocks, 256, 0, s>>>(n, idx_out, nvalid, qmin, scale, q);
    // 3 kernels of ours; the sort and the scan launch CUB's own
    return cudaGetLastError() == cudaSuccess ? 3 : -1;
}

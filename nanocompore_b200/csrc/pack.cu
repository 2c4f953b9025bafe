// Host-side CSR packer: the reference's dense per-transcript tensor (positions x reads x vars,
// NaN = the read does not cover the position; run.py:621-855 readers) -> the ragged batch layout
// of ncb_compare (valid reads only, position after position, read order preserved).
// Multi-threaded over positions; no CUDA calls.  This is the "reader -> pinned CSR" step of
// a reader that decodes SQLite blobs / BAM tags straight into these arrays skips
// the dense tensor altogether.
#include <math.h>
#include <string.h>
#include <thread>
#include <vector>
#include "ncb.h"

namespace {

template <class F>
void parallel_for(int64_t n, int threads, F f) {
    if (threads < 1) threads = 1;
    if (threads == 1 || n < 2 * threads) {
        f(0, n);
        return;
    }
    std::vector<std::thread> pool;
    const int64_t chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        const int64_t lo = t * chunk, hi = lo + chunk < n ? lo + chunk : n;
        if (lo >= hi) break;
        pool.emplace_back([=] { f(lo, hi); });
    }
    for (auto &th : pool) th.join();
}

inline bool covered(const float *row, int V) {
    // a read is kept at a position when intensity or dwell is a number
    return !(row[0] != row[0]) || !(row[1] != row[1]);
}

}  // namespace

extern "C" {

int ncb_pack_count(const float *data, int64_t n_positions, int32_t n_reads, int32_t n_vars,
                   int64_t *counts, int threads) {
    if (!data || !counts || n_positions < 0 || n_reads < 0 || n_vars < 2) return NCB_ERR_INVALID;
    parallel_for(n_positions, threads, [=](int64_t lo, int64_t hi) {
        for (int64_t p = lo; p < hi; ++p) {
            const float *row = data + p * (int64_t)n_reads * n_vars;
            int64_t c = 0;
            for (int r = 0; r < n_reads; ++r) c += covered(row + (int64_t)r * n_vars, n_vars) ? 1 : 0;
            counts[p] = c;
        }
    });
    return NCB_OK;
}

int ncb_pack_fill(const float *data, const uint8_t *label, int64_t n_positions, int32_t n_reads,
                  int32_t n_vars, const int64_t *pos_offsets, float *signal_out, uint8_t *label_out,
                  int threads) {
    if (!data || !label || !pos_offsets || !signal_out || !label_out || n_positions < 0 ||
        n_reads < 0 || n_vars < 2)
        return NCB_ERR_INVALID;
    parallel_for(n_positions, threads, [=](int64_t lo, int64_t hi) {
        for (int64_t p = lo; p < hi; ++p) {
            const float *row = data + p * (int64_t)n_reads * n_vars;
            int64_t o = pos_offsets[p];
            for (int r = 0; r < n_reads; ++r) {
                const float *e = row + (int64_t)r * n_vars;
                if (covered(e, n_vars)) {
                    signal_out[2 * o] = e[0];
                    signal_out[2 * o + 1] = e[1];
                    label_out[o] = label[r];
                    ++o;
                }
            }
        }
    });
    return NCB_OK;
}

}  // extern "C"

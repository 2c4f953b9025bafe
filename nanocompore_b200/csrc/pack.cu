// Host-side CSR packer: the reference's dense per-transcript tensor (positions x reads x vars,
// NaN = the read does not cover the position; run.py:621-855 readers) -> the ragged batch layout
// of ncb_compare (valid reads only, position after position, read order preserved).
// Multi-threaded over positions; no CUDA calls.  This is the "reader -> pinned CSR" step of
// a reader that decodes SQLite blobs / BAM tags straight into these arrays skips
// the dense tensor altogether.
#include <math.h>
#include <string.h>
#include <atomic>
#include <exception>
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
        try {
            pool.emplace_back([=] { f(lo, hi); });
        } catch (const std::exception &) {
            f(lo, hi);   // no thread (or no memory for one) to be had: the chunk runs here
        }
    }
    for (auto &th : pool) th.join();
}

inline bool covered(const float *row, int V) {
    // a read is kept at a position when intensity or dwell is a number
    return !(row[0] != row[0]) || !(row[1] != row[1]);
}

// float32 -> int16 code of NCB_LAYOUT_CSR_I16: NaN -> -32768; an integer in [-32767, 32767] -> itself;
// anything else is not representable (`bad` is raised, the code written is meaningless)
inline int16_t code_i16(float x, bool &bad) {
    if (x != x) return (int16_t)-32768;
    if (!(x >= -32767.0f && x <= 32767.0f)) { bad = true; return 0; }
    const int16_t c = (int16_t)x;
    if ((float)c != x) bad = true;
    return c;
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

int ncb_pack_fill_i16(const float *data, const uint8_t *label, int64_t n_positions, int32_t n_reads,
                      int32_t n_vars, const int64_t *pos_offsets, int16_t *signal_out, uint8_t *label_out,
                      int threads) {
    if (!data || !label || !pos_offsets || !signal_out || !label_out || n_positions < 0 ||
        n_reads < 0 || n_vars < 2)
        return NCB_ERR_INVALID;
    std::atomic<int> any_bad(0);
    std::atomic<int> *flag = &any_bad;
    parallel_for(n_positions, threads, [=](int64_t lo, int64_t hi) {
        bool bad = false;
        for (int64_t p = lo; p < hi; ++p) {
            const float *row = data + p * (int64_t)n_reads * n_vars;
            int64_t o = pos_offsets[p];
            for (int r = 0; r < n_reads; ++r) {
                const float *e = row + (int64_t)r * n_vars;
                if (covered(e, n_vars)) {
                    signal_out[2 * o] = code_i16(e[0], bad);
                    signal_out[2 * o + 1] = code_i16(e[1], bad);
                    label_out[o] = label[r];
                    ++o;
                }
            }
        }
        if (bad) flag->store(1);
    });
    return any_bad.load() ? NCB_ERR_INVALID : NCB_OK;
}

// ---- planar read-major rows -> CSR ---------------------------------------------------------
// Rows are what the readers produce per read: intensity[n_positions], dwell[n_positions] with NaN
// where the read does not cover the position (SQLite signal_data blobs, run.py:795-801, used in
// place; rows decoded from Uncalled4 BAM tags by ncb_bam_read_uncalled4).  The dense
// (positions, reads, vars) tensor of the reference (run.py:833-838: concatenate, transpose,
// stack) is never built.  Threads own tiles of positions; inside a tile the rows are walked in
// row order, so the reads of a position keep the row order (the order the mixture fit sees).
// motor_offset > 0 adds the third variable of Worker._prepare_data (run.py:575-584): the dwell of
// the same read `motor_offset` positions downstream.

namespace {
const int64_t ROW_TILE = 512;   // positions per tile: 2 KB of every row, cursors stay in L1

inline bool is_num(float x) { return x == x; }
}  // namespace

int ncb_pack_rows_count(const float *const *intensity, const float *const *dwell, int64_t n_rows,
                        int64_t n_positions, int32_t motor_offset, int64_t *counts, int threads) {
    if (n_rows < 0 || n_positions < 0 || motor_offset < 0 || (n_positions > 0 && !counts) ||
        (n_rows > 0 && (!intensity || !dwell)))
        return NCB_ERR_INVALID;
    const int64_t tiles = (n_positions + ROW_TILE - 1) / ROW_TILE;
    parallel_for(tiles, threads, [=](int64_t t_lo, int64_t t_hi) {
        for (int64_t t = t_lo; t < t_hi; ++t) {
            const int64_t p0 = t * ROW_TILE, p1 = p0 + ROW_TILE < n_positions ? p0 + ROW_TILE : n_positions;
            int64_t c[ROW_TILE];
            for (int64_t i = 0; i < p1 - p0; ++i) c[i] = 0;
            for (int64_t r = 0; r < n_rows; ++r) {
                const float *x = intensity[r], *y = dwell[r];
                for (int64_t p = p0; p < p1; ++p) {
                    bool keep = is_num(x[p]) || is_num(y[p]);
                    if (motor_offset > 0 && p + motor_offset < n_positions) keep |= is_num(y[p + motor_offset]);
                    c[p - p0] += keep ? 1 : 0;
                }
            }
            for (int64_t i = 0; i < p1 - p0; ++i) counts[p0 + i] = c[i];
        }
    });
    return NCB_OK;
}

int ncb_pack_rows_fill(const float *const *intensity, const float *const *dwell, const uint8_t *row_label,
                       int64_t n_rows, int64_t n_positions, int32_t motor_offset,
                       const int64_t *pos_offsets, float *signal_out, uint8_t *label_out, int threads) {
    if (n_rows < 0 || n_positions < 0 || motor_offset < 0 ||
        (n_positions > 0 && (!pos_offsets || !signal_out || !label_out)) ||
        (n_rows > 0 && (!intensity || !dwell || !row_label)))
        return NCB_ERR_INVALID;
    const int V = motor_offset > 0 ? 3 : 2;
    const float NaN = nanf("");
    const int64_t tiles = (n_positions + ROW_TILE - 1) / ROW_TILE;
    parallel_for(tiles, threads, [=](int64_t t_lo, int64_t t_hi) {
        for (int64_t t = t_lo; t < t_hi; ++t) {
            const int64_t p0 = t * ROW_TILE, p1 = p0 + ROW_TILE < n_positions ? p0 + ROW_TILE : n_positions;
            int64_t cur[ROW_TILE];
            for (int64_t i = 0; i < p1 - p0; ++i) cur[i] = pos_offsets[p0 + i];
            for (int64_t r = 0; r < n_rows; ++r) {
                const float *x = intensity[r], *y = dwell[r];
                const uint8_t lab = row_label[r];
                for (int64_t p = p0; p < p1; ++p) {
                    const float xv = x[p], yv = y[p];
                    float mv = NaN;
                    if (motor_offset > 0 && p + motor_offset < n_positions) mv = y[p + motor_offset];
                    if (is_num(xv) || is_num(yv) || is_num(mv)) {
                        const int64_t o = cur[p - p0]++;
                        signal_out[V * o] = xv;
                        signal_out[V * o + 1] = yv;
                        if (V == 3) signal_out[V * o + 2] = mv;
                        label_out[o] = lab;
                    }
                }
            }
        }
    });
    return NCB_OK;
}

int ncb_pack_rows_fill_i16(const float *const *intensity, const float *const *dwell,
                           const uint8_t *row_label, int64_t n_rows, int64_t n_positions,
                           int32_t motor_offset, const int64_t *pos_offsets, int16_t *signal_out,
                           uint8_t *label_out, int threads) {
    if (n_rows < 0 || n_positions < 0 || motor_offset < 0 ||
        (n_positions > 0 && (!pos_offsets || !signal_out || !label_out)) ||
        (n_rows > 0 && (!intensity || !dwell || !row_label)))
        return NCB_ERR_INVALID;
    const int V = motor_offset > 0 ? 3 : 2;
    const float NaN = nanf("");
    const int64_t tiles = (n_positions + ROW_TILE - 1) / ROW_TILE;
    std::atomic<int> any_bad(0);
    std::atomic<int> *flag = &any_bad;
    parallel_for(tiles, threads, [=](int64_t t_lo, int64_t t_hi) {
        bool bad = false;
        for (int64_t t = t_lo; t < t_hi; ++t) {
            const int64_t p0 = t * ROW_TILE, p1 = p0 + ROW_TILE < n_positions ? p0 + ROW_TILE : n_positions;
            int64_t cur[ROW_TILE];
            for (int64_t i = 0; i < p1 - p0; ++i) cur[i] = pos_offsets[p0 + i];
            for (int64_t r = 0; r < n_rows; ++r) {
                const float *x = intensity[r], *y = dwell[r];
                const uint8_t lab = row_label[r];
                for (int64_t p = p0; p < p1; ++p) {
                    const float xv = x[p], yv = y[p];
                    float mv = NaN;
                    if (motor_offset > 0 && p + motor_offset < n_positions) mv = y[p + motor_offset];
                    if (is_num(xv) || is_num(yv) || is_num(mv)) {
                        const int64_t o = cur[p - p0]++;
                        signal_out[V * o] = code_i16(xv, bad);
                        signal_out[V * o + 1] = code_i16(yv, bad);
                        if (V == 3) signal_out[V * o + 2] = code_i16(mv, bad);
                        label_out[o] = lab;
                    }
                }
            }
        }
        if (bad) flag->store(1);
    });
    return any_bad.load() ? NCB_ERR_INVALID : NCB_OK;
}

int ncb_pack_label_runs(const int64_t *pos_offsets, const uint8_t *label, int64_t n_positions,
                        const uint8_t *run_labels, int32_t n_runs, uint16_t *run_ends, int threads) {
    if (n_positions < 0 || n_runs < 1 || n_runs > NCB_MAX_RUNS || !run_labels ||
        (n_positions > 0 && (!pos_offsets || !run_ends)))
        return NCB_ERR_INVALID;
    int16_t run_of[256];
    for (int i = 0; i < 256; ++i) run_of[i] = -1;
    for (int j = 0; j < n_runs; ++j) {
        if (run_of[run_labels[j]] >= 0) return NCB_ERR_INVALID;   // a label names one run
        run_of[run_labels[j]] = (int16_t)j;
    }
    const int16_t *lut = run_of;
    std::atomic<int> any_bad(0);
    std::atomic<int> *flag = &any_bad;
    parallel_for(n_positions, threads, [=](int64_t lo, int64_t hi) {
        bool bad = false;
        for (int64_t p = lo; p < hi; ++p) {
            const int64_t o = pos_offsets[p], n = pos_offsets[p + 1] - o;
            uint16_t *ends = run_ends + p * (int64_t)n_runs;
            if (n < 0 || n > 65535 || (n > 0 && !label)) { bad = true; for (int j = 0; j < n_runs; ++j) ends[j] = 0; continue; }
            int run = 0;
            for (int64_t e = 0; e < n; ++e) {
                const int r = lut[label[o + e]];
                if (r < run) { bad = true; break; }           // unknown label (-1) or out of order
                while (run < r) ends[run++] = (uint16_t)e;
            }
            while (run < n_runs) ends[run++] = (uint16_t)n;
        }
        if (bad) flag->store(1);
    });
    return any_bad.load() ? NCB_ERR_INVALID : NCB_OK;
}

}  // extern "C"

/*
 * ncb.h -- C ABI of libncb: nanocompore's per-position two-condition comparison on B200.
 *
 * This is the drop-in boundary for ONE path of nanocompore v2.0.0:
 *   Worker._prepare_data / _filter_low_cov_positions        nanocompore/run.py:511-515, 532-601
 *   TranscriptComparator.compare_transcript                 nanocompore/comparisons.py:41-148
 *     _add_shift_stats                                      comparisons.py:406-433
 *     outlier filter + standardise                          comparisons.py:95-120
 *     _nonparametric_test (KS / MW / TT via SciPy)          comparisons.py:218-252
 *     _gmm_test_split (gmm_gpu.GMM, contingency, chi2, LOR) comparisons.py:338-392
 *     _get_cluster_counts / _get_mod_cluster                comparisons.py:477-605
 *   Postprocessor._multipletests_filter_nan (BH FDR)        nanocompore/postprocessing.py:255-285
 *
 * The reference has no FFI: it is pure Python calling torch / SciPy / gmm_gpu.  The
 * entry points below are what a ctypes binding inside the reference's
 * TranscriptComparator would call (INTEGRATION.md shows that stub).
 *
 * Conventions: plain pointers and sizes, no C++ / torch types; every function returns
 * 0 on success or a negative ncb_status; nothing throws across the ABI;
 * ncb_last_error() gives the message of the last failure on that handle.  The caller
 * owns every buffer it passes in; the library owns its device scratch.  One handle per
 * process per device; a handle is not thread-safe.
 */
#ifndef NCB_H
#define NCB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NCB_ABI_VERSION 3

typedef struct ncb_handle ncb_handle;

typedef enum {
    NCB_OK = 0,
    NCB_ERR_INVALID = -1,      /* bad argument                                          */
    NCB_ERR_CUDA = -2,         /* CUDA runtime error (message in ncb_last_error)        */
    NCB_ERR_NOMEM = -3,        /* device/pinned allocation failed (maps to torch.OutOfMemoryError) */
    NCB_ERR_TOO_MANY_READS = -4 /* returned by ncb_wait: a position of a batch since the last ncb_wait holds
                                   more entries than NCB_MAX_READS (or than ncb_batch.n_reads declared); that
                                   position has status NCB_POS_TOO_MANY_READS and the NaN / 0 defaults in its
                                   result row, every other position of the batch is complete             */
} ncb_status;

/* Most valid reads one position may hold (the reference caps at 5000 per condition,
 * config.py:200 `downsample_high_coverage`, run.py:518-529). */
#define NCB_MAX_READS 10240
#define NCB_MAX_SAMPLES 64

/* comparison_methods (config.py:152-164), as a bit mask */
enum {
    NCB_TEST_KS = 1,
    NCB_TEST_MW = 2,
    NCB_TEST_TT = 4,
    NCB_TEST_GMM = 8,
    NCB_TEST_LOGIT = 16,  /* extension: adds GMM_logit_* (no reference code) */
    NCB_TEST_AUTO = 32    /* comparisons.py:122-126, 165-167: coverage < 500 -> KS else GMM */
};

/* cluster_counts (config.py:171, comparisons.py:368-376) */
enum { NCB_COUNTS_NONE = 0, NCB_COUNTS_HARD = 1, NCB_COUNTS_SOFT = 2 };

/* where a buffer lives */
enum { NCB_MEM_HOST = 0, NCB_MEM_DEVICE = 1 };

/* input layout */
enum {
    NCB_LAYOUT_CSR = 0,   /* ragged: pos_offsets[P+1], signal float32[E][V], label[E]     */
    NCB_LAYOUT_DENSE = 1, /* the reference's tensor: signal[P][R][V] with NaN, label[R]   */
    NCB_LAYOUT_CSR_I16 = 2, /* ragged like NCB_LAYOUT_CSR with the entries as int16 codes: signal points at
                             int16[E][V], value = (float)code, -32768 = NaN.  This is what Uncalled4 stores
                             (tags uc / ul are int16 arrays copied into the float32 tensor as they are,
                             uncalled4.py:118-192, -32768 -> NaN at 103-107), so for that input the layout is
                             lossless at 5 instead of 9 bytes per entry on the host-to-device link          */
    NCB_LAYOUT_CSR_I16_RUNS = 3 /* NCB_LAYOUT_CSR_I16 without the label byte of every entry: the reference
                             concatenates the reads of a transcript sample after sample (run.py:820-826), so
                             the entries of a position are RUNS of equal labels in one fixed order.
                             ncb_batch.run_labels[n_runs] (host memory, whatever ncb_batch.memory says) names
                             the label of every run, and ncb_batch.label points at uint16 run_ends[P][n_runs]:
                             entry e of position p (0-based inside the position) belongs to the first run j
                             with e < run_ends[p][j]; run_ends[p] is non-decreasing and its last element is
                             the number of entries of the position.  4 bytes per entry on the link; the
                             label bytes are rebuilt on the device (one small kernel) before the statistics
                             kernel.  ncb_pack_label_runs writes run_ends from per-entry labels               */
};
#define NCB_MAX_RUNS 16

/* per-position status bits; the row is dropped by the host mirror when
 * (status & NCB_POS_DROP_MASK) != 0 */
enum {
    NCB_POS_OK = 0,
    NCB_POS_LOW_COVERAGE = 1,   /* run.py:511-515                                        */
    NCB_POS_BAD_STD = 2,        /* comparisons.py:108-120: std NaN or 0 after the filter */
    NCB_POS_TOO_MANY_READS = 4, /* more than NCB_MAX_READS entries (or than ncb_batch.n_reads): the row holds
                                   the defaults and ncb_wait returns NCB_ERR_TOO_MANY_READS */
    NCB_POS_DROP_MASK = 7,
    NCB_POS_AUTO_GMM = 256,     /* informational, with NCB_TEST_AUTO: the automatic test of
                                   this position is GMM (coverage >= 500), else KS       */
    NCB_POS_MW_EXACT_SKIPPED = 512 /* MW p-value left NaN: SciPy would enumerate the exact null
                                   distribution (a sample of <= 8 values, no ties) and the
                                   other sample is too large for the on-chip recurrence  */
};

typedef struct {
    int32_t struct_size;        /* sizeof(ncb_opts), for forward compatibility           */
    int32_t tests;              /* NCB_TEST_* mask                                       */
    int32_t n_samples;          /* S: sample ids are 0..S-1 (config.py:443-445)          */
    int32_t depleted_condition; /* condition id of `depleted_condition` (comparisons.py:586-587) */
    int32_t min_coverage;       /* per condition, on valid intensities (run.py:511-515); 0 = off */
    int32_t dwell_is_raw;       /* 1: apply log10(dwell + 1e-10) (run.py:573); 0: already done */
    int32_t cluster_counts;     /* NCB_COUNTS_*                                          */
    int32_t em_fp64;            /* 1 (default): EM runs in fp64 on the float32 standardised values -- the
                                   definition in oracle/gmm.py, bit-exact counts.  0: the per-read E-step
                                   (densities, responsibilities, log-likelihood terms) runs in float32 like
                                   the reference's fit (dtype=torch.float32, comparisons.py:33, 345), the
                                   sufficient statistics are still accumulated in fp64 and the M-step is
                                   fp64: faster, and within float32 rounding of the definition -- a read
                                   whose two densities are within ~1e-6 of each other, or a stop test
                                   within ~1e-7 of em_tol, can fall the other way (mixture
                                   kernel: measured flip rate)                                          */
    int32_t em_max_iter;        /* 100                                                   */
    int32_t diagnostics;        /* 1: fill ncb_results.diag_* (integer statistics, GMM parameters) */
    double em_tol;              /* 1e-3                                                  */
    double reg_covar;           /* 1e-6                                                  */
    int32_t input_standardised; /* 1: the values are already filtered and standardised (what
                                   compare_transcript hands to _nonparametric_test / _gmm_test_split,
                                   comparisons.py:95-120, 218, 338): no outlier rule, no scaling, no
                                   bad-std drop; the shift statistics are those of the values given */
    int32_t reserved;           /* 0                                                      */
} ncb_opts;

/* label byte of a read: bit 0 = condition id (0/1), bits 1..7 = sample id */
#define NCB_LABEL(cond, sample) ((uint8_t)(((sample) << 1) | ((cond) & 1)))

typedef struct {
    int32_t layout;             /* NCB_LAYOUT_*                                          */
    int32_t memory;             /* NCB_MEM_*: where signal/label/pos_offsets live        */
    int64_t n_positions;        /* P                                                     */
    int64_t n_entries;          /* CSR: E = pos_offsets[P]; dense: P*R                   */
    const int64_t *pos_offsets; /* CSR: [P+1]; dense: NULL                               */
    const void *signal;         /* CSR: float32[E][V] {intensity, dwell[, motor dwell]}; CSR_I16: int16[E][V];
                                   dense: float32[P][R][V]                                */
    const uint8_t *label;       /* CSR: [E]; dense: [R]                                  */
    int32_t n_reads;            /* dense: R.  CSR: an upper bound on the entries of a position if the
                                   packer knows one (ncb_pack_* callers do), 0 = unknown.  The kernels of
                                   the size classes above the bound are not launched; a position that
                                   exceeds it gets NCB_POS_TOO_MANY_READS                                */
    int32_t n_vars;             /* V: 2, or 3 with the motor-dwell column of
                                   Worker._prepare_data (run.py:575-584); CSR: 0 means 2  */
    const uint8_t *run_labels;  /* NCB_LAYOUT_CSR_I16_RUNS: [n_runs] label of every run, HOST memory; else NULL */
    int32_t n_runs;             /* NCB_LAYOUT_CSR_I16_RUNS: 1..NCB_MAX_RUNS; else 0       */
    int32_t reserved;           /* 0                                                      */
} ncb_batch;

/* rows of ncb_results.pvalues ([NCB_PV_ROWS][P], float64, NaN = not tested) */
enum {
    NCB_PV_KS_INTENSITY = 0, NCB_PV_KS_DWELL = 1,
    NCB_PV_MW_INTENSITY = 2, NCB_PV_MW_DWELL = 3,
    NCB_PV_TT_INTENSITY = 4, NCB_PV_TT_DWELL = 5,
    NCB_PV_GMM_CHI2 = 6,
    NCB_PV_GMM_LOGIT = 7, NCB_PV_GMM_LOGIT_COEF = 8,
    NCB_PV_ROWS = 9
};

/* rows of ncb_results.stats ([NCB_ST_ROWS][P], float32): the 12 shift statistics in the
 * reference's column order (comparisons.py:425-433), then GMM_LOR */
enum {
    NCB_ST_C1_MEAN_INTENSITY = 0, NCB_ST_C1_MEAN_DWELL = 1,
    NCB_ST_C1_MEDIAN_INTENSITY = 2, NCB_ST_C1_MEDIAN_DWELL = 3,
    NCB_ST_C1_SD_INTENSITY = 4, NCB_ST_C1_SD_DWELL = 5,
    NCB_ST_C2_MEAN_INTENSITY = 6, NCB_ST_C2_MEAN_DWELL = 7,
    NCB_ST_C2_MEDIAN_INTENSITY = 8, NCB_ST_C2_MEDIAN_DWELL = 9,
    NCB_ST_C2_SD_INTENSITY = 10, NCB_ST_C2_SD_DWELL = 11,
    NCB_ST_GMM_LOR = 12,
    /* motor-dwell shift statistics, written only for 3-variable batches (motor_dwell_offset > 0,
     * comparisons.py:411-433): c1/c2 x mean/median/sd of the third variable */
    NCB_ST_C1_MEAN_MOTOR = 13, NCB_ST_C1_MEDIAN_MOTOR = 14, NCB_ST_C1_SD_MOTOR = 15,
    NCB_ST_C2_MEAN_MOTOR = 16, NCB_ST_C2_MEDIAN_MOTOR = 17, NCB_ST_C2_SD_MOTOR = 18,
    NCB_ST_ROWS = 19
};

/* rows of ncb_results.diag_i64 ([NCB_DI_ROWS][P]) -- exact integer statistics */
enum {
    NCB_DI_N0_INTENSITY = 0, NCB_DI_N1_INTENSITY = 1,   /* valid reads after the outlier filter */
    NCB_DI_N0_DWELL = 2, NCB_DI_N1_DWELL = 3,
    NCB_DI_KS_H_INTENSITY = 4, NCB_DI_KS_H_DWELL = 5,   /* h = max|c0*n1/g - c1*n0/g|            */
    NCB_DI_MW_U2_INTENSITY = 6, NCB_DI_MW_U2_DWELL = 7, /* 2*U1 (average ranks)                  */
    NCB_DI_MW_TIE_INTENSITY = 8, NCB_DI_MW_TIE_DWELL = 9, /* sum t^3 - t                          */
    NCB_DI_CONT_00 = 10, NCB_DI_CONT_01 = 11, NCB_DI_CONT_10 = 12, NCB_DI_CONT_11 = 13,
    NCB_DI_EM_ITERS = 14, NCB_DI_N_GMM = 15, NCB_DI_N_OUTLIERS = 16,
    NCB_DI_ROWS = 17
};

/* rows of ncb_results.diag_f64 ([NCB_DF_ROWS][P]) -- GMM fit */
enum {
    NCB_DF_BIC1 = 0, NCB_DF_BIC2 = 1,
    NCB_DF_W0 = 2,
    NCB_DF_MU0_X = 3, NCB_DF_MU0_Y = 4, NCB_DF_C0_XX = 5, NCB_DF_C0_XY = 6, NCB_DF_C0_YY = 7,
    NCB_DF_MU1_X = 8, NCB_DF_MU1_Y = 9, NCB_DF_C1_XX = 10, NCB_DF_C1_XY = 11, NCB_DF_C1_YY = 12,
    NCB_DF_MEAN2_X = 13, NCB_DF_MEAN2_Y = 14, NCB_DF_STD2_X = 15, NCB_DF_STD2_Y = 16,
    NCB_DF_ROWS = 17
};

/* Host results (memory == NCB_MEM_HOST): only the rows the configured tests write are copied back --
 * status; the p-value rows of the tests in ncb_opts.tests (GMM_LOGIT rows with NCB_TEST_LOGIT); stats rows
 * 0..11 (+ GMM_LOR with a mixture test, + the motor rows of a 3-variable batch); counts with a mixture test
 * and cluster_counts != NCB_COUNTS_NONE.  Rows of tests that were not asked for are left as the caller's
 * buffer had them (fill it with NaN once to read them as "not tested"); ncb_results_host_bytes() gives the
 * bytes a call copies.  Device results (NCB_MEM_DEVICE): every row of every array is written (NaN / 0 where
 * a test did not run). */
typedef struct {
    int32_t memory;        /* NCB_MEM_*: where the arrays below live                      */
    int32_t reserved;
    int32_t *status;       /* [P]                 NCB_POS_* bits                          */
    double *pvalues;       /* [NCB_PV_ROWS][P]                                            */
    float *stats;          /* [NCB_ST_ROWS][P]                                            */
    uint32_t *counts;      /* [S][2][P] hard: {<sample>_mod, <sample>_unmod} as uint32
                              (comparisons.py:513-514); soft: the same slots hold float32 */
    int64_t *diag_i64;     /* [NCB_DI_ROWS][P] or NULL                                    */
    double *diag_f64;      /* [NCB_DF_ROWS][P] or NULL                                    */
} ncb_results;

/* --- lifecycle ------------------------------------------------------------------ */

int ncb_abi_version(void);
/* sizeof(ncb_opts), sizeof(ncb_batch), sizeof(ncb_results) as this library was built: a binding checks its own
 * struct definitions against them (ncb_default_opts writes a whole ncb_opts). */
int ncb_sizeof_opts(void);
int ncb_sizeof_batch(void);
int ncb_sizeof_results(void);

/* Fills *opts with the reference's defaults (config.py:196-214): tests = GMM|KS,
 * n_samples = 4, min_coverage = 30, hard counts, 100 EM iterations, tol 1e-3, reg_covar 1e-6. */
void ncb_default_opts(ncb_opts *opts);

/* Creates a handle on CUDA device `device`.  Fails (NCB_ERR_CUDA) if there is no
 * usable GPU: there is no CPU fallback. */
int ncb_create(ncb_handle **out, int device, const ncb_opts *opts);
int ncb_set_opts(ncb_handle *h, const ncb_opts *opts);
void ncb_destroy(ncb_handle *h);
const char *ncb_last_error(const ncb_handle *h);

/* --- the hot path ----------------------------------------------------------------
 * ncb_compare: replaces the body of TranscriptComparator.compare_transcript
 * (comparisons.py:83-137) plus Worker._filter_low_cov_positions / the log10 of
 * _prepare_data (run.py:511-515, 573) for a batch of positions.
 *
 * `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  With
 * NCB_MEM_HOST buffers the call copies inputs host->device and results device->host
 * on that stream (use pinned memory for the copies to be asynchronous) and returns
 * after enqueueing; ncb_wait() blocks until the results are in place.  With
 * NCB_MEM_DEVICE buffers nothing is copied and the kernels are only enqueued.        */
int ncb_compare(ncb_handle *h, const ncb_batch *batch, ncb_results *results, void *stream);
int ncb_wait(ncb_handle *h, void *stream);
/* Bytes ncb_compare copies device->host for a batch of P positions with host results (see ncb_results);
 * n_vars as in ncb_batch, with_counts = results->counts != NULL. */
int64_t ncb_results_host_bytes(const ncb_handle *h, int64_t n_positions, int32_t n_vars, int with_counts);

/* --- unit entry points (each is one stage of ncb_compare, exposed for parity tests) */

/* Exact two-sided two-sample KS p-values from integer statistics: replaces
 * scipy.stats._stats_py._attempt_exact_2kssamp (scipy/stats/_stats_py.py:7825-7870),
 * reached from comparisons.py:222 via scipy.stats.mstats.ks_twosamp.  n0,n1,h: [n].  */
int ncb_ks_exact_pvalues(ncb_handle *h, int64_t n, const int32_t *n0, const int32_t *n1,
                         const int64_t *hstat, double *p_out, int memory, void *stream);

/* Benjamini-Hochberg q-values ignoring NaNs: replaces
 * Postprocessor._multipletests_filter_nan (postprocessing.py:255-285).
 * scale multiplies every q (postprocessing.py:232-235 uses N/n for auto_qvalue); 1.0 otherwise. */
int ncb_bh_fdr(ncb_handle *h, int64_t n, const double *pvalues, double *qvalues, double scale,
               int memory, void *stream);

/* The same correction for ONE VALUE RANGE of a column whose rows are spread over several GPUs (the
 * reference corrects a column over the rows of the whole run, postprocessing.py:209-285): the column
 * holds m_total numbers (NaNs not counted) of which n_above are larger than every number passed
 * here.  q[i] = min over the passed values p_j >= p_i of p_j * m_total / rank_j with rank_j the rank of
 * p_j in the whole column; the caller completes it with the minimum of the ranges above (one number
 * per range) -- nanocompore_b200/sharding.py:fdr_columns.  m_total < 0: the values are the whole
 * column (= ncb_bh_fdr). */
int ncb_bh_fdr_segment(ncb_handle *h, int64_t n, const double *pvalues, double *qvalues, double scale,
                       int64_t n_above, int64_t m_total, int memory, void *stream);

/* Sequence-context combination of p-value columns (Hou's weighted Fisher combination of
 * dependent p-values): replaces TranscriptComparator._combine_context_pvalues
 * (comparisons.py:436-474) with cross_corr_matrix (627-651) and combine_pvalues_hou (659-705),
 * for n_tx transcripts and n_cols p-value columns in one launch.
 *   tx_offsets  HOST int64[n_tx + 1]   rows of transcript t: tx_offsets[t] .. tx_offsets[t + 1]
 *   pos         HOST int64[n]          reference position of every row (distinct within a transcript)
 *   pvalues     float64[n_cols][n]     NaN = position not tested by that column   (`memory`)
 *   weights     HOST float64[2 * context + 1]   (harmonic or uniform, comparisons.py:439-443)
 *   combined    float64[n_cols][n]     NaN where the row's own p-value is NaN     (`memory`)
 *   status      HOST int32[n_cols][n_tx]   NCB_CTX_*: where the reference raises, the status says
 *                                       why and the rows of that (column, transcript) are NaN
 * 1 <= context <= NCB_CTX_MAX.  The call returns when the results are in place. */
#define NCB_CTX_MAX 8
enum {
    NCB_CTX_OK = 0,
    NCB_CTX_TOO_SHORT = 1,       /* "Not enough p-values for a context of order" (comparisons.py:632-633) */
    NCB_CTX_INVALID_PVALUE = 2   /* "At least one p-value is invalid" (comparisons.py:636-637)          */
};
int ncb_context_combine(ncb_handle *h, int64_t n_tx, const int64_t *tx_offsets, const int64_t *pos,
                        int32_t n_cols, const double *pvalues, int32_t context, const double *weights,
                        double *combined, int32_t *status, int memory, void *stream);

/* --- host-side CSR packer (no GPU work) ---------------------------------------------
 * Dense per-transcript tensor of the reference's readers (run.py:621-855: data[P][R][V] float32,
 * NaN = the read does not cover the position) -> the NCB_LAYOUT_CSR arrays.  Two passes so that
 * many transcripts can be packed into one batch: ncb_pack_count gives the reads per position,
 * the caller turns the counts of all transcripts into pos_offsets (running sum), ncb_pack_fill
 * writes signal/label at those offsets (pos_offsets points at this transcript's first
 * position; the values are offsets into the whole batch).  A read is kept at a position when
 * its intensity or its dwell is a number; read order is preserved.  `threads` host threads. */
int ncb_pack_count(const float *data, int64_t n_positions, int32_t n_reads, int32_t n_vars,
                   int64_t *counts, int threads);
int ncb_pack_fill(const float *data, const uint8_t *label, int64_t n_positions, int32_t n_reads,
                  int32_t n_vars, const int64_t *pos_offsets, float *signal_out, uint8_t *label_out,
                  int threads);
/* Same, writing NCB_LAYOUT_CSR_I16 entries (int16[E][2]).  Every number of the tensor must be an integer in
 * [-32767, 32767] (Uncalled4 input: the float32 tensor holds int16 tag values, uncalled4.py:118-192); NaN
 * becomes -32768.  Returns NCB_ERR_INVALID, with nothing usable written, when a value is not representable
 * (the caller then packs float32 entries with ncb_pack_fill). */
int ncb_pack_fill_i16(const float *data, const uint8_t *label, int64_t n_positions, int32_t n_reads,
                      int32_t n_vars, const int64_t *pos_offsets, int16_t *signal_out, uint8_t *label_out,
                      int threads);

/* --- transcript readers (host code) ---------------------------------------
 * Planar rows: one read = intensity[n_positions] + dwell[n_positions] float32, NaN where the
 * read does not cover the position.  This is what the reference's SQLite store keeps per read
 * (signal_data blobs, database.py:69-76, decoded at run.py:795-801) and what
 * ncb_bam_read_uncalled4 decodes from Uncalled4 BAM tags; ncb_pack_rows_* turn the rows of all
 * samples of a transcript into the CSR arrays of ncb_compare without building the dense
 * (positions, reads, vars) tensor of GenericWorker._read_data / Uncalled4.get_data
 * (run.py:833-838, uncalled4.py:80-115).
 *
 * ncb_pack_rows_count / ncb_pack_rows_fill: rows are passed as arrays of row pointers (so SQLite
 * blobs are used in place, and ordering / filtering reads is a permutation of pointers), in the
 * read order the fit must see (run.py:698-700, 820-826).  motor_offset > 0 adds the third
 * variable of Worker._prepare_data (run.py:575-584) -- the raw dwell of the same read
 * motor_offset positions downstream -- and the output has 3 floats per entry.  A read is kept at
 * a position when any of its variables is a number.  Two passes like ncb_pack_count/_fill. */
int ncb_pack_rows_count(const float *const *intensity, const float *const *dwell, int64_t n_rows,
                        int64_t n_positions, int32_t motor_offset, int64_t *counts, int threads);
int ncb_pack_rows_fill(const float *const *intensity, const float *const *dwell,
                       const uint8_t *row_label, int64_t n_rows, int64_t n_positions,
                       int32_t motor_offset, const int64_t *pos_offsets, float *signal_out,
                       uint8_t *label_out, int threads);
/* ncb_pack_rows_fill writing NCB_LAYOUT_CSR_I16 entries (int16[E][2 or 3]); representability rule and
 * return value as ncb_pack_fill_i16. */
int ncb_pack_rows_fill_i16(const float *const *intensity, const float *const *dwell,
                           const uint8_t *row_label, int64_t n_rows, int64_t n_positions,
                           int32_t motor_offset, const int64_t *pos_offsets, int16_t *signal_out,
                           uint8_t *label_out, int threads);
/* Label runs of a packed CSR batch (NCB_LAYOUT_CSR_I16_RUNS): run_ends[p][j] = the number of entries of
 * position p whose label is one of run_labels[0..j].  NCB_ERR_INVALID when a label is not in run_labels, when
 * the labels of a position do not come in the order of run_labels (the batch then travels with its label
 * bytes), when a position has more than 65535 entries, or n_runs is not in 1..NCB_MAX_RUNS. */
int ncb_pack_label_runs(const int64_t *pos_offsets, const uint8_t *label, int64_t n_positions,
                        const uint8_t *run_labels, int32_t n_runs, uint16_t *run_ends, int threads);
/* common.get_reads_invalid_ratio (common.py:298-329): per row, the share of positions between
 * its first and last valid intensity that are NaN; NaN for a row without any valid position. */
int ncb_rows_invalid_ratio(const float *const *intensity, int64_t n_rows, int64_t n_positions,
                           double *ratios, int threads);

/* BGZF/BAM reader with the Uncalled4 tag decoder: replaces pysam.AlignmentFile.fetch/count and
 * Uncalled4.get_data/_copy_signal_to_tensor/_resolve_skips (uncalled4.py:57-201) for one
 * reference (transcript).  ncb_bam_open indexes the file in one pass (`threads` inflate BGZF
 * blocks in parallel); no .bai is needed.  On failure ncb_bam_open may still return a handle so
 * that ncb_bam_last_error can be read; close it.  A corrupt or truncated file is an error
 * (NCB_ERR_INVALID + message), never a crash: every length field is checked before it sizes a
 * buffer, and no C++ exception leaves the library (NCB_ERR_NOMEM when memory runs out). */
typedef struct ncb_bam ncb_bam;
int ncb_bam_open(ncb_bam **out, const char *path, int threads);
void ncb_bam_close(ncb_bam *b);
const char *ncb_bam_last_error(const ncb_bam *b);
int32_t ncb_bam_n_references(const ncb_bam *b);
const char *ncb_bam_reference_name(const ncb_bam *b, int32_t i);
int64_t ncb_bam_reference_length(const ncb_bam *b, int32_t i);
/* every record placed on the reference, secondary and supplementary included
 * (pysam's count(reference=...), uncalled4.py:76-77): the capacity to allocate */
int64_t ncb_bam_reference_records(const ncb_bam *b, int32_t i);
int32_t ncb_bam_reference_index(const ncb_bam *b, const char *name);   /* -1: not in the header */
/* Decodes the primary reads of reference `ref`: read k fills intensity[k*ref_len ..] and
 * dwell[k*ref_len ..] (NaN elsewhere; the int16 value -32768 is Uncalled4's NaN) and
 * names[k*name_stride ..] (NUL-terminated query name).  kit_len / kit_center: the pore model
 * (common.py:29-32: RNA002 5/4, RNA004 9/5).  Returns the number of reads decoded (file order),
 * or a negative ncb_status (message in ncb_bam_last_error). */
int64_t ncb_bam_read_uncalled4(ncb_bam *b, int32_t ref, int64_t ref_len, int32_t kit_len,
                               int32_t kit_center, float *intensity, float *dwell, char *names,
                               int32_t name_stride, int64_t max_reads);

/* --- device-side Uncalled4 decoder (csrc/decode.cu) -----------------------------------------
 * The same step as ncb_bam_read_uncalled4 + ncb_pack_rows_* for a whole batch of transcripts, with the
 * work on the GPU: the host looks the BGZF members of the transcripts up in the handles' indexes and
 * copies their COMPRESSED bytes into pinned memory (ncb_decoder_stage, no GPU work, may run on another
 * thread while the device works on the batch of another slot); ncb_decoder_launch enqueues the upload
 * and the kernels -- inflate, record walk, tag decoding (uncalled4.py:118-201), invalid-kmer filter
 * (common.py:298-329, run.py:711-718), read order (run.py:698-700), CSR assembly -- and fills `batch`
 * with a DEVICE-resident NCB_LAYOUT_CSR_I16 batch that ncb_compare takes on the same stream (the
 * pointers belong to the slot and stay valid until its next ncb_decoder_stage).  Positions of the batch:
 * transcript after transcript, ref_len[t] positions each.
 *
 *   file_label  NCB_LABEL(condition, sample) of every file; files in the order of the config's `data`
 *               section (equal read names keep that order, readers.Uncalled4Reader)
 *   refs        [n_tx][n_files] reference index of the transcript in the file (ncb_bam_reference_index),
 *               < 0: none
 *
 * ncb_decoder_collect waits for the slot's kernels and reports per transcript the reads kept and a flag:
 * 0 = decoded; NCB_DECODE_ERROR = a member does not inflate or a record / tag is malformed;
 * NCB_DECODE_NEEDS_HOST = well-formed input this path does not reproduce to the bit (values that are no
 * int16 codes, overlapping or more than 48 ur segments, non-ASCII read names).  A flagged transcript has
 * no entries in the batch: the caller reads it with ncb_bam_read_uncalled4, which also owns the error
 * messages.  Downsampling (run.py:518-529) is not done here: transcripts with more records than
 * `downsample_high_coverage` belong to the host path as well. */
typedef struct ncb_decoder ncb_decoder;
enum { NCB_DECODE_OK = 0, NCB_DECODE_ERROR = 1, NCB_DECODE_NEEDS_HOST = 2 };
int ncb_decoder_create(ncb_decoder **out, int device, int n_slots);
void ncb_decoder_destroy(ncb_decoder *d);
const char *ncb_decoder_last_error(const ncb_decoder *d);
int ncb_decoder_stage(ncb_decoder *d, int slot, int32_t n_files, ncb_bam *const *bams, const uint8_t *file_label,
                      int32_t n_tx, const int32_t *refs, const int64_t *ref_len, int threads);
int ncb_decoder_launch(ncb_decoder *d, int slot, int32_t kit_len, int32_t kit_center, int32_t motor_offset,
                       double max_invalid_ratio, ncb_batch *batch, void *stream);
int ncb_decoder_collect(ncb_decoder *d, int slot, int32_t *tx_flags, int32_t *tx_reads, int64_t *n_entries);
/* Compressed bytes staged in the slot (what ncb_decoder_launch uploads); optionally the inflated bytes,
 * BGZF members and records behind them.  -1: nothing staged. */
int64_t ncb_decoder_staged_bytes(const ncb_decoder *d, int slot, int64_t *text_bytes, int64_t *n_members,
                                 int64_t *n_records);
int64_t ncb_decoder_kernel_launches(const ncb_decoder *d, int slot);
/* The slot's batch copied to host arrays (tests, debugging): waits for the slot's kernels.  pos_offsets
 * [P + 1]; signal int16 [E][n_vars] and label [E] with E = pos_offsets[P] <= cap_entries. */
int ncb_decoder_download(ncb_decoder *d, int slot, int64_t *pos_offsets, int16_t *signal, uint8_t *label,
                         int64_t cap_entries, int32_t n_vars);

/* --- introspection ---------------------------------------------------------------- */

/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
int64_t ncb_kernel_launches(const ncb_handle *h);
/* Device milliseconds of the position kernels / KS p-value kernels of the last ncb_compare
 * whose batch was timed (set NCB_TIMING=1 in the environment or call ncb_set_timing). */
int ncb_set_timing(ncb_handle *h, int enabled);
int ncb_last_timing(ncb_handle *h, float *ms_position_kernels, float *ms_ks_kernels, float *ms_total);
/* Sums over every ncb_compare since the last ncb_set_timing(h, 1): number of batches and the
 * device milliseconds (CUDA events on the caller's stream) of the staging copy + classification,
 * of the statistics kernels, of the mixture kernels and of the KS p-value kernels.  Blocks until
 * those batches are done. */
int ncb_timing_totals(ncb_handle *h, int64_t *n_batches, double *ms_classify, double *ms_stats,
                      double *ms_gmm, double *ms_ks);
/* The mixture share of ncb_timing_totals by kernel, over the same batches: initialisation (K = 1 fit,
 * 2-means), EM passes, final pass + tests of the split mixture kernels of the warp size classes
 * (csrc/gmm_split.cu); the kernels of larger classes count with the final pass. */
int ncb_timing_mixture(ncb_handle *h, double *ms_init, double *ms_em, double *ms_final);

/* Non-tensor fp64 rate of the device in TFLOP/s, measured with a register-resident DFMA
 * microbenchmark (best of 3 after a warm-up): the compute roofline the fp64 mixture kernel is
 * reported against in bench.py. */
int ncb_probe_fp64_peak(ncb_handle *h, double *tflops, void *stream);

/* Task counter shared by the worker processes of one box: the reference hands transcripts to its workers
 * through a multiprocessing queue (run.py:104-131, 380-420), so a worker that is served faster by the host takes
 * more of them; with one process per GPU the same is a 64-bit counter in POSIX shared memory that every rank
 * advances with an atomic fetch-add (list scheduling over a task list every rank knows: sharding.TaskQueue).
 * ncb_counter_open maps the counter `name` ("/..."; `create` != 0: creates it if need be and sets it to 0);
 * ncb_counter_add returns the value BEFORE the addition; ncb_counter_close unmaps it and, with `unlink` != 0,
 * removes the name.  Host only. */
typedef struct ncb_counter ncb_counter;
int ncb_counter_open(const char *name, int create, ncb_counter **out);
int64_t ncb_counter_add(ncb_counter *c, int64_t n);
void ncb_counter_set(ncb_counter *c, int64_t value);
void ncb_counter_close(ncb_counter *c, int unlink);

#ifdef __cplusplus
}
#endif
#endif /* NCB_H */

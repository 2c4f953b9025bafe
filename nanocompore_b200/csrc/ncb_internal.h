// Internal declarations shared by the libncb translation units (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ncb.h"

// Everything the device code needs to know about one ncb_compare call.
struct NcbDeviceArgs {
    // input
    const float *signal;         // CSR [E][2] or dense [P][R][V]
    const uint8_t *label;        // CSR [E] or dense [R]
    const int64_t *pos_offsets;  // CSR [P+1] or nullptr (dense)
    int64_t n_positions;
    int32_t n_reads;             // dense R
    int32_t n_vars;              // floats per entry (2 for CSR, V for dense)
    int32_t max_entries;         // upper bound on the entries of a position (dense: R), 0 = unknown
    int32_t signal_i16;          // 1: `signal` holds int16 codes (NCB_LAYOUT_CSR_I16), value = (float)code
    int32_t sort32;              // 1: the warp classes of an int16 batch sort 32-bit keys (NCB_SORT32=0: A/B runs)
    int32_t *error_flag;         // mapped host word: NCB_POS_TOO_MANY_READS is OR-ed in when a position is rejected
    // options
    int32_t tests;
    int32_t n_samples;
    int32_t depleted_condition;
    int32_t min_coverage;
    int32_t dwell_is_raw;
    int32_t cluster_counts;
    int32_t em_max_iter;
    int32_t em_fp32;             // 1: float32 E-step (ncb_opts.em_fp64 == 0)
    int32_t input_standardised;  // 1: no outlier rule, no scaling (ncb_opts.input_standardised)
    double em_tol;
    double reg_covar;
    // output
    int32_t *status;
    double *pvalues;
    float *stats;
    uint32_t *counts;
    int64_t *diag_i64;
    double *diag_f64;
    // scratch: hand-over from the statistics kernel to the mixture kernel
    float2 *zbuf;                // standardised (intensity, dwell) per entry, NaN x = not in the fit
    int32_t *gmm_n;              // [P] reads in the 2-variable fit, 0 = none on this position
    double *gmm_rec;             // [P][NCB_GMM_REC] hand-over between the three mixture kernels of the warp
                                 // size classes (gmm_split.cu); NULL: the fused kernel takes those classes too
    float *zbuf3;                // 3-variable batches: standardised motor dwell per entry
    int32_t *gmm_n3;             // 3-variable batches: [P] reads in the 3-variable fit, 0 = none
    // scratch: rows of the exact Mann-Whitney counting recurrence that do not fit on chip, one per
    // resident group (4 CAP + 8 doubles of the group's size class); NULL when the batch cannot need it
    double *mw_scratch;
    int64_t mw_scratch_len;      // doubles
    // scratch: exact-KS work items, one per (variable, position): index v * P + p
    int32_t *ks_n0;
    int32_t *ks_n1;
    int64_t *ks_h;
};

// size classes of the position kernel, by entries per position
//   class 0..2: one warp per position, 4 / 8 / 16 entries per lane
//   class 3   : one 128-thread CTA per position, 8 entries per thread (a 256-thread CTA spends the EM
//               iteration of a 600-read position in its block-wide reduction: 4.4x the cost of 300 reads)
//   class 4   : one 256-thread CTA per position, 8 entries per thread
//   class 5   : one 512-thread CTA per position, 8 entries per thread
//   class 6   : one 1024-thread CTA per position, 10 entries per thread
// doubles of the Mann-Whitney scratch per SM: (resident groups per SM) x (4 CAP + 8) is 8 320, 24 768, 32 896,
// 32 832, 32 800, 32 784 and 40 968 for the seven size classes (the kernels of one handle's classes run one
// after the other on its stream, so they share the rows); a group whose row would not fit keeps the flag
#define NCB_MW_SCRATCH_PER_SM (4 * (int64_t)NCB_MAX_READS + 8)
#define NCB_TIMING_LOG_MAX 4096   /* batches kept in a handle's timing log */
#define NCB_TIMING_EVENTS 7       /* start, classify, statistics, mixture, exact KS; mixture: after init, after EM */
// phases (each has its own work counter per class): statistics / mixture, 2 / 3 variables; the mixture
// phase of the warp classes is three kernels (gmm_split.cu): initialisation (the counter of
// NCB_PHASE_GMM), EM passes, final pass
enum { NCB_PHASE_STATS = 0, NCB_PHASE_GMM = 1, NCB_PHASE_STATS3 = 2, NCB_PHASE_GMM3 = 3, NCB_PHASE_GMM_EM = 4,
       NCB_PHASE_GMM_FINAL = 5 };
enum { NCB_NUM_CLASSES = 7, NCB_NUM_PHASES = 6 };
#define NCB_GMM_REC 24   /* doubles per position in NcbDeviceArgs.gmm_rec */
#define NCB_WL_COUNTERS (NCB_NUM_CLASSES * (1 + NCB_NUM_PHASES))
static const int NCB_CLASS_CAP[NCB_NUM_CLASSES] = {128, 256, 512, 1024, 2048, 4096, NCB_MAX_READS};

// worklists: class_count[c] positions in class_list[c * P ...]
struct NcbWorklists {
    int32_t *class_count;        // [NCB_NUM_CLASSES] device, followed by the work counters of the
                                 // position kernels: [NCB_NUM_PHASES][NCB_NUM_CLASSES]
    int32_t *class_list;         // [NCB_NUM_CLASSES][P] device
};

// exact-KS p-value worklists: problems are binned by cost so that the threads of a warp
// (thread-per-problem kernels) and consecutive warps (warp-per-problem kernel) do similar work
#define NCB_KS_BUCKETS 64      /* cost buckets per kind (2 per octave of cost)           */
enum { NCB_KS_KIND_SMALL = 0,  /* unequal sizes, DP window <= NCB_KS_SMALL_WINDOW: thread per problem */
       NCB_KS_KIND_EQUAL = 1,  /* equal sizes: Horner product, thread per problem        */
       NCB_KS_KIND_LARGE = 2,  /* unequal sizes, band span <= 1024 columns: warp wavefront, 8 KB ring */
       NCB_KS_KIND_HUGE = 3,   /* unequal sizes, wider band: CTA lockstep wavefront over 256 rows, full row
                                  (41 KB) per CTA; it also takes the LARGE problems that would outlast the batch */
       NCB_KS_KINDS = 4 };
#define NCB_KS_SMALL_WINDOW 64
#define NCB_KS_RING_SMALL 1024   /* doubles in the ring of the LARGE kind */
#define NCB_KS_RING_FULL (NCB_MAX_READS / 2 + 1)   /* the HUGE kind holds the whole row: no wrap */
#define NCB_KS_NBINS (NCB_KS_KINDS * NCB_KS_BUCKETS)

#define NCB_KS_RTAB_SIZE (NCB_MAX_READS + 320)  /* reciprocals of i + j <= n0 + n1 <= NCB_MAX_READS; the strip form
                                                   reads up to 256 entries past the last divisor it uses */

struct NcbKsWork {
    const double *rtab; // [NCB_KS_RTAB_SIZE] correctly rounded reciprocals of the integers (per handle)
    int32_t *hist;      // [NCB_KS_NBINS + 1] counts -> exclusive offsets after the scan
    int32_t *cursor;    // [NCB_KS_NBINS]
    int32_t *list;      // [n_problems]
    uint8_t *bin;       // [n_problems] bin of each problem, 255 = answered directly
    int32_t *fetch;     // [NCB_KS_KINDS] dynamic work counters of the wavefront kernels;
                        // [NCB_KS_KINDS]: LARGE problems taken over by the CTA kernel
};

// launchers (each returns the number of kernels launched, or <0 on launch error)
int ncb_launch_classify(const NcbDeviceArgs &a, const NcbWorklists &w, cudaStream_t s);
// phases & 1: statistics kernels; phases & 2: (if a.zbuf) mixture kernels; `between` (optional) is
// recorded after the statistics kernels; `marks` (optional): two events recorded after the initialisation
// and after the EM kernel of the split mixture fit (with `between` when that fit does not run)
int ncb_launch_position_kernels(const NcbDeviceArgs &a, const NcbWorklists &w, int sm_count,
                                cudaStream_t s, cudaEvent_t between, int phases, cudaEvent_t *marks = nullptr);
int ncb_launch_ks_pvalues(int64_t n_problems, const int32_t *n0, const int32_t *n1,
                          const int64_t *h, double *p_out, const NcbKsWork &w, int sm_count,
                          cudaStream_t s);
// fills the handle's table of reciprocals (once, at ncb_create)
int ncb_launch_ks_rtab(double *rtab, cudaStream_t s);
// `above`, `m_total`: the values are one value range of a longer column (ncb_bh_fdr_segment); a whole
// column has above = 0, m_total = -1 (its own count of numbers)
int ncb_launch_bh_fdr(int64_t n, const double *p, double *q, double scale, int64_t above, int64_t m_total,
                      void *scratch, size_t scratch_bytes, cudaStream_t s);
size_t ncb_bh_fdr_scratch_bytes(int64_t n);
int ncb_launch_context(int64_t n_tx, int64_t n_rows, int64_t track_total, int32_t n_cols,
                       const int64_t *tx_offsets, const int64_t *track_offsets, const int32_t *idx,
                       const double *pvalues, double *tracks, double *combined, int32_t *status,
                       int32_t context, const double *weights, cudaStream_t s);
int ncb_configure_kernels(void);  // cudaFuncSetAttribute for large dynamic shared memory
// mixture fit of the warp size classes as three kernels (gmm_split.cu): classes 0 and 1 in one launch of
// each, class 2 (`wide`) in its own; returns the number launched
int ncb_configure_gmm_split(void);
// marks (or NULL): two events, recorded after the initialisation and after the EM kernel
int ncb_launch_gmm_split(const NcbDeviceArgs &a, const NcbWorklists &w, bool wide, int sm_count, cudaStream_t s,
                         cudaEvent_t *marks);

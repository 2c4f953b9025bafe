"""
ctypes binding of libncb.so (include/ncb.h).  This is the only way the Python host code
reaches the CUDA kernels; there is no CPU fallback -- if the library cannot be loaded, or
there is no GPU, every compute entry point raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# NCB_LIB_PATH lets a tuning run load an experimental build of the same ABI
LIB_PATH = os.environ.get('NCB_LIB_PATH') or os.path.join(HERE, 'libncb.so')

NCB_ABI_VERSION = 3
NCB_MAX_READS = 10240
NCB_MAX_SAMPLES = 64

NCB_OK, NCB_ERR_INVALID, NCB_ERR_CUDA, NCB_ERR_NOMEM, NCB_ERR_TOO_MANY_READS = 0, -1, -2, -3, -4

NCB_TEST_KS, NCB_TEST_MW, NCB_TEST_TT, NCB_TEST_GMM, NCB_TEST_LOGIT, NCB_TEST_AUTO = 1, 2, 4, 8, 16, 32
NCB_COUNTS_NONE, NCB_COUNTS_HARD, NCB_COUNTS_SOFT = 0, 1, 2
NCB_MEM_HOST, NCB_MEM_DEVICE = 0, 1
NCB_LAYOUT_CSR, NCB_LAYOUT_DENSE, NCB_LAYOUT_CSR_I16, NCB_LAYOUT_CSR_I16_RUNS = 0, 1, 2, 3
NCB_MAX_RUNS = 16

NCB_POS_LOW_COVERAGE, NCB_POS_BAD_STD, NCB_POS_TOO_MANY_READS = 1, 2, 4
NCB_POS_DROP_MASK = 7
NCB_POS_AUTO_GMM = 256
NCB_POS_MW_EXACT_SKIPPED = 512
NCB_CTX_MAX = 8
NCB_CTX_OK, NCB_CTX_TOO_SHORT, NCB_CTX_INVALID_PVALUE = 0, 1, 2
NCB_DECODE_OK, NCB_DECODE_ERROR, NCB_DECODE_NEEDS_HOST = 0, 1, 2

PV = dict(KS_INTENSITY=0, KS_DWELL=1, MW_INTENSITY=2, MW_DWELL=3, TT_INTENSITY=4, TT_DWELL=5,
          GMM_CHI2=6, GMM_LOGIT=7, GMM_LOGIT_COEF=8)
NCB_PV_ROWS = 9
ST = dict(C1_MEAN_INTENSITY=0, C1_MEAN_DWELL=1, C1_MEDIAN_INTENSITY=2, C1_MEDIAN_DWELL=3,
          C1_SD_INTENSITY=4, C1_SD_DWELL=5, C2_MEAN_INTENSITY=6, C2_MEAN_DWELL=7,
          C2_MEDIAN_INTENSITY=8, C2_MEDIAN_DWELL=9, C2_SD_INTENSITY=10, C2_SD_DWELL=11, GMM_LOR=12,
          C1_MEAN_MOTOR=13, C1_MEDIAN_MOTOR=14, C1_SD_MOTOR=15, C2_MEAN_MOTOR=16, C2_MEDIAN_MOTOR=17,
          C2_SD_MOTOR=18)
NCB_ST_ROWS = 19
DI = dict(N0_INTENSITY=0, N1_INTENSITY=1, N0_DWELL=2, N1_DWELL=3, KS_H_INTENSITY=4, KS_H_DWELL=5,
          MW_U2_INTENSITY=6, MW_U2_DWELL=7, MW_TIE_INTENSITY=8, MW_TIE_DWELL=9,
          CONT_00=10, CONT_01=11, CONT_10=12, CONT_11=13, EM_ITERS=14, N_GMM=15, N_OUTLIERS=16)
NCB_DI_ROWS = 17
DF = dict(BIC1=0, BIC2=1, W0=2, MU0_X=3, MU0_Y=4, C0_XX=5, C0_XY=6, C0_YY=7,
          MU1_X=8, MU1_Y=9, C1_XX=10, C1_XY=11, C1_YY=12,
          MEAN2_X=13, MEAN2_Y=14, STD2_X=15, STD2_Y=16)
NCB_DF_ROWS = 17

EXPORTS = ['ncb_abi_version', 'ncb_sizeof_opts', 'ncb_sizeof_batch', 'ncb_sizeof_results', 'ncb_default_opts', 'ncb_create', 'ncb_set_opts', 'ncb_destroy',
           'ncb_last_error', 'ncb_compare', 'ncb_wait', 'ncb_results_host_bytes',
           'ncb_pack_fill_i16', 'ncb_pack_rows_fill_i16', 'ncb_pack_label_runs', 'ncb_ks_exact_pvalues', 'ncb_bh_fdr', 'ncb_bh_fdr_segment',
           'ncb_context_combine',
           'ncb_kernel_launches', 'ncb_set_timing', 'ncb_last_timing', 'ncb_timing_totals', 'ncb_timing_mixture',
           'ncb_pack_count', 'ncb_pack_fill', 'ncb_pack_rows_count', 'ncb_pack_rows_fill',
           'ncb_rows_invalid_ratio', 'ncb_bam_open', 'ncb_bam_close', 'ncb_bam_last_error',
           'ncb_bam_n_references', 'ncb_bam_reference_name', 'ncb_bam_reference_length',
           'ncb_bam_reference_records', 'ncb_bam_reference_index', 'ncb_bam_read_uncalled4',
           'ncb_probe_fp64_peak',
           'ncb_decoder_create', 'ncb_decoder_destroy', 'ncb_decoder_last_error', 'ncb_decoder_stage',
           'ncb_decoder_launch', 'ncb_decoder_collect', 'ncb_decoder_staged_bytes', 'ncb_decoder_kernel_launches',
           'ncb_decoder_download',
           'ncb_counter_open', 'ncb_counter_add', 'ncb_counter_set', 'ncb_counter_close']


class NcbOpts(C.Structure):
    _fields_ = [('struct_size', C.c_int32), ('tests', C.c_int32), ('n_samples', C.c_int32),
                ('depleted_condition', C.c_int32), ('min_coverage', C.c_int32),
                ('dwell_is_raw', C.c_int32), ('cluster_counts', C.c_int32), ('em_fp64', C.c_int32),
                ('em_max_iter', C.c_int32), ('diagnostics', C.c_int32),
                ('em_tol', C.c_double), ('reg_covar', C.c_double),
                ('input_standardised', C.c_int32), ('reserved', C.c_int32)]


class NcbBatch(C.Structure):
    _fields_ = [('layout', C.c_int32), ('memory', C.c_int32), ('n_positions', C.c_int64),
                ('n_entries', C.c_int64), ('pos_offsets', C.c_void_p), ('signal', C.c_void_p),
                ('label', C.c_void_p), ('n_reads', C.c_int32), ('n_vars', C.c_int32),
                ('run_labels', C.c_void_p), ('n_runs', C.c_int32), ('reserved', C.c_int32)]


class NcbResults(C.Structure):
    _fields_ = [('memory', C.c_int32), ('reserved', C.c_int32), ('status', C.c_void_p),
                ('pvalues', C.c_void_p), ('stats', C.c_void_p), ('counts', C.c_void_p),
                ('diag_i64', C.c_void_p), ('diag_f64', C.c_void_p)]


class NcbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libncb error {code}: {msg}")
        self.code = code


_libs = {}


def load(path=None):
    """Loads libncb.so (built in-tree by ``python -m nanocompore_b200.build``).  ``path`` selects
    another build of the same ABI (tuning runs)."""
    path = path or LIB_PATH
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -m nanocompore_b200.build` "
            "(nanocompore_b200 has no CPU fallback)")
    lib = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.ncb_abi_version.restype = C.c_int
    lib.ncb_default_opts.argtypes = [C.POINTER(NcbOpts)]
    lib.ncb_default_opts.restype = None
    lib.ncb_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(NcbOpts)]
    lib.ncb_create.restype = C.c_int
    lib.ncb_set_opts.argtypes = [vp, C.POINTER(NcbOpts)]
    lib.ncb_set_opts.restype = C.c_int
    lib.ncb_destroy.argtypes = [vp]
    lib.ncb_destroy.restype = None
    lib.ncb_last_error.argtypes = [vp]
    lib.ncb_last_error.restype = C.c_char_p
    lib.ncb_compare.argtypes = [vp, C.POINTER(NcbBatch), C.POINTER(NcbResults), vp]
    lib.ncb_compare.restype = C.c_int
    lib.ncb_wait.argtypes = [vp, vp]
    lib.ncb_wait.restype = C.c_int
    lib.ncb_ks_exact_pvalues.argtypes = [vp, i64, vp, vp, vp, vp, C.c_int, vp]
    lib.ncb_ks_exact_pvalues.restype = C.c_int
    lib.ncb_bh_fdr.argtypes = [vp, i64, vp, vp, C.c_double, C.c_int, vp]
    lib.ncb_bh_fdr.restype = C.c_int
    lib.ncb_bh_fdr_segment.argtypes = [vp, i64, vp, vp, C.c_double, i64, i64, C.c_int, vp]
    lib.ncb_bh_fdr_segment.restype = C.c_int
    lib.ncb_context_combine.argtypes = [vp, i64, vp, vp, i32, vp, i32, vp, vp, vp, C.c_int, vp]
    lib.ncb_context_combine.restype = C.c_int
    lib.ncb_kernel_launches.argtypes = [vp]
    lib.ncb_kernel_launches.restype = i64
    lib.ncb_set_timing.argtypes = [vp, C.c_int]
    lib.ncb_set_timing.restype = C.c_int
    lib.ncb_last_timing.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.ncb_last_timing.restype = C.c_int
    lib.ncb_timing_totals.argtypes = [vp, C.POINTER(i64)] + [C.POINTER(C.c_double)] * 4
    lib.ncb_timing_totals.restype = C.c_int
    lib.ncb_timing_mixture.argtypes = [vp] + [C.POINTER(C.c_double)] * 3
    lib.ncb_timing_mixture.restype = C.c_int
    lib.ncb_pack_count.argtypes = [vp, i64, i32, i32, vp, C.c_int]
    lib.ncb_pack_count.restype = C.c_int
    lib.ncb_pack_fill.argtypes = [vp, vp, i64, i32, i32, vp, vp, vp, C.c_int]
    lib.ncb_pack_fill.restype = C.c_int
    lib.ncb_pack_fill_i16.argtypes = [vp, vp, i64, i32, i32, vp, vp, vp, C.c_int]
    lib.ncb_pack_fill_i16.restype = C.c_int
    lib.ncb_pack_rows_fill_i16.argtypes = [vp, vp, vp, i64, i64, i32, vp, vp, vp, C.c_int]
    lib.ncb_pack_rows_fill_i16.restype = C.c_int
    lib.ncb_pack_label_runs.argtypes = [vp, vp, i64, vp, i32, vp, C.c_int]
    lib.ncb_pack_label_runs.restype = C.c_int
    lib.ncb_counter_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    lib.ncb_counter_open.restype = C.c_int
    lib.ncb_counter_add.argtypes = [vp, i64]
    lib.ncb_counter_add.restype = i64
    lib.ncb_counter_set.argtypes = [vp, i64]
    lib.ncb_counter_set.restype = None
    lib.ncb_counter_close.argtypes = [vp, C.c_int]
    lib.ncb_counter_close.restype = None
    lib.ncb_results_host_bytes.argtypes = [vp, i64, i32, C.c_int]
    lib.ncb_results_host_bytes.restype = i64
    lib.ncb_pack_rows_count.argtypes = [vp, vp, i64, i64, i32, vp, C.c_int]
    lib.ncb_pack_rows_count.restype = C.c_int
    lib.ncb_pack_rows_fill.argtypes = [vp, vp, vp, i64, i64, i32, vp, vp, vp, C.c_int]
    lib.ncb_pack_rows_fill.restype = C.c_int
    lib.ncb_rows_invalid_ratio.argtypes = [vp, i64, i64, vp, C.c_int]
    lib.ncb_rows_invalid_ratio.restype = C.c_int
    lib.ncb_bam_open.argtypes = [C.POINTER(vp), C.c_char_p, C.c_int]
    lib.ncb_bam_open.restype = C.c_int
    lib.ncb_bam_close.argtypes = [vp]
    lib.ncb_bam_close.restype = None
    lib.ncb_bam_last_error.argtypes = [vp]
    lib.ncb_bam_last_error.restype = C.c_char_p
    lib.ncb_bam_n_references.argtypes = [vp]
    lib.ncb_bam_n_references.restype = i32
    lib.ncb_bam_reference_name.argtypes = [vp, i32]
    lib.ncb_bam_reference_name.restype = C.c_char_p
    lib.ncb_bam_reference_length.argtypes = [vp, i32]
    lib.ncb_bam_reference_length.restype = i64
    lib.ncb_bam_reference_records.argtypes = [vp, i32]
    lib.ncb_bam_reference_records.restype = i64
    lib.ncb_bam_reference_index.argtypes = [vp, C.c_char_p]
    lib.ncb_bam_reference_index.restype = i32
    lib.ncb_bam_read_uncalled4.argtypes = [vp, i32, i64, i32, i32, vp, vp, vp, i32, i64]
    lib.ncb_bam_read_uncalled4.restype = i64
    lib.ncb_probe_fp64_peak.argtypes = [vp, C.POINTER(C.c_double), vp]
    lib.ncb_probe_fp64_peak.restype = C.c_int
    lib.ncb_decoder_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int]
    lib.ncb_decoder_create.restype = C.c_int
    lib.ncb_decoder_destroy.argtypes = [vp]
    lib.ncb_decoder_destroy.restype = None
    lib.ncb_decoder_last_error.argtypes = [vp]
    lib.ncb_decoder_last_error.restype = C.c_char_p
    lib.ncb_decoder_stage.argtypes = [vp, C.c_int, i32, vp, vp, i32, vp, vp, C.c_int]
    lib.ncb_decoder_stage.restype = C.c_int
    lib.ncb_decoder_launch.argtypes = [vp, C.c_int, i32, i32, i32, C.c_double, C.POINTER(NcbBatch), vp]
    lib.ncb_decoder_launch.restype = C.c_int
    lib.ncb_decoder_collect.argtypes = [vp, C.c_int, vp, vp, C.POINTER(i64)]
    lib.ncb_decoder_collect.restype = C.c_int
    lib.ncb_decoder_staged_bytes.argtypes = [vp, C.c_int, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    lib.ncb_decoder_staged_bytes.restype = i64
    lib.ncb_decoder_kernel_launches.argtypes = [vp, C.c_int]
    lib.ncb_decoder_kernel_launches.restype = i64
    lib.ncb_decoder_download.argtypes = [vp, C.c_int, vp, vp, vp, i64, i32]
    lib.ncb_decoder_download.restype = C.c_int
    if lib.ncb_abi_version() != NCB_ABI_VERSION:
        raise RuntimeError("libncb.so ABI version mismatch: rebuild with python -m nanocompore_b200.build")
    for fn, struct in ((lib.ncb_sizeof_opts, NcbOpts), (lib.ncb_sizeof_batch, NcbBatch), (lib.ncb_sizeof_results, NcbResults)):
        fn.restype = C.c_int
        if fn() != C.sizeof(struct):
            raise RuntimeError(f"{path}: {struct.__name__} is {C.sizeof(struct)} bytes here, {fn()} in the library")
    _libs[path] = lib
    return lib


def default_opts():
    o = NcbOpts()
    load().ncb_default_opts(C.byref(o))
    return o

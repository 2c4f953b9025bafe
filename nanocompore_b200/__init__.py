"""
nanocompore_b200 -- B200-native implementation of nanocompore's per-position two-condition
comparison hot path (reference: tleonardi/nanocompore v2.0.0, comparisons.py / run.py).

    from nanocompore_b200.comparisons import TranscriptComparator   # drop-in mirror
    from nanocompore_b200.engine import Engine, pack_csr            # batch (CSR) entry point

Everything numeric runs in the CUDA kernels of libncb.so (nanocompore_b200/csrc, C ABI in
include/ncb.h).  There is no CPU fallback.
"""
__version__ = '0.1.0'

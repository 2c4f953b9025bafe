"""
Mirror of the p-value correction of the reference's post-processing
(/root/reference/nanocompore/postprocessing.py):

  * ``multipletests_filter_nan``   = ``Postprocessor._multipletests_filter_nan`` (255-285)
  * ``correct_pvalues``            = ``Postprocessor._correct_pvalues`` (209-252): a ``*_qvalue``
    column per ``*_pvalue`` column, and for ``auto_pvalue`` the per-test-type correction scaled by
    N / n (211-235)

The Benjamini-Hochberg correction itself runs on the GPU (libncb ``ncb_bh_fdr``: key transform,
radix sort, scan, scatter; NaNs skipped).  ``fdr_bh`` is the only ``correction_method`` the
reference's configuration schema admits (config.py:175, default config.py:214), so it is the only one
provided: ``check_correction_method`` refuses anything else, and ``config.PathConfig`` calls it when the
configuration is read -- a job does not run for hours to fail in its last step.
"""
import numpy as np

SUPPORTED_CORRECTION_METHODS = ('fdr_bh',)


def check_correction_method(method):
    if method not in SUPPORTED_CORRECTION_METHODS:
        raise ValueError(f"correction method {method!r} is not available: the configuration schema of "
                         f"nanocompore admits {SUPPORTED_CORRECTION_METHODS[0]!r} only (config.py:175)")
    return method


def multipletests_filter_nan(engine, pvalues, method='fdr_bh', scale=1.0):
    """q-values of ``pvalues`` (NaNs ignored and returned as NaN)."""
    check_correction_method(method)
    pvalues = np.asarray(pvalues, dtype=np.float64)
    if pvalues.size == 0:
        return pvalues.copy()
    return engine.bh_fdr(pvalues, scale=scale)


def correct_pvalues(engine, df, method='fdr_bh', auto_test=None):
    """Adds the q-value columns to ``df`` (in place, columns sorted like the reference does) and
    returns it.  ``auto_test``: the ``auto_test`` column (a pandas Series or array) when the frame
    has an ``auto_pvalue`` column."""
    check_correction_method(method)
    for column in list(df.columns):
        if column == 'auto_pvalue' and method == 'fdr_bh':
            pvals = np.array(df[column].values, dtype=float)
            tests = np.asarray(auto_test)
            all_qvals = np.full(pvals.shape, np.nan, dtype=float)
            for test_type in dict.fromkeys(tests.tolist()):
                mask = tests == test_type
                n = int(mask.sum())
                if n:
                    # BH inside the test type, times N / n (postprocessing.py:229-235)
                    all_qvals[mask] = multipletests_filter_nan(engine, pvals[mask], method, scale=len(pvals) / n)
            df['auto_qvalue'] = all_qvals
        elif 'pvalue' in column:
            pvals = np.array(df[column].values, dtype=float)
            df[column.replace('pvalue', 'qvalue')] = multipletests_filter_nan(engine, pvals, method)
    df.sort_index(axis=1, inplace=True)
    return df

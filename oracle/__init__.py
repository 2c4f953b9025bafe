"""
oracle/ -- CPU restatement of nanocompore's per-position comparison path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
anything from here.  ``nanocompore_b200`` never imports ``oracle`` and has no
CPU fallback: without the CUDA library it raises.

What it restates (reference = /root/reference, nanocompore v2.0.0):

* ``oracle.stats``          comparisons.py:95-120, 406-433, 622-624 (shift stats, outlier filter, standardise)
* ``oracle.prep``           run.py:511-601 (log10 dwell, coverage filter, downsample)
* ``oracle.nonparametric``  comparisons.py:218-252 + scipy ks_2samp / mannwhitneyu / ttest_ind
* ``oracle.ks_exact``       scipy/stats/_stats_py.py:7713-7870 (exact two-sample KS p-value)
* ``oracle.gmm``            the gmm_gpu.GMM contract used at comparisons.py:338-361
* ``oracle.gmm_test``       comparisons.py:338-392, 477-605, 708-744 (contingency, counts, LOR, chi2)
* ``oracle.context``        comparisons.py:436-474, 627-705 (Hou context combination)
* ``oracle.fdr``            postprocessing.py:209-285 (Benjamini-Hochberg with NaNs)
* ``oracle.comparator``     comparisons.py:41-215 (whole compare_transcript)

Parity status
-------------
PINNED (checked against the unmodified reference imported from /root/reference in
the build container -- ``oracle/refharness.py`` + ``tests/golden/make_golden.py`` --
and against the reference's own known-answer tests, see tests/test_oracle_*.py):
shift stats, outlier/standardise, KS / MW / TT p-values, contingency tables,
cluster counts, LOR, chi2, Hou combination, BH FDR.

PARITY UNPINNED: the Gaussian-mixture fit.  The reference delegates it to the
third-party package ``gmm-gpu==0.2.8`` (pyproject.toml:29), which is neither
vendored in /root/reference nor installed, and there is no network.  ``oracle.gmm``
is a *definition* following scikit-learn ``GaussianMixture(covariance_type='full')``
semantics with a deterministic initialisation; it is pinned
only by the two behavioural tests the reference has for it
(tests/test_comparisons.py:464-504) and by scikit-learn itself as a loose second
oracle.  The optional GMM-logit column has no reference code at all (v1-era only).

Numeric convention shared by this oracle and the CUDA kernels: every float32
reduction of the reference (torch ``nanmean`` / ``nansum``, whose summation order is
implementation-defined and differs between torch's own CPU and CUDA back ends) is
restated as "exact sum (float64 accumulation), rounded once to float32", and
``log10`` of the dwell as "float64 log10 rounded to float32".  All other float32
operations are mirrored one by one.
"""

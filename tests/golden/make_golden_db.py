"""
Extracts the reference's golden results database into tests/golden/golden_db.npz.

/root/reference/tests/fixtures/out_sampComp_sql.db is the result store of a real nanocompore run
on the fixture BAMs (tests/fixtures/{kd1,kd2,wt1,wt2}.bam): 3437 rows of table ``kmer_results``
with the 12 shift statistics, KS_{intensity,dwell}_pvalue, GMM_chi2_pvalue, GMM_LOR and the
per-sample cluster counts {WT1,WT2,KD1,KD2}_{mod,unmod}.  It was written by an OLDER build of the
reference (sample labels WT*/KD*, condition 1 = WT: its c1_* columns are today's c2_*; cluster
counts stored also where the BIC gate rejects the 2-component fit) with the real SciPy and the
real gmm_gpu -- the only numbers under /root/reference that gmm_gpu itself produced.

Only numbers are extracted (no reference source).  Run in the build container:

    python tests/golden/make_golden_db.py

tests/test_oracle_golden_db.py checks the oracle against the extract.
"""
import os
import sqlite3

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DB = '/root/reference/tests/fixtures/out_sampComp_sql.db'
SHIFT = [f"c{c}_{s}_{d}" for c in (1, 2) for s in ('mean', 'median', 'sd') for d in ('intensity', 'dwell')]
COUNTS = [f"{s}_{k}" for s in ('WT1', 'WT2', 'KD1', 'KD2') for k in ('mod', 'unmod')]
COLUMNS = ['pos', 'kmer'] + SHIFT + ['KS_intensity_pvalue', 'KS_dwell_pvalue', 'GMM_chi2_pvalue', 'GMM_LOR'] + COUNTS


def main():
    con = sqlite3.connect(f'file:{DB}?mode=ro', uri=True)
    out = {'columns': np.array(COLUMNS)}
    names = list(con.execute('select id, name from transcripts order by id'))
    out['transcripts'] = np.array([n for _, n in names])
    total = 0
    for i, (tid, _) in enumerate(names):
        rows = list(con.execute(f"select {','.join(COLUMNS)} from kmer_results where transcript_id = ? order by pos",
                                (tid,)))
        a = np.array([[np.nan if v is None else v for v in r] for r in rows], dtype=np.float64)
        out[f'tx{i}'] = a
        total += len(a)
    out['total_rows'] = np.array(total)
    np.savez_compressed(os.path.join(HERE, 'golden_db.npz'), **out)
    print('golden_db: rows', total)


if __name__ == '__main__':
    main()

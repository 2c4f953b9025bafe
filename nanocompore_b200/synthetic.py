"""
Seeded synthetic transcripts of the BASELINE.json shapes.

One generator feeds the oracle, the kernel parity tests and bench.py so that all
three see identical inputs.  Output is the reference reader's layout
(/root/reference/nanocompore/run.py:621-855): a dense ``(positions, reads, 2)``
float32 array, NaN = read does not cover the position, column 0 = intensity (pA),
column 1 = RAW dwell in seconds (``_prepare_data`` takes the log10 later,
run.py:573), plus ``samples`` / ``conditions`` per read in the reference's order
(sample ids follow the config order ``[kd1, kd2, wt1, wt2] -> 0..3``, config.py:443-445;
conditions are the blocks ``[0, 0, 1, 1]``, run.py:820-826).

"reads per position" is PER CONDITION (n0 = n1 = n before missingness).
"""
from dataclasses import dataclass

import numpy as np

BASE_SEED = 20251019


@dataclass
class SyntheticTranscript:
    name: str
    data: np.ndarray          # (P, R, 2) float32, NaN = missing, dwell in seconds
    samples: np.ndarray       # (R,) int16
    conditions: np.ndarray    # (R,) int16
    modified: np.ndarray      # (P,) bool -- ground truth, not used by the tests


def make_transcript(rng, n_positions, reads_per_cond, *, quantised=False, uncalled4=False,
                    mod_fraction=0.10, gap_rate=0.02, span_start_max=0.3,
                    name='tx', n_replicates=2):
    """``uncalled4``: the values are what the Uncalled4 reader hands over (uncalled4.py:118-192, the
    int16 tags ``uc`` / ``ul`` copied into the float32 tensor): the current as an integer code,
    250 codes per pA around 95 pA (the fixture BAMs hold codes of -14 000 .. 18 000 with a
    per-position spread of ~1 000), the dwell as a whole number of samples at 4 kHz (>= 1).  The same
    draws as the default generator, so a position has the same shape in both."""
    P, n = int(n_positions), int(reads_per_cond)
    R = 2 * n
    S = 2 * n_replicates
    # sample blocks: each condition's reads are split evenly over its replicates
    per = [n // n_replicates + (1 if i < n % n_replicates else 0) for i in range(n_replicates)]
    samples = np.concatenate([np.full(c, s, dtype=np.int16)
                              for s, c in enumerate(per + per)])
    conditions = (samples >= n_replicates).astype(np.int16)
    assert samples.size == R and S == 2 * n_replicates

    f32 = np.float32
    mu = rng.uniform(60.0, 130.0, size=P).astype(f32)
    sigma = rng.uniform(2.0, 6.0, size=P).astype(f32)
    inten = rng.standard_normal((P, R), dtype=f32) * sigma[:, None] + mu[:, None]
    dwell = np.exp(rng.standard_normal((P, R), dtype=f32) * f32(0.5) + f32(np.log(0.01)))

    modified = rng.random(P) < mod_fraction
    frac = rng.uniform(0.3, 0.9, size=P).astype(f32)
    hit = (rng.random((P, R), dtype=f32) < frac[:, None]) & modified[:, None] & (conditions[None, :] == 1)
    inten = np.where(hit, inten + f32(3.0) * sigma[:, None], inten)
    dwell = np.where(hit, dwell * f32(1.6), dwell)

    if uncalled4:
        inten = np.clip(np.round((inten - f32(95.0)) * f32(250.0)), f32(-32767.0), f32(32767.0))
        dwell = np.clip(np.round(dwell * f32(4000.0)), f32(1.0), f32(32767.0))
    elif quantised:
        # mimics Uncalled4 int16 current levels / integer sample counts: true ties
        inten = np.round(inten * f32(64.0)) / f32(64.0)
        dwell = np.maximum(np.round(dwell * f32(4000.0)), f32(1.0)) / f32(4000.0)

    start = (rng.uniform(0.0, span_start_max, size=R) * P).astype(np.int64)
    covered = np.arange(P)[:, None] >= start[None, :]
    covered &= rng.random((P, R), dtype=f32) >= gap_rate

    data = np.empty((P, R, 2), dtype=np.float32)
    data[:, :, 0] = np.where(covered, inten, np.nan)
    data[:, :, 1] = np.where(covered, dwell, np.nan)
    return SyntheticTranscript(name, data, samples, conditions, modified)


def make_config(config_index, n_transcripts, n_positions, reads_per_cond, **kw):
    """Deterministic list of transcripts for BASELINE config ``config_index`` (1-based)."""
    rng = np.random.default_rng(BASE_SEED + config_index)
    return [make_transcript(rng, n_positions, reads_per_cond, name=f'cfg{config_index}_tx{t}', **kw)
            for t in range(n_transcripts)]


# BASELINE.json configs (index -> full-size shape and comparison methods); config 1 is
# the reference's bundled fixture (tests/golden/fixture_cfg1.npz), not synthetic.
BASELINE_CONFIGS = {
    2: dict(n_transcripts=100, n_positions=1000, reads_per_cond=50, methods=['KS']),
    3: dict(n_transcripts=1000, n_positions=2000, reads_per_cond=100, methods=['GMM', 'KS']),
    4: dict(n_transcripts=20000, n_positions=1000, reads_per_cond=150, methods=['GMM', 'MW']),
    5: dict(n_transcripts=200, n_positions=1000, reads_per_cond=5000, methods=['GMM', 'KS']),
}


def make_skew_tail(rng, n_transcripts, n_positions=1000, lo=30, hi=200, mean=80, **kw):
    """Low-coverage tail of config 5: reads/condition ~ truncated geometric in [lo, hi]."""
    out = []
    for t in range(n_transcripts):
        n = int(min(hi, lo + rng.geometric(1.0 / max(mean - lo, 1))))
        out.append(make_transcript(rng, n_positions, n, name=f'cfg5_tail{t}', **kw))
    return out

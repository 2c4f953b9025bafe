"""Lane-level emulation of the strip form of the exact-KS wavefront (csrc/ks_pvalue.cu:strip_problem):
lane l owns the columns [l C, l C + C) of the band relative to its left edge and runs l rows behind
lane l - 1; the cell above the last column of a strip comes from lane l + 1's first column of the
row before (computed in the first step of the same period), the cell left of a strip from lane
l - 1's last column.  The emulation follows the kernel's schedule step by step (same shuffles, same
"exists above" rule) and must reproduce the oracle's lattice recurrence bitwise -- it checks the
schedule and the indexing without a GPU; the kernel itself is checked bitwise against SciPy in
tests/test_gpu_parity.py."""
import math

import numpy as np

from oracle.ks_exact import outer_prob_inside_method

def strip_emulation(m, n, g, h, c):
    if m < n: m, n = n, m
    mg, ng = m // g, n // g
    maxj0 = min((h + mg - 1) // mg, n + 1)
    L = 32
    # per-lane state
    U = np.zeros((L, c)); V = np.zeros((L, c))
    prev_min = np.zeros(L, int); prev_w = np.full(L, maxj0)
    result = None
    def limits(i):
        a0 = ng * i - h
        minj = min(max(a0 // mg + 1, 0), n)           # python floor division
        maxj = min(-((-(ng * i + h)) // mg), n + 1)
        return minj, maxj
    for p in range(m + L - 1):
        # row setup
        i_l = p - np.arange(L) + 1
        act = (i_l >= 1) & (i_l <= m)
        minj = np.zeros(L, int); maxj = np.zeros(L, int)
        for l in range(L):
            if act[l]: minj[l], maxj[l] = limits(int(i_l[l]))
        w = np.maximum(maxj - minj, 0)
        s = minj - prev_min
        assert ((s == 0) | (s == 1) | ~act).all(), (s, act)
        assert (w <= L * c).all()
        left = np.ones(L)
        for q in range(c):
            if q == 0:
                sh = np.roll(V[:, c - 1], 1)           # shfl_up(V[c-1], 1)
                left = np.where(np.arange(L) == 0, 1.0, sh)
            # up
            Unext = np.roll(V[:, 0], -1) if q == c - 1 else None   # shfl_down(V[0], 1)
            up = np.empty(L)
            for l in range(L):
                rp = l * c + q + s[l]
                if s[l] == 0: val = U[l, q]
                elif q + 1 < c: val = U[l, q + 1]
                else: val = Unext[l]
                up[l] = val if rp < prev_w[l] else 1.0
            rel = np.arange(L) * c + q
            active = act & (rel < w)
            j = minj + rel
            di = i_l.astype(float); dj = j.astype(float)
            with np.errstate(all='ignore'):
                val = (up * di + left * dj) / (di + dj)
            V[:, q] = np.where(active, val, V[:, q])
            left = V[:, q].copy()
        # row end
        for l in range(L):
            if act[l]:
                if i_l[l] == m:
                    rel = n - minj[l]
                    if 0 <= rel < w[l] and rel // c == l: result = V[l, rel % c]
                    # column n may belong to another lane processing row m at a different period
                U[l] = V[l]; prev_min[l] = minj[l]; prev_w[l] = w[l]
        # result by other lanes: handled when that lane processes row m
    return min(max(result, 0.0), 1.0) if result is not None else None


def test_strip_schedule_reproduces_the_lattice_recurrence():
    rng = np.random.default_rng(0)
    cases = [(37, 23, 0.3), (64, 63, 0.2), (100, 7, 0.5), (33, 32, 0.9), (90, 45, 0.25), (128, 97, 0.15),
             (40, 39, 0.05), (200, 3, 0.7), (65, 1, 0.5), (96, 95, 0.6)]
    for _ in range(25):
        a = int(rng.integers(2, 140))
        b = int(rng.integers(1, 140))
        if a == b:
            b += 1
        cases.append((a, b, float(rng.uniform(0.02, 0.9))))
    for a, b, d in cases:
        m, n = max(a, b), min(a, b)
        g = math.gcd(m, n)
        h = max(1, round(d * (m // g * n)))
        wmax = min(2 * h // (m // g) + 2, n + 1)     # strip_dispatch's bound on the cells of a row
        c = max(2, -(-wmax // 32))
        want = min(max(outer_prob_inside_method(m, n, g, h), 0.0), 1.0)
        assert strip_emulation(m, n, g, h, c) == want, (m, n, h, c)
        assert strip_emulation(m, n, g, h, c + 1) == want, (m, n, h, c + 1)   # a wider strip than needed

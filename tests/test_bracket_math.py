"""CPU check of the algorithm behind the bracketed threshold solve (fries_b200/csrc/compress.cuh: bracket_solve).

The CUDA kernels decide the exactly-preserved set of a compression from (a) the exact count and sum of everything at
or above a bracket [t_lo, t_hi) around the expected threshold, (b) Newton rounds restricted to the elements inside the
bracket, and two validity tests.  This file restates that procedure in numpy and checks, against the oracle's
find_preserve (the reference's rounds, compress_utils.cpp:29-105), that
  * a valid bracket yields exactly the reference's preserved set and budget, wherever the bracket sits,
  * a bracket that misses the fixed point is always reported invalid (the kernels then fall back to the plain rounds),
so the shortcut cannot silently change a result."""
import numpy as np
import pytest

import oraclelib


def bracket_solve(x, m, t_lo, t_hi):
    """numpy restatement: returns (valid, keep mask, budget left)"""
    hi = x >= t_hi
    cand = (x >= t_lo) & ~hi
    c_hi, s_hi = int(hi.sum()), float(x[hi].sum())
    nrem0, R0 = m - c_hi, float(x.sum()) - s_hi
    if nrem0 <= 0 or not (t_hi * nrem0 >= R0):  # lemma: H = {x >= t_hi} is preserved iff t_hi >= R(H) / n(H)
        return False, None, None
    cx = x[cand]
    kept = np.zeros(cx.size, bool)
    R, nrem = R0, nrem0
    while True:
        new = ~kept & (cx * nrem >= R)
        if not new.any():
            break
        kept |= new
        if kept.sum() >= nrem0:
            return False, None, None
        nrem = nrem0 - int(kept.sum())
        R = R0 - float(cx[kept].sum())
    if not (t_lo * nrem < R):  # nothing below the bracket may reach the final threshold
        return False, None, None
    x_cut = cx[kept].min() if kept.any() else t_hi
    return True, x >= min(x_cut, t_hi), nrem


def values(rng, n, kind):
    if kind == "lognormal":
        v = rng.lognormal(0, 2.5, n)
    elif kind == "uniform":
        v = rng.random(n)
    else:  # a few giants + dust + a cluster of equal magnitudes, like a compressed FRI iterate
        v = np.concatenate([rng.lognormal(6, 1, max(1, n // 100)), rng.lognormal(-3, 2, n // 2),
                            np.full(n - n // 2 - max(1, n // 100), 0.37)])
        rng.shuffle(v)
    return v


@pytest.mark.parametrize("kind", ["lognormal", "uniform", "fri"])
@pytest.mark.parametrize("n,budget", [(50, 10), (2000, 300), (20000, 2500), (20000, 15000)])
def test_bracket_gives_the_reference_fixed_point(kind, n, budget):
    rng = np.random.default_rng(n + budget)
    x = values(rng, n, kind)
    loc, glob, left, keep = oraclelib.find_preserve(x, budget)
    keep = keep.astype(bool)
    if keep.all() or left == 0:
        pytest.skip("degenerate: everything preserved")
    # the final threshold of the reference's rounds
    t_star = x[~keep].sum() / left
    n_valid = 0
    for centre in (1.0, 0.97, 1.04, 0.8, 1.3, 0.3, 3.0):
        for h in (1e-3, 0.02, 0.2):
            t_lo, t_hi = t_star * centre * (1 - h), t_star * centre * (1 + h)
            valid, k, nrem = bracket_solve(x, budget, t_lo, t_hi)
            if valid:
                n_valid += 1
                assert np.array_equal(k, keep), (centre, h, int((k != keep).sum()))
                assert nrem == left
            else:
                # an invalid verdict is only allowed when the bracket really misses the fixed point: the smallest
                # preserved and the largest resampled magnitude must not both lie inside it with room to spare
                x_keep_min = x[keep].min() if keep.any() else np.inf
                inside = t_lo <= t_star < t_hi and t_lo < x[~keep].max() and (not keep.any() or x_keep_min < t_hi)
                hi_ok = t_hi * (budget - (x >= t_hi).sum()) >= x[x < t_hi].sum()
                assert not (inside and hi_ok and t_lo * left < x[~keep].sum()), (centre, h)
    assert n_valid >= 3  # the centred brackets are valid


def test_lemma_downward_induction():
    """H = {x >= t} is inside the fixed point iff t >= R(H) / n(H): brute force over all prefixes of sorted vectors"""
    rng = np.random.default_rng(1)
    for _ in range(200):
        n, m = int(rng.integers(3, 30)), int(rng.integers(1, 25))
        x = np.sort(rng.lognormal(0, 1.5, n))[::-1]
        loc, glob, left, keep = oraclelib.find_preserve(x, m)
        k_star = int(keep.sum())
        assert keep[:k_star].all() and not keep[k_star:].any()  # the fixed point is a prefix of the sorted vector
        for c in range(1, min(n, m)):
            t = x[c - 1]  # H = the c largest elements
            if x[c] == t:
                continue
            lemma = t * (m - c) >= x[c:].sum()
            assert lemma == (c <= k_star), (c, k_star, x, m)

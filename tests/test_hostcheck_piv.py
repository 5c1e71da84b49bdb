"""CPU tier: the parallel restatement of the pivotal family that the kernels of fries_b200/csrc/piv.cu execute (the
__host__ __device__ rules of piv.cuh, run phase by phase on the host by tests/hostcheck) against the sequential oracle,
and the product's host-side piv_budget against the oracle and the compiled reference."""
import numpy as np
import pytest

import oraclelib as ol
import reflib
from hostcheck import hc
from test_oracle_piv import piv_case


def hc_piv_samp(v, norm, n_samp, keep, draws):
    import ctypes as C
    vv, kk = np.array(v, np.float64), np.array(keep, np.uint8)
    an = C.c_uint64(0)
    used = hc.lib().hc_piv_samp(vv, len(vv), kk, norm, n_samp, np.ascontiguousarray(draws, np.uint32), C.byref(an))
    return vv, kk, used, an.value


def last_unit_start(v, keep, norm, n_samp):
    """index of the first element that reaches into the last sampling unit"""
    w = np.where(keep == 0, np.abs(v), 0.0)
    E = np.cumsum(w)
    return int(np.searchsorted(E, (n_samp - 1) * (norm / n_samp) * (1 - 1e-12)))


def same_up_to_closing_unit(v, keep, norm, n_samp, a_v, a_k, b_v, b_k, rtol=0.0):
    """Two pivotal samplings of the same input with the same draws agree, except that the closing unit may pick a
    different one of its candidates: whether the prefix total reaches fl(n_samp * unit) decides if that unit has a
    straddling element, and two ways of summing differ in the last bit there.  The swap involves an element of the last
    unit and possibly the element carried into it (which can come from far back)."""
    diff = np.flatnonzero((a_v != 0) != (b_v != 0))
    assert len(diff) <= 2, diff
    if len(diff):
        assert diff.max() >= last_unit_start(v, keep, norm, n_samp), diff
    same = np.ones(len(v), bool)
    same[diff] = False
    assert np.array_equal(a_k[same], b_k[same])
    if rtol:
        assert np.allclose(a_v[same], b_v[same], rtol=rtol, atol=0)
    else:
        assert np.array_equal(a_v[same], b_v[same])
    return len(diff)


@pytest.mark.parametrize("seed,n,n_samp,fk,zeros", [(1, 200, 20, 0.0, 0.0), (2, 5000, 700, 0.1, 0.05),
                                                     (3, 40000, 9000, 0.3, 0.0), (4, 64, 40, 0.0, 0.0),
                                                     (5, 3000, 1, 0.2, 0.0), (6, 3000, 0, 0.2, 0.1),
                                                     (7, 20000, 5000, 0.0, 0.3), (8, 300000, 100000, 0.2, 0.0)])
def test_parallel_pivotal_rules_match_the_sequential_sweep(seed, n, n_samp, fk, zeros):
    v, keep, norm = piv_case(seed, n, max(n_samp, 1), fk, zeros)
    draws = ol.mt19937(seed, 2 * n_samp + 8)
    ov, ok, oused = ol.piv_samp_serial(v, norm, n_samp, keep, draws)
    hv, hk, hused, anomalies = hc_piv_samp(v, norm, n_samp, keep, draws)
    assert anomalies == 0
    assert hused == oused
    if n_samp == 0:
        assert np.array_equal(hv, ov) and np.array_equal(hk, ok)
        return
    same_up_to_closing_unit(v, keep, norm, n_samp, hv, hk, ov, ok)
    samp = (keep == 0) & (hv != 0)
    assert samp.sum() == n_samp and np.all(np.abs(hv[samp]) == norm / n_samp)
    assert np.array_equal(hv[keep == 1], v[keep == 1]) and not hk[keep == 1].any()
    assert np.array_equal(hk == 1, hv == 0)


def test_parallel_pivotal_is_unbiased():
    # E[output] = input for every element (the compression is unbiased): CLT bound over many draws
    rng = np.random.default_rng(5)
    n, n_samp, reps = 60, 12, 20000
    v, keep, norm = piv_case(11, n, n_samp, 0.1)
    acc = np.zeros(n)
    for r in range(reps):
        draws = rng.integers(0, 2**32, 2 * n_samp, dtype=np.uint64).astype(np.uint32)
        hv, _, _, an = hc_piv_samp(v, norm, n_samp, keep, draws)
        assert an == 0
        acc += hv
    mean = acc / reps
    unit = norm / n_samp
    p = np.where(keep == 1, 0.0, np.abs(v) / unit)  # inclusion probability of a resampled element
    sd = unit * np.sqrt(p * (1 - p) / reps)
    assert np.all(np.abs(mean - v) <= 5 * sd + 1e-12)
    assert np.allclose(mean[keep == 1], v[keep == 1], rtol=1e-12, atol=0)


@pytest.mark.parametrize("seed,n,n_tot,exp_loc,up", [(1, 500, 1000, 37.4, True), (2, 500, 1000, 37.4, False),
                                                      (3, 3000, 9000, 411.9, True), (4, 3000, 9000, 411.05, False),
                                                      (5, 100, 50, 9.5, True), (6, 100, 50, 9.5, False),
                                                      (7, 500, 1000, 37.4, None)])
def test_parallel_adjust_probs_matches_the_walk(seed, n, n_tot, exp_loc, up):
    import ctypes as C
    rng = np.random.default_rng(200 + seed)
    v = rng.random(n) * np.where(rng.random(n) < 0.5, -1.0, 1.0)
    keep = (rng.random(n) < 0.1).astype(np.uint8)
    tot_norm = 5000.0
    unit = tot_norm / n_tot
    free = np.flatnonzero(keep == 0)
    if up is None:  # nothing too big: untouched
        v[free] *= exp_loc * unit / np.abs(v[free]).sum()
        up = True
    else:
        hot = rng.choice(free, 5, replace=False)
        rest = np.setdiff1d(free, hot)
        v[hot] = np.sign(v[hot]) * 0.9999 * unit
        v[rest] *= (exp_loc - 5 * 0.9999) * unit / np.abs(v[rest]).sum()
    n_loc = int(np.ceil(exp_loc)) if up else int(exp_loc)
    ov, ok, on, onorm = ol.adjust_probs(v, n_loc, exp_loc, n_tot, tot_norm, keep)
    hv, hk = v.copy(), keep.copy()
    nl = C.c_uint32(n_loc)
    hnorm = hc.lib().hc_adjust_probs(hv, n, hk, C.byref(nl), exp_loc, n_tot, tot_norm)
    assert nl.value == on and hnorm == onorm
    assert np.array_equal(hk, ok)
    # values: the element where the walk stops gets a correction computed from a prefix sum instead of two running
    # counters -> 1e-12 relative to one sampling unit
    assert np.allclose(hv, ov, rtol=0, atol=1e-12 * unit)
    assert (hv != ov).sum() <= 1


@pytest.mark.parametrize("seed,n_procs,n_samp", [(1, 1, 100), (2, 2, 7), (3, 4, 1000), (4, 8, 33), (5, 8, 100000),
                                                  (6, 3, 2), (7, 5, 5)])
def test_product_piv_budget(seed, n_procs, n_samp):
    import fries_b200
    rng = np.random.default_rng(100 + seed)
    ln = rng.random(n_procs) * 1000
    draws = ol.mt19937(seed, 2 * n_procs + 8)
    ob, oused = ol.piv_budget(ln, n_samp, draws)
    pb, pused = fries_b200.piv_budget(ln, n_samp, draws)
    assert pused == oused and int(pb.sum()) == n_samp
    exp = ln / ln.sum() * n_samp
    assert np.all(np.abs(pb.astype(float) - exp) < 1 + 1e-9)
    # same budgets as the sequential sweep except when the closing unit differs (see above): at most one unit moves
    assert np.abs(pb.astype(int) - ob.astype(int)).sum() <= 2
    if reflib.available():
        rb, rused = reflib.piv_budget(ln, n_samp, seed)
        assert np.array_equal(ob, rb) and rused == oused


@pytest.mark.parametrize("n_procs,n_samp", [(1, 5), (2, 4), (3, 2), (8, 1), (4, 100)])
def test_product_piv_budget_nothing_left(n_procs, n_samp):
    """Every rank's residual norm is zero (all elements preserved) while samples are left over -- the exact limit of the
    HB-PP pipeline (tests/test_hbpp_exact_limit.py).  The reference's sweep then runs on a zero unit, steps over one rank
    per unit and still draws twice each time: budgets stay zero, and the generator must advance by the same count."""
    import fries_b200
    ln = np.zeros(n_procs)
    draws = ol.mt19937(11, 2 * n_procs + 8)
    ob, oused = ol.piv_budget(ln, n_samp, draws)
    pb, pused = fries_b200.piv_budget(ln, n_samp, draws)
    assert np.array_equal(pb, ob) and not pb.any()
    assert pused == oused == 2 * min(n_procs, n_samp)
    # the compiled reference is not consulted here: 0 / 0 -> (uint32_t)NaN is undefined behaviour in piv_budget
    # (compress_utils.cpp:570) and this build of it faults; the oracle restates the x86 outcome (budget 0)

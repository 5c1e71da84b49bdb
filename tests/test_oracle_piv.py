"""CPU tier: the plain-C restatement of the pivotal compression family (oracle/fries_oracle.c: fo_piv_*) pinned
bit-for-bit against the compiled reference (compress_utils.cpp:354-681) on seeded inputs, plus the reference's own
known-answer test for piv_samp_serial (tests/test_compression.cpp:64-118: with every element preserved and a zero
budget the vector comes back unchanged)."""
import numpy as np
import pytest

import oraclelib as ol
import reflib
from golden_cases import piv_samp_inputs

needs_ref = pytest.mark.skipif(not reflib.available(), reason="oracle/_ref not built")


def piv_case(seed, n, n_samp, frac_keep=0.0, zeros=0.0):
    return piv_samp_inputs((seed, n, n_samp, frac_keep, zeros))


@needs_ref
def test_mt19937_matches_std():
    for seed in (0, 1, 5489, 123456789):
        assert np.array_equal(ol.mt19937(seed, 1500), reflib.mt19937(seed, 1500))


def test_mt19937_known_answer():
    # the C++ standard's check value: the 10000th output of default-seeded mt19937 is 4123659995
    assert int(ol.mt19937(5489, 10000)[-1]) == 4123659995


@needs_ref
@pytest.mark.parametrize("seed,n,n_samp,fk,zeros", [(1, 200, 20, 0.0, 0.0), (2, 5000, 700, 0.1, 0.05),
                                                     (3, 40000, 9000, 0.3, 0.0), (4, 64, 40, 0.0, 0.0),
                                                     (5, 3000, 1, 0.2, 0.0), (6, 3000, 0, 0.2, 0.1),
                                                     (7, 20000, 5000, 0.0, 0.3)])
def test_piv_samp_serial_matches_reference(seed, n, n_samp, fk, zeros):
    v, keep, norm = piv_case(seed, n, max(n_samp, 1), fk, zeros)
    draws = ol.mt19937(seed, 2 * n_samp + 8)
    ov, ok, oused = ol.piv_samp_serial(v, norm, n_samp, keep, draws)
    rv, rk, rused = reflib.piv_samp_serial(v, norm, n_samp, keep, seed)
    assert oused == rused
    assert np.array_equal(ok, rk)
    assert np.array_equal(ov, rv)
    if n_samp:
        # exactly n_samp samples of magnitude norm / n_samp, preserved elements untouched
        samp = (keep == 0) & (ov != 0)
        assert samp.sum() == n_samp and np.allclose(np.abs(ov[samp]), norm / n_samp, rtol=0, atol=0)
        assert np.array_equal(ov[keep == 1], v[keep == 1]) and not ok[keep == 1].any()
        assert np.array_equal(np.sign(ov[samp]), np.sign(v[samp]))


def test_piv_samp_serial_identity_known_answer():
    # tests/test_compression.cpp:96-117: all preserved, n_samp = 0 -> unchanged
    rng = np.random.default_rng(0)
    v = rng.standard_normal(10)
    ov, ok, used = ol.piv_samp_serial(v, 0.0, 0, np.ones(10, np.uint8), np.zeros(4, np.uint32))
    assert np.array_equal(ov, v) and not ok.any() and used == 0


@needs_ref
@pytest.mark.parametrize("seed,n_procs,n_samp", [(1, 1, 100), (2, 2, 7), (3, 4, 1000), (4, 8, 33), (5, 8, 100000),
                                                  (6, 3, 2), (7, 5, 5)])
def test_piv_budget_matches_reference(seed, n_procs, n_samp):
    rng = np.random.default_rng(100 + seed)
    ln = rng.random(n_procs) * 1000
    draws = ol.mt19937(seed, 2 * n_procs + 8)
    ob, oused = ol.piv_budget(ln, n_samp, draws)
    rb, rused = reflib.piv_budget(ln, n_samp, seed)
    assert np.array_equal(ob, rb) and oused == rused
    assert int(ob.sum()) == n_samp
    exp = ln / ln.sum() * n_samp
    assert np.all(np.abs(ob.astype(float) - exp) < 1 + 1e-9)


@needs_ref
@pytest.mark.parametrize("seed,n,n_tot,exp_loc,up", [(1, 500, 1000, 37.4, True), (2, 500, 1000, 37.4, False),
                                                      (3, 3000, 9000, 411.9, True), (4, 3000, 9000, 411.05, False),
                                                      (5, 100, 50, 9.5, True), (6, 100, 50, 9.5, False)])
def test_adjust_probs_matches_reference(seed, n, n_tot, exp_loc, up):
    rng = np.random.default_rng(200 + seed)
    v = rng.random(n) * np.where(rng.random(n) < 0.5, -1.0, 1.0)
    keep = (rng.random(n) < 0.1).astype(np.uint8)
    # this rank's residual norm is exp_loc sampling units of the global tot_norm / n_tot; no element reaches one unit
    # (find_preserve took those), but a few exceed loc_norm / ceil(exp_loc), which is what triggers the adjustment
    tot_norm = 5000.0
    unit = tot_norm / n_tot
    free = np.flatnonzero(keep == 0)
    hot = rng.choice(free, 5, replace=False)
    rest = np.setdiff1d(free, hot)
    v[hot] = np.sign(v[hot]) * 0.9999 * unit
    v[rest] *= (exp_loc - 5 * 0.9999) * unit / np.abs(v[rest]).sum()
    assert (np.abs(v[rest]) < 0.9 * unit).all()
    n_loc = int(np.ceil(exp_loc)) if up else int(exp_loc)
    ov, ok, on, onorm = ol.adjust_probs(v, n_loc, exp_loc, n_tot, tot_norm, keep)
    rv, rk, rn, rnorm = reflib.adjust_probs(v, n_loc, exp_loc, n_tot, tot_norm, keep)
    assert on == rn and onorm == rnorm
    assert np.array_equal(ok, rk) and np.array_equal(ov, rv)
    assert (ov != v).any()  # the adjustment happened
    # afterwards the inclusion probabilities |v| / unit of the non-preserved elements add up to the integer budget
    assert np.abs(ov[ok == 0]).sum() / unit == pytest.approx(on, abs=1e-9)


@needs_ref
@pytest.mark.parametrize("seed,n,m", [(1, 300, 40), (2, 20000, 3000), (3, 20000, 19990), (4, 1000, 2000)])
def test_piv_comp_parallel_matches_reference(seed, n, m):
    rng = np.random.default_rng(300 + seed)
    v = rng.standard_normal(n) * np.exp(3 * rng.standard_normal(n))
    draws = ol.mt19937(seed, 2 * m + 8)
    ov, ok, oused = ol.piv_comp(v, m, draws)
    rv, rk, rused = reflib.piv_comp_parallel(v, m, seed)
    assert oused == rused
    assert np.array_equal(ok, rk) and np.array_equal(ov, rv)
    assert (ov != 0).sum() <= m and np.array_equal(ok == 1, ov == 0)


# ---- apply_HBPP_piv (heat_bathPP.cpp:1014-1419): the pivotal twin of apply_HBPP_sys --------------------------------
@needs_ref
@pytest.mark.parametrize("case,new_hb,n_det,n_samp", [(("ne", 2, False), 1, 1, 50), (("ne", 2, False), 0, 300, 1000),
                                                       (("ne", 2, True), 1, 300, 1000), (("n2", 7, True), 1, 2000, 4000),
                                                       (("h2o", 3, True), 0, 2000, 3000),
                                                       (((8, 6, 0, [0, 0, 1, 2, 3, 0, 1, 2]), 4, True), 1, 40, 3000)])
def test_apply_hbpp_piv_matches_reference(case, new_hb, n_det, n_samp):
    from fries_b200.synth import SynthMol
    from golden_cases import make_values
    sm = SynthMol(*case)
    om, rm = ol.OracleMol(sm), reflib.RefMol(sm)
    rng = np.random.default_rng(n_det + new_hb)
    keys = np.concatenate([[sm.hf], sm.random_dets(n_det - 1, rng, 0)]).astype(np.uint64) if n_det > 1 else \
        np.array([sm.hf], np.uint64)
    vals = make_values(rng, n_det, "fri")
    vals[0] = 100.0
    cap = 4 * n_samp + 4 * n_det
    for seed in (1, 2):
        rv, rd, ro, rused = rm.apply_hbpp_piv(keys, vals, 0.97, new_hb, seed, n_samp, cap)
        ov, od, oo, oused = om.apply_hbpp_piv(keys, vals, 0.97, new_hb, ol.mt19937(seed, 12 * n_samp + 64), n_samp, cap)
        assert oused == rused and len(ov) == len(rv) > 0
        assert np.array_equal(od, rd) and np.array_equal(oo, ro)
        assert np.allclose(ov, rv, rtol=1e-12, atol=0)
        print(f"apply_hbpp_piv {case[0]} new_hb={new_hb}: {len(rv)} samples, {rused} draws")


@needs_ref
def test_apply_hbpp_piv_exact_limit_equals_sys():
    """tests/test_hamiltonian.cpp:507-519: with a budget above the number of excitations the pivotal and the systematic
    pipeline return the same tuples (here also the same values up to the sign the pivotal variant includes)"""
    from fries_b200.synth import SynthMol
    sm = SynthMol("ne", 2, True)
    om = ol.OracleMol(sm)
    hf = np.array([sm.hf], np.uint64)
    n_ex = len(om.sing_ex(int(sm.hf))) + len(om.doub_ex(int(sm.hf)))
    sv, sd, so = om.apply_hbpp_sys(hf, np.ones(1), 0.95, 1, np.full(5, 0.5), 8 * n_ex, 16 * n_ex)
    pv, pd, po, used = om.apply_hbpp_piv(hf, np.ones(1), 0.95, 1, ol.mt19937(0, 64), 8 * n_ex, 16 * n_ex)
    assert used == 0 or used % 2 == 0
    assert np.array_equal(po, so) and len(pv) == n_ex
    assert np.allclose(np.abs(pv), np.abs(sv), rtol=1e-12, atol=0)

"""GPU tier: the pivotal compression family (fries_piv_samp_serial / _dev, fries_adjust_probs, fries_piv_comp;
reference compress_utils.cpp:354-681) through the C-ABI against the sequential oracle (itself pinned bit-for-bit
against the compiled reference in tests/test_oracle_piv.py) on the same inputs and the same mt19937 draws.

Bar: identical sampled index sets and values, except (a) the closing sampling unit -- whether the prefix total reaches
fl(n_samp * unit) decides if it has a straddling element, and a chunked device sum and a running sum differ in the last
bit there -- and (b) counted FP-boundary ties; adjust_probs values to 1e-12 of a sampling unit."""
import ctypes as C

import numpy as np
import pytest

import oraclelib as ol
from test_hostcheck_piv import same_up_to_closing_unit
from test_oracle_piv import piv_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import fries_b200
    c = fries_b200.Context(0)
    yield c
    c.close()


def check_samples(v, keep, norm, n_samp, gv, gk, ov, ok, rtol=0.0):
    n_diff = same_up_to_closing_unit(v, keep, norm, n_samp, gv, gk, ov, ok, rtol)
    print(f"n={len(v)} n_samp={n_samp}: elements swapped by the closing unit: {n_diff}")
    samp = (keep == 0) & (gv != 0)
    assert samp.sum() == n_samp
    assert np.array_equal(np.sign(gv[samp]), np.sign(v[samp]))
    assert np.array_equal(gk == 1, gv == 0)
    assert np.array_equal(gv[keep == 1], v[keep == 1])


@pytest.mark.parametrize("seed,n,n_samp,fk,zeros", [(1, 200, 20, 0.0, 0.0), (2, 5000, 700, 0.1, 0.05),
                                                     (3, 40000, 9000, 0.3, 0.0), (4, 64, 40, 0.0, 0.0),
                                                     (5, 3000, 1, 0.2, 0.0), (6, 3000, 0, 0.2, 0.1),
                                                     (7, 20000, 5000, 0.0, 0.3), (8, 300000, 100000, 0.2, 0.0),
                                                     (9, 2000000, 400000, 0.1, 0.0), (10, 1, 1, 0.0, 0.0),
                                                     (11, 513, 100, 0.0, 0.0)])
def test_piv_samp_serial(ctx, seed, n, n_samp, fk, zeros):
    import fries_b200
    v, keep, norm = piv_case(seed, n, max(n_samp, 1), fk, zeros)
    draws = ol.mt19937(seed, 2 * n_samp + 8)
    ov, ok, oused = ol.piv_samp_serial(v, norm, n_samp, keep, draws)
    gv, gk, gused = fries_b200.piv_samp_serial(ctx, v, norm, n_samp, keep, draws)
    assert gused == oused
    if n_samp == 0:
        assert np.array_equal(gv, ov) and np.array_equal(gk, ok)
        return
    check_samples(v, keep, norm, n_samp, gv, gk, ov, ok)
    assert np.all(np.abs(gv[(keep == 0) & (gv != 0)]) == norm / n_samp)


def test_piv_samp_empty_and_all_preserved(ctx):
    import fries_b200
    # tests/test_compression.cpp:96-117: all preserved, n_samp = 0 -> unchanged
    rng = np.random.default_rng(0)
    v = rng.standard_normal(10)
    gv, gk, used = fries_b200.piv_samp_serial(ctx, v, 0.0, 0, np.ones(10, np.uint8), np.zeros(4, np.uint32))
    assert np.array_equal(gv, v) and not gk.any() and used == 0
    gv, gk, used = fries_b200.piv_samp_serial(ctx, np.zeros(0), 0.0, 0, np.zeros(0, np.uint8), np.zeros(4, np.uint32))
    assert gv.size == 0 and used == 0


def test_piv_samp_dev_resident(ctx):
    """device-resident form on torch buffers, on the caller's stream"""
    import torch
    from fries_b200._capi import check, lib
    v, keep, norm = piv_case(21, 100000, 30000, 0.2)
    n_samp = 30000
    draws = ol.mt19937(21, 2 * n_samp)
    ov, ok, _ = ol.piv_samp_serial(v, norm, n_samp, keep, draws)
    dv = torch.from_numpy(v.copy()).cuda()
    dk = torch.from_numpy(keep.copy()).cuda()
    dd = torch.from_numpy(draws.view(np.int32).copy()).cuda()
    work = torch.empty(8 * len(v) + 12 * n_samp + 2048, dtype=torch.uint8, device="cuda")
    res = torch.zeros(4, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    check(lib.fries_piv_samp_dev(ctx.h, dv.data_ptr(), len(v), norm, n_samp, dk.data_ptr(), dd.data_ptr(),
                                 work.data_ptr(), work.numel(), res.data_ptr()))
    ctx.sync()
    gv, gk, r = dv.cpu().numpy(), dk.cpu().numpy(), res.cpu().numpy()
    check_samples(v, keep, norm, n_samp, gv, gk, ov, ok)
    assert r[1] == n_samp and r[2] == n_samp and r[3] == 0
    assert r[0] == pytest.approx(np.abs(gv).sum(), rel=1e-12)
    # too small a work area is refused
    from fries_b200._capi import FriesError
    with pytest.raises(FriesError):
        check(lib.fries_piv_samp_dev(ctx.h, dv.data_ptr(), len(v), norm, n_samp, dk.data_ptr(), dd.data_ptr(),
                                     work.data_ptr(), 1000, None))


@pytest.mark.parametrize("seed,n,n_tot,exp_loc,up", [(1, 500, 1000, 37.4, True), (2, 500, 1000, 37.4, False),
                                                      (3, 3000, 9000, 411.9, True), (4, 3000, 9000, 411.05, False),
                                                      (5, 100, 50, 9.5, True), (6, 100, 50, 9.5, False),
                                                      (7, 500, 1000, 37.4, None), (8, 400000, 900000, 41100.6, True),
                                                      (9, 400000, 900000, 41100.6, False)])
def test_adjust_probs(ctx, seed, n, n_tot, exp_loc, up):
    import fries_b200
    rng = np.random.default_rng(200 + seed)
    v = rng.random(n) * np.where(rng.random(n) < 0.5, -1.0, 1.0)
    keep = (rng.random(n) < 0.1).astype(np.uint8)
    tot_norm = 5000.0
    unit = tot_norm / n_tot
    free = np.flatnonzero(keep == 0)
    if up is None:
        v[free] *= exp_loc * unit / np.abs(v[free]).sum()
        up = True
    else:
        hot = rng.choice(free, 5, replace=False)
        rest = np.setdiff1d(free, hot)
        v[hot] = np.sign(v[hot]) * 0.9999999 * unit
        v[rest] *= (exp_loc - 5 * 0.9999999) * unit / np.abs(v[rest]).sum()
    n_loc = int(np.ceil(exp_loc)) if up else int(exp_loc)
    ov, ok, on, onorm = ol.adjust_probs(v, n_loc, exp_loc, n_tot, tot_norm, keep)
    gv, gk, gn, gnorm = fries_b200.adjust_probs(ctx, v, n_loc, exp_loc, n_tot, tot_norm, keep)
    assert gn == on and gnorm == onorm
    assert np.array_equal(gk, ok)
    assert np.allclose(gv, ov, rtol=0, atol=1e-10 * unit)
    assert (gv != ov).sum() <= 1


@pytest.mark.parametrize("seed,n,m", [(1, 300, 40), (2, 20000, 3000), (3, 20000, 19990), (4, 1000, 2000),
                                      (5, 1000000, 200000)])
def test_piv_comp_parallel(ctx, seed, n, m):
    import fries_b200
    rng = np.random.default_rng(300 + seed)
    v = rng.standard_normal(n) * np.exp(3 * rng.standard_normal(n))
    draws = ol.mt19937(seed, 2 * m + 8)
    ov, ok, oused = ol.piv_comp(v, m, draws)
    gv, gk, gused = fries_b200.piv_comp(ctx, v, m, draws)
    assert gused == oused
    assert np.array_equal(gk == 1, gv == 0) and (gv != 0).sum() <= m
    # same preserved set (bit-exact, as find_preserve), same samples up to the closing unit; the sampling unit comes from a
    # residual norm summed in a different order -> values to 1e-12
    loc, glob, left, keep = ol.find_preserve(v, m)
    assert np.array_equal(gv[keep == 1], v[keep == 1])
    if left and loc > 0:
        check_samples(v, keep, loc, left, gv, gk, ov, ok, rtol=1e-12)
    else:
        assert np.array_equal(gv, ov)


def test_piv_comp_is_unbiased(ctx):
    import fries_b200
    rng = np.random.default_rng(8)
    n, m, reps = 80, 20, 4000
    v = rng.standard_normal(n) * np.exp(1.5 * rng.standard_normal(n))
    loc, glob, left, keep = ol.find_preserve(v, m)
    acc = np.zeros(n)
    for r in range(reps):
        draws = rng.integers(0, 2**32, 2 * m + 2, dtype=np.uint64).astype(np.uint32)
        gv, gk, used = fries_b200.piv_comp(ctx, v, m, draws)
        assert (gv != 0).sum() == m
        acc += gv
    mean = acc / reps
    unit = loc / left
    p = np.where(keep == 1, 0.0, np.abs(v) / unit)
    sd = unit * np.sqrt(p * (1 - p) / reps)
    assert np.all(np.abs(mean - v) <= 5 * sd + 1e-9 * np.abs(v))


@pytest.mark.parametrize("method", ["piv", "sys", "multi"])
def test_vec_compress_rows(ctx, method):
    """compress_vecs / compress_vecs_sys / compress_vecs_multi (vec_utils.cpp:10-127) on the resident store: rows 1 and 2 of a 3-row vector are
    compressed, row 0 is left alone, and elements that end up zero in all three rows disappear"""
    import fries_b200
    from fries_b200.synth import SynthMol
    from test_hostcheck_piv import same_up_to_closing_unit
    sm = SynthMol("ne", 2, False)
    rng = np.random.default_rng(17)
    n, m = 6000, 700
    keys = sm.random_dets(n, rng, None).astype(np.uint64)
    vals = rng.standard_normal((3, n)) * np.exp(2 * rng.standard_normal((3, n)))
    vals[0, rng.random(n) < 0.6] = 0       # row 0: many zeros, so that some elements die
    vals[2, rng.random(n) < 0.2] = 0
    scr = rng.integers(0, 2**32, sm.n_bits, dtype=np.uint64).astype(np.uint32)
    vec = fries_b200.Vec(ctx, 2 * n, sm.n_bits, sm.n_elec, 3, scr, scr)
    vec.upload(keys, vals)
    k0, v0 = vec.download()
    assert np.array_equal(k0, keys) and np.array_equal(v0, vals)
    draws = ol.mt19937(99, 8 * m + 16)
    used = vec.compress(1, 3, m, draws, method)
    # expected: the oracle row by row on consecutive draws
    exp = vals.copy()
    o_used = 0
    swaps = []
    for r in (1, 2):
        if method == "piv":
            ov, ok, u = ol.piv_comp(vals[r], m, draws[o_used:])
            o_used += u
        elif method == "multi":
            ov, u = ol.compress_multi_row(vals[r], m, draws[o_used:])
            o_used += u
        else:
            loc, glob, left, keep = ol.find_preserve(vals[r], m)
            ov, ok, _ = ol.sys_comp(vals[r], [loc], left, keep, draws[o_used] / 4294967296.0)
            o_used += 1
        exp[r] = ov
    assert used == o_used
    gk, gv = vec.download()
    # map the surviving elements back to their original positions (compaction is stable)
    pos = np.flatnonzero(np.isin(keys, gk))
    assert np.array_equal(keys[pos], gk)
    full = np.zeros_like(vals)
    full[:, pos] = gv
    assert np.array_equal(full[0], vals[0])
    for r in (1, 2):
        if method == "multi":
            # multinomial: every value is a whole number of norm / m quanta; the same draws land on the same elements (the
            # row's one-norm is summed in another order on the device: the quantum may differ in the last place)
            assert np.array_equal(full[r] != 0, exp[r] != 0) and 0 < (full[r] != 0).sum() <= m
            assert np.allclose(full[r], exp[r], rtol=1e-12, atol=0)
            quanta = np.abs(full[r]) / (np.abs(vals[r]).sum() / m)
            assert np.allclose(quanta, np.round(quanta), atol=1e-9) and int(np.round(quanta).sum()) == m
            continue
        assert (full[r] != 0).sum() == m
        if method == "piv":
            loc, glob, left, keep = ol.find_preserve(vals[r], m)
            kz = np.zeros(n, np.uint8)
            swaps.append(same_up_to_closing_unit(vals[r], keep, loc, left, full[r], kz, exp[r], kz, rtol=1e-12))
        else:
            ties = int(((full[r] != 0) != (exp[r] != 0)).sum())
            assert ties <= 2  # FP-boundary ties of the systematic grid, as in test_gpu_parity.py
            same = (full[r] != 0) == (exp[r] != 0)
            assert np.allclose(full[r][same], exp[r][same], rtol=1e-12, atol=0)
    # exactly the elements that are zero in every row are gone
    dead = ~np.any(full != 0, axis=0)
    assert dead.sum() > 0 and len(gk) == n - dead.sum()
    assert not np.isin(keys[dead], gk).any()
    vec.close()

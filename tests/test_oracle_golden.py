"""CPU tier: the plain-C oracle against the committed golden fixture (tests/golden/fries_golden.npz, generated from
the compiled reference by tests/golden/make_golden.py).  Runs where /root/reference does not exist."""
import os

import numpy as np
import pytest

import oraclelib
from fries_b200.synth import SynthMol
from golden_cases import (COMP_SUB_CASES, HBPP_CASES, MOL_CASES, PIV_ADJUST_CASES, PIV_BUDGET_CASES, PIV_COMP_CASES,
                          PIV_HBPP_CASES, PIV_SAMP_CASES, VEC_COMP_CASES, comp_sub_inputs, hbpp_inputs, mol_keys, piv_adjust_inputs,
                          piv_budget_inputs, piv_comp_inputs, piv_samp_inputs, vec_values)

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fries_golden.npz"))
GP = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "piv_golden.npz"))


def test_hash_golden():
    for tag in ("tv1", "tv2", "rnd"):
        h, o = oraclelib.hash_keys(G[f"hash_{tag}_keys"], G[f"hash_{tag}_scr"], 8)
        assert np.array_equal(h, G[f"hash_{tag}_h"]) and np.array_equal(o, G[f"hash_{tag}_o8"])


@pytest.mark.parametrize("i", range(len(VEC_COMP_CASES)))
def test_vector_compression_golden(i):
    case = VEC_COMP_CASES[i]
    v = vec_values(case)
    loc, glob, left, keep = oraclelib.find_preserve(v, case[1])
    assert loc == G[f"fp{i}_loc"] and glob == G[f"fp{i}_glob"] and left == G[f"fp{i}_left"]
    assert np.array_equal(keep, G[f"fp{i}_keep"])
    for j, rn in enumerate((0.0, 0.37, 0.999999)):
        ov, ok, on = oraclelib.sys_comp(v, [loc], left, keep, rn)
        assert np.array_equal(ov, G[f"sc{i}_{j}_v"]) and np.array_equal(ok, G[f"sc{i}_{j}_k"]) and on[0] == G[f"sc{i}_{j}_norm"]


@pytest.mark.parametrize("i", range(len(COMP_SUB_CASES)))
def test_comp_sub_golden(i):
    case = COMP_SUB_CASES[i]
    v, nd, sw, ss = comp_sub_inputs(case)
    for j, rn in enumerate((0.123, 0.9)):
        ov, oi, _, _ = oraclelib.comp_sub(v, nd, sw, ss, case[2], rn, 4 * max(case[2], case[0]) + 64)
        assert np.array_equal(oi, G[f"cs{i}_{j}_i"]) and np.array_equal(ov, G[f"cs{i}_{j}_v"])


@pytest.mark.parametrize("i", range(len(MOL_CASES)))
def test_molecule_golden(i):
    sm = SynthMol(*MOL_CASES[i])
    om = oraclelib.OracleMol(sm)
    for k, t in om.hb_tables().items():
        assert np.array_equal(t, G[f"mol{i}_{k}"]), k
    keys = mol_keys(sm)
    assert np.array_equal(om.diag(keys), G[f"mol{i}_diag"])
    so, do = G[f"mol{i}_sing_off"], G[f"mol{i}_doub_off"]
    for n, k in enumerate(keys[:6]):
        se, de = om.sing_ex(k), om.doub_ex(k)
        assert np.array_equal(se, G[f"mol{i}_sing_ex"][so[n]:so[n + 1]])
        assert np.array_equal(de, G[f"mol{i}_doub_ex"][do[n]:do[n + 1]])
        assert np.array_equal(om.sing_el([k] * len(se), se), G[f"mol{i}_sing_el"][so[n]:so[n + 1]])
        assert np.array_equal(om.doub_el(de), G[f"mol{i}_doub_el"][do[n]:do[n + 1]])
    w0 = np.concatenate([[om.hb_wt(0, k, o) for o in om.doub_ex(k)[::9]] for k in keys[:6]])
    w1 = np.concatenate([[om.hb_wt(1, k, o) for o in om.doub_ex(k)[::9]] for k in keys[:6]])
    assert np.allclose(w0, G[f"mol{i}_wt0"], rtol=1e-14, atol=0) and np.allclose(w1, G[f"mol{i}_wt1"], rtol=1e-14, atol=0)
    hk, hv = om.h_apply(keys[:5], np.linspace(-1, 1, 5) + 0.1, 1.0, -0.01)
    assert np.array_equal(hk, G[f"mol{i}_hv_keys"])
    assert np.allclose(hv, G[f"mol{i}_hv_vals"], rtol=1e-12, atol=1e-13 * np.abs(hv).max())


@pytest.mark.parametrize("i", range(len(HBPP_CASES)))
def test_apply_hbpp_sys_golden(i):
    case = HBPP_CASES[i]
    sm = SynthMol(*case[0])
    om = oraclelib.OracleMol(sm)
    keys, vals = hbpp_inputs(sm, case)
    ov, od, oo = om.apply_hbpp_sys(keys, vals, 0.97, case[3], G[f"hb{i}_uni"], case[2], 4 * case[2] + 4 * case[1])
    assert np.array_equal(od, G[f"hb{i}_d"]) and np.array_equal(oo, G[f"hb{i}_o"])
    assert np.allclose(ov, G[f"hb{i}_v"], rtol=1e-13, atol=0)


# ---- pivotal family (tests/golden/piv_golden.npz, generator tests/golden/make_piv_golden.py) -----------------------
def test_mt19937_golden():
    assert np.array_equal(oraclelib.mt19937(5489, 10000)[-4:], GP["mt5489_10000"])


@pytest.mark.parametrize("i", range(len(PIV_SAMP_CASES)))
def test_piv_samp_serial_golden(i):
    case = PIV_SAMP_CASES[i]
    v, keep, norm = piv_samp_inputs(case)
    ov, ok, used = oraclelib.piv_samp_serial(v, norm, case[2], keep, oraclelib.mt19937(case[0], 2 * case[2] + 8))
    assert used == GP[f"ps{i}_used"] and np.array_equal(ov, GP[f"ps{i}_v"]) and np.array_equal(ok, GP[f"ps{i}_k"])


@pytest.mark.parametrize("i", range(len(PIV_BUDGET_CASES)))
def test_piv_budget_golden(i):
    case = PIV_BUDGET_CASES[i]
    b, used = oraclelib.piv_budget(piv_budget_inputs(case), case[2], oraclelib.mt19937(case[0], 2 * case[1] + 8))
    assert used == GP[f"pb{i}_used"] and np.array_equal(b, GP[f"pb{i}_b"])


@pytest.mark.parametrize("i", range(len(PIV_ADJUST_CASES)))
def test_adjust_probs_golden(i):
    case = PIV_ADJUST_CASES[i]
    v, keep, n_loc, tot_norm = piv_adjust_inputs(case)
    ov, ok, on, onorm = oraclelib.adjust_probs(v, n_loc, case[3], case[2], tot_norm, keep)
    assert on == GP[f"pa{i}_n"] and onorm == GP[f"pa{i}_norm"]
    assert np.array_equal(ov, GP[f"pa{i}_v"]) and np.array_equal(ok, GP[f"pa{i}_k"])


@pytest.mark.parametrize("i", range(len(PIV_COMP_CASES)))
def test_piv_comp_parallel_golden(i):
    case = PIV_COMP_CASES[i]
    ov, ok, used = oraclelib.piv_comp(piv_comp_inputs(case), case[2], oraclelib.mt19937(case[0], 2 * case[2] + 8))
    nz = np.flatnonzero(ov)
    assert used == GP[f"pc{i}_used"] and np.array_equal(nz, GP[f"pc{i}_idx"]) and np.array_equal(ov[nz], GP[f"pc{i}_val"])
    assert np.array_equal(ok == 1, ov == 0)


@pytest.mark.parametrize("i", range(len(PIV_HBPP_CASES)))
def test_apply_hbpp_piv_golden(i):
    case = PIV_HBPP_CASES[i]
    sm = SynthMol(*case[0])
    keys, vals = hbpp_inputs(sm, case)
    ov, od, oo, used = oraclelib.OracleMol(sm).apply_hbpp_piv(keys, vals, 0.97, case[3],
                                                             oraclelib.mt19937(case[4], 12 * case[2] + 64), case[2],
                                                             4 * case[2] + 4 * case[1])
    assert used == GP[f"ph{i}_used"] and np.array_equal(od, GP[f"ph{i}_d"]) and np.array_equal(oo, GP[f"ph{i}_o"])
    assert np.allclose(ov, GP[f"ph{i}_v"], rtol=1e-12, atol=0)

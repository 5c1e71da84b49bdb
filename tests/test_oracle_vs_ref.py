"""CPU tier: pin the plain-C oracle (oracle/fries_oracle.c) against the compiled reference (oracle/_ref) --
bit for bit, on the same inputs the GPU parity tests use."""
import ctypes as C

import numpy as np
import pytest

import oraclelib
import reflib
from fries_b200.synth import SynthMol

pytestmark = pytest.mark.skipif(not reflib.available(), reason="oracle/_ref not built (needs /root/reference)")


def make_values(rng, n, kind):
    if kind == "lognormal":
        v = rng.lognormal(0, 2.5, n)
    elif kind == "uniform":
        v = rng.random(n)
    else:
        v = np.concatenate([rng.lognormal(6, 1, max(1, n // 100)), rng.lognormal(-3, 2, n - max(1, n // 100))])
        rng.shuffle(v)
    v *= rng.choice([-1.0, 1.0], n)
    v[rng.random(n) < 0.05] = 0.0
    return v


def test_hash_and_bits():
    rng = np.random.default_rng(0)
    scr = rng.integers(0, 2**32, 52, dtype=np.uint64).astype(np.uint32)
    keys = rng.integers(0, 2**52, 5000, dtype=np.uint64)
    for n_procs in (1, 3, 8):
        h, o = oraclelib.hash_keys(keys, scr, n_procs)
        rh, ro = np.zeros(keys.size, np.uint64), np.zeros(keys.size, np.int32)
        reflib.lib().ref_hash_keys(keys, keys.size, 52, scr, n_procs, rh, ro)
        assert np.array_equal(h, rh) and np.array_equal(o, ro)
    L, R = oraclelib.lib(), reflib.lib()
    for k in keys[:500]:
        k = int(k)
        a, b = (int(x) for x in rng.choice(52, 2, replace=False))
        assert L.fo_bits_between(k, a, b) == R.ref_bits_between(k, a, b)
    for n_orb, n_elec in [(4, 4), (10, 6), (22, 8), (26, 10), (31, 12)]:  # tests/test_bitstrings.cpp:40-92 shapes
        assert L.fo_gen_hf_bitstring(n_orb, n_elec) == R.ref_gen_hf_bitstring(n_orb, n_elec)


@pytest.mark.parametrize("n,budget,kind", [(1, 1, "uniform"), (7, 3, "lognormal"), (1000, 100, "lognormal"),
                                            (1000, 2000, "uniform"), (50000, 5000, "fri"), (300000, 250000, "lognormal")])
def test_find_preserve_sys_comp(n, budget, kind):
    rng = np.random.default_rng(n + budget)
    v = make_values(rng, n, kind)
    r = reflib.find_preserve(v, budget)
    o = oraclelib.find_preserve(v, budget)
    assert o[0] == r[0] and o[1] == r[1] and o[2] == r[2] and np.array_equal(o[3], r[3])
    for rn in (0.0, 0.37, 0.999999):
        rv, rk, rnorm = reflib.sys_comp(v, r[0], r[2], r[3], rn)
        ov, ok, onorm = oraclelib.sys_comp(v, [o[0]], o[2], o[3], rn)
        assert np.array_equal(ov, rv) and np.array_equal(ok, rk) and onorm[0] == rnorm


@pytest.mark.parametrize("n,n_sub,budget,jagged", [(5, 2, 4, False), (200, 8, 50, False), (3000, 11, 700, True),
                                                    (40000, 18, 9000, True), (40000, 2, 60000, False)])
def test_comp_sub(n, n_sub, budget, jagged):
    rng = np.random.default_rng(n * 7 + n_sub)
    v = rng.lognormal(0, 2, n)
    v[rng.random(n) < 0.03] = 0
    nd = np.where(rng.random(n) < 0.4, rng.integers(1, 30, n), 0).astype(np.uint32)
    sw = rng.random((n, n_sub)) ** 3
    sw[rng.random((n, n_sub)) < 0.1] = 0
    ss = None
    if jagged:
        ss = rng.integers(1, n_sub + 1, n).astype(np.uint16)
        for i in range(n):
            sw[i, ss[i]:] = 0
    tot = sw.sum(1, keepdims=True)
    tot[tot == 0] = 1
    sw = sw / tot
    cap = 4 * max(budget, n) + 64
    for rn in (0.123, 0.9):
        rv, ri = reflib.comp_sub(v, nd, sw, ss, budget, rn, cap)
        ov, oi, _, _ = oraclelib.comp_sub(v, nd, sw, ss, budget, rn, cap)
        assert np.array_equal(oi, ri) and np.array_equal(ov, rv)


@pytest.fixture(scope="module", params=[("ne", 2, False), ("ne", 2, True), ("h2o", 3, True), ("n2", 7, True)])
def mols(request):
    name, seed, frozen = request.param
    sm = SynthMol(name, seed, frozen)
    return sm, reflib.RefMol(sm), oraclelib.OracleMol(sm)


def test_mol_tables_elements_enumeration(mols):
    sm, rm, om = mols
    assert np.array_equal(om.packed_eris(), sm.eris_packed)
    rt, ot = rm.hb_tables(), om.hb_tables()
    for k in rt:
        assert np.array_equal(ot[k], rt[k]), k
    rng = np.random.default_rng(5)
    keys = np.concatenate([[sm.hf], sm.random_dets(30, rng, None)]).astype(np.uint64)
    assert np.array_equal(om.diag(keys), rm.diag(keys))
    for k in keys[:10]:
        se, de = rm.sing_ex(k), rm.doub_ex(k)
        assert np.array_equal(om.sing_ex(k), se) and np.array_equal(om.doub_ex(k), de)
        assert oraclelib.lib().fo_mol_count_singex(om.h, int(k)) == reflib.lib().ref_mol_count_singex(rm.h, int(k))
        assert np.array_equal(om.sing_el([k] * len(se), se), rm.sing_el(np.full(len(se), k, np.uint64), se))
        assert np.array_equal(om.doub_el(de[::5]), rm.doub_el(de[::5]))
        for o in de[:: max(1, len(de) // 40)]:
            for nrm in (0, 1):
                assert om.hb_wt(nrm, k, o) == pytest.approx(rm.hb_wt(nrm, k, o), rel=1e-14)


def test_mol_hb_rows(mols):
    sm, rm, om = mols
    rng = np.random.default_rng(6)
    M, ne = sm.n_orb, sm.n_elec
    for k in sm.random_dets(20, rng, None):
        k = int(k)
        occ = [i for i in range(2 * M) if (k >> i) & 1]
        vir = [i for i in range(2 * M) if i not in occ]
        cases = [(0, 0, 0, 0), (0, 1, 0, 0)]
        for o1 in range(ne):
            cases += [(1, o1, 0, 0), (3, occ[o1], 0, 0), (3, occ[o1], 1, 0)]
            if o1:
                cases.append((2, o1, 0, 0))
        for _ in range(20):
            o1i, o2i = sorted(rng.choice(ne, 2, replace=False))[::-1]
            u1 = int(rng.choice([x for x in vir if x // M == occ[o1i] // M]))
            cases += [(4, occ[o1i], occ[o2i], u1), (5, occ[o1i], occ[o2i], u1)]
        for which, a0, a1, a2 in cases:
            r, rrow = rm.hb_row(which, k, a0, a1, a2)
            o, orow = om.hb_row(which, k, a0, a1, a2)
            assert np.array_equal(orow, rrow, equal_nan=True) and (o == r or (np.isnan(o) and np.isnan(r)))


@pytest.mark.parametrize("new_hb", [0, 1])
@pytest.mark.parametrize("n_det,n_samp", [(1, 50), (500, 2000), (5000, 8000)])
def test_apply_hbpp_sys(mols, new_hb, n_det, n_samp):
    sm, rm, om = mols
    rng = np.random.default_rng(n_det + new_hb)
    keys = np.concatenate([[sm.hf], sm.random_dets(n_det - 1, rng, 0)]).astype(np.uint64) if n_det > 1 else \
        np.array([sm.hf], np.uint64)
    vals = make_values(rng, n_det, "fri")
    vals[0] = 100.0
    cap = 4 * n_samp + 4 * n_det
    for seed in (1, 2):
        uni, rv, rd, ro = rm.apply_hbpp_sys(keys, vals, 0.97, new_hb, seed, n_samp, cap)
        ov, od, oo = om.apply_hbpp_sys(keys, vals, 0.97, new_hb, uni, n_samp, cap)
        assert np.array_equal(od, rd) and np.array_equal(oo, ro)
        assert np.allclose(ov, rv, rtol=1e-13, atol=0)


def test_h_apply(mols):
    sm, rm, om = mols
    rng = np.random.default_rng(12)
    scr = rng.integers(0, 2**32, sm.n_bits, dtype=np.uint64).astype(np.uint32)
    keys = np.concatenate([[sm.hf], sm.random_dets(9, rng, 0)]).astype(np.uint64)
    vals = rng.normal(size=10)
    rk, rv = rm.h_apply(keys, vals, 1.0, -0.01, 200000, scr, scr)
    ok, ov = om.h_apply(keys, vals, 1.0, -0.01)
    ro = np.argsort(rk)
    nz = ov != 0
    assert np.array_equal(rk[ro], ok)
    assert np.allclose(ov, rv[ro], rtol=1e-12, atol=1e-13 * np.abs(rv).max())

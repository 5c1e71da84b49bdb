"""CPU tier: the device pipeline of fries_apply_hbpp_piv (apply_HBPP_piv, heat_bathPP.cpp:1014-1419), step by step on the
host.  tests/hostcheck compiles the product's per-input arithmetic for the host -- the stage providers of hbpp_prov.cuh
(the same ones the systematic stage kernels use), the group prep / fill of the "long" vector and the finalize of a sample
-- and restates the scans, the collapse and the buffer ping-pong of hbpp.cu; the pivotal compression between expand and
collapse is the oracle's piv_comp_parallel here.  With the same compression the composition must reproduce the oracle's
apply_HBPP_piv (pinned to the compiled reference, tests/test_oracle_piv.py): same samples, same draw count.  What stays
for the GPU tier (tests/test_gpu_hbpp_piv.py): launch geometry, the one-CTA scans, the resident compression."""
import numpy as np
import pytest

import oraclelib as ol
from fries_b200.synth import SynthMol
from golden_cases import make_values
from hostcheck import hc

CASES = [(("ne", 2, False), 1, 1, 50), (("ne", 2, True), 1, 300, 1000), (("ne", 2, True), 0, 300, 1000),
         (("h2o", 3, True), 0, 1000, 1500), (("h2o", 3, True), 1, 200, 4000), (("n2", 7, True), 1, 500, 800),
         (((8, 6, 0, [0, 0, 1, 2, 3, 0, 1, 2]), 4, True), 1, 40, 3000),
         (((8, 6, 0, [0, 0, 1, 2, 3, 0, 1, 2]), 4, True), 0, 40, 200)]


def host_pipeline(L, hm, sm, keys, vals, p_doub, new_hb, draws, n_samp, cap, cutoff=1e-12):
    ne, M = sm.n_elec, sm.n_orb
    n_states = max(ne, M - ne // 2, int(np.bincount(sm.symm).max()), 2)
    long_cap = cap * n_states  # HBCompressPiv::long_vec heat_bathPP.hpp:303-310
    lng = np.zeros(long_cap)
    h = L.hc_hbpiv_begin(hm, keys, vals, len(keys), p_doub, int(new_hb), cap)
    used = 0
    sizes = []
    try:
        for s in range(5):
            n_long = L.hc_hbpiv_expand(h, s, lng, long_cap)
            assert n_long != 2**64 - 1, f"stage {s} exceeds the long vector"
            if n_long:
                v, zeroed, u = ol.piv_comp(lng[:n_long], n_samp, draws[used:])
                used += u
            else:
                v, zeroed = np.zeros(1), np.zeros(1, np.uint8)
            n_out = L.hc_hbpiv_collapse(h, s, np.ascontiguousarray(v), np.ascontiguousarray(zeroed))
            assert n_out <= cap
            sizes.append((n_long, n_out))
        ov, od, oo = np.zeros(cap), np.zeros(cap, np.uint64), np.zeros((cap, 4), np.uint8)
        n = L.hc_hbpiv_finalize(h, cutoff, ov, od, oo.reshape(-1))
    finally:
        L.hc_hbpiv_end(h)
    return ov[:n].copy(), od[:n].copy(), oo[:n].copy(), used, sizes


@pytest.mark.parametrize("case,new_hb,n_det,n_samp", CASES)
def test_host_pipeline_reproduces_the_oracle(case, new_hb, n_det, n_samp):
    sm = SynthMol(*case)
    om = ol.OracleMol(sm)
    t = om.hb_tables()
    L = hc.lib()
    hm = L.hc_mol_create(sm.n_orb, sm.n_elec_total, sm.n_frz, sm.hcore.reshape(-1), sm.eris_packed, sm.eris_packed.size,
                         sm.symm, t["d_diff"], t["d_same"], t["s_tens"], float(t["s_norm"][0]), t["exch_sqrt"],
                         t["diag_sqrt"], t["exch_norms"])
    try:
        rng = np.random.default_rng(n_det + new_hb)
        keys = np.concatenate([[sm.hf], sm.random_dets(n_det - 1, rng, 0)]).astype(np.uint64) if n_det > 1 else \
            np.array([sm.hf], np.uint64)
        vals = make_values(rng, n_det, "fri")
        vals[0] = 100.0
        if n_det > 5:
            vals[3] = 0.0  # an empty input: no group in the reference, one entry of weight zero here
        cap = 4 * n_samp + 4 * n_det
        draws = ol.mt19937(3, 12 * n_samp + 64)
        ov, od, oo, oused = om.apply_hbpp_piv(keys, vals, 0.97, new_hb, draws, n_samp, cap)
        hv, hd, ho, hused, sizes = host_pipeline(L, hm, sm, keys, vals, 0.97, new_hb, draws, n_samp, cap)
    finally:
        L.hc_mol_destroy(hm)
    # multisets: with new_hb = 0 the same excitation is reached along several paths (either order of the two electrons)
    def bag(dd, orbs, vv):
        out = {}
        for d, o, v in zip(dd, orbs.tolist(), vv):
            out.setdefault((int(d), tuple(o)), []).append(float(v))
        return {k: sorted(v) for k, v in out.items()}
    ob, hb = bag(od, oo, ov), bag(hd, ho, hv)
    n_diff = sum(abs(len(ob.get(k, [])) - len(hb.get(k, []))) for k in set(ob) | set(hb))
    print(f"{case[0]} new_hb={new_hb}: {len(ov)} samples, {n_diff} differ, draws {hused} / {oused}, stages {sizes}")
    # the long vectors agree to rounding (value x (raw x 1/norm) here, (raw / norm) x value there): a sample whose
    # inclusion is decided within an ulp may move, and with it a few downstream ones
    assert n_diff <= max(2, len(ov) // 500)
    for k in set(ob) & set(hb):
        if len(ob[k]) == len(hb[k]):
            assert hb[k] == pytest.approx(ob[k], rel=1e-9)
    assert abs(hused - oused) <= 4


# ---- the same providers in the systematic pipeline (apply_HBPP_sys, heat_bathPP.cpp:686-992) ---------------------------
def host_pipeline_sys(L, hm, sm, keys, vals, p_doub, new_hb, uniforms5, n_samp, cap):
    ne, M = sm.n_elec, sm.n_orb
    cols_of = [2, ne - (1 if new_hb else 0), ne - (1 if new_hb else 0), M - ne // 2, int(np.bincount(sm.symm).max())]
    h = L.hc_hbpiv_begin(hm, keys, vals, len(keys), p_doub, int(new_hb), cap)
    try:
        for s in range(5):
            cols = cols_of[s]
            values, ndiv = np.zeros(cap), np.zeros(cap, np.uint32)
            subwts, nsub = np.zeros(cap * cols), np.zeros(cap, np.uint16)
            n = L.hc_hbsys_rows(h, s, values, ndiv, subwts, cols, nsub)
            assert n != 2**64 - 1, f"stage {s}: a sub-weight exceeds the provider's bound wmax"
            jag = nsub[:n] if (s == 4 or (s == 2 and new_hb)) else None  # where the reference passes sub_sizes
            nv, ni, _, _ = ol.comp_sub(values[:n], ndiv[:n], subwts[:n * cols].reshape(n, cols), jag, n_samp, uniforms5[s], cap)
            L.hc_hbsys_accept(h, s, np.ascontiguousarray(nv), np.ascontiguousarray(ni.reshape(-1)), len(nv))
        ov, od, oo = np.zeros(cap), np.zeros(cap, np.uint64), np.zeros((cap, 4), np.uint8)
        k = L.hc_hbpiv_finalize(h, 1e-9, ov, od, oo.reshape(-1))
    finally:
        L.hc_hbpiv_end(h)
    return ov[:k].copy(), od[:k].copy(), oo[:k].copy()


@pytest.mark.parametrize("case,new_hb,n_det,n_samp", CASES)
def test_host_sys_pipeline_reproduces_the_oracle(case, new_hb, n_det, n_samp):
    """providers + finalize of the systematic stage kernels on the host, the oracle's comp_sub in between: the oracle's
    apply_HBPP_sys sample for sample (a CPU regression harness for changes to the row generators)"""
    sm = SynthMol(*case)
    om = ol.OracleMol(sm)
    t = om.hb_tables()
    L = hc.lib()
    hm = L.hc_mol_create(sm.n_orb, sm.n_elec_total, sm.n_frz, sm.hcore.reshape(-1), sm.eris_packed, sm.eris_packed.size,
                         sm.symm, t["d_diff"], t["d_same"], t["s_tens"], float(t["s_norm"][0]), t["exch_sqrt"],
                         t["diag_sqrt"], t["exch_norms"])
    try:
        rng = np.random.default_rng(n_det + new_hb)
        keys = np.concatenate([[sm.hf], sm.random_dets(n_det - 1, rng, 0)]).astype(np.uint64) if n_det > 1 else \
            np.array([sm.hf], np.uint64)
        vals = make_values(rng, n_det, "fri")
        vals[0] = 100.0
        cap = 4 * n_samp + 4 * n_det
        u5 = rng.random(5)
        ov, od, oo = om.apply_hbpp_sys(keys, vals, 0.97, new_hb, u5, n_samp, cap)
        hv, hd, ho = host_pipeline_sys(L, hm, sm, keys, vals, 0.97, new_hb, u5, n_samp, cap)
    finally:
        L.hc_mol_destroy(hm)
    same = len(ov) == len(hv) and np.array_equal(od, hd) and np.array_equal(oo, ho)
    n_diff = 0 if same else len(set(zip(od.tolist(), map(tuple, oo.tolist()))) ^ set(zip(hd.tolist(), map(tuple, ho.tolist()))))
    print(f"{case[0]} new_hb={new_hb}: {len(ov)} samples, same order and content: {same}, set difference {n_diff}")
    # rounding of value x weight (different association than the reference) may move a sample at a tie
    assert n_diff <= max(2, len(ov) // 500)
    if same:
        assert np.allclose(np.abs(hv), np.abs(ov), rtol=1e-9, atol=0)

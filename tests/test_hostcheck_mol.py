"""CPU tier: the product's __host__ __device__ arithmetic (mol.cuh, common.cuh), compiled for the host by
tests/hostcheck, against the compiled reference (oracle/_ref).  Skipped where the reference build is absent."""
import ctypes as C

import numpy as np
import pytest

import reflib
from fries_b200.synth import SynthMol
from hostcheck import hc

pytestmark = pytest.mark.skipif(not reflib.available(), reason="oracle/_ref not built")


def occ_of(key):
    return [i for i in range(64) if (key >> i) & 1]


@pytest.fixture(scope="module", params=[("ne", 2, True), ("ne", 2, False), ("h2o", 3, True), ("n2", 7, True)])
def mols(request):
    name, seed, frozen = request.param
    sm = SynthMol(name, seed, frozen)
    rm = reflib.RefMol(sm)
    t = rm.hb_tables()
    L = hc.lib()
    h = L.hc_mol_create(sm.n_orb, sm.n_elec_total, sm.n_frz, sm.hcore.reshape(-1), sm.eris_packed, sm.eris_packed.size,
                        sm.symm, t["d_diff"], t["d_same"], t["s_tens"], float(t["s_norm"][0]), t["exch_sqrt"],
                        t["diag_sqrt"], t["exch_norms"])
    yield sm, rm, h
    L.hc_mol_destroy(h)


def test_matrix_elements_and_enumeration(mols):
    sm, rm, h = mols
    L = hc.lib()
    rng = np.random.default_rng(11)
    keys = np.concatenate([[sm.hf], sm.random_dets(40, rng, None)]).astype(np.uint64)
    ref_d = rm.diag(keys)
    for k, rd in zip(keys, ref_d):
        assert L.hc_diag(h, int(k)) == rd
    for k in keys[:12]:
        k = int(k)
        se = rm.sing_ex(k)
        buf = np.zeros((4096, 2), np.uint8)
        n = L.hc_sing_ex(h, k, buf.reshape(-1))
        assert n == len(se) and np.array_equal(buf[:n], se)
        assert L.hc_count_singex(h, k) == reflib.lib().ref_mol_count_singex(rm.h, k)
        de = rm.doub_ex(k)
        buf = np.zeros((1 << 16, 4), np.uint8)
        n = L.hc_doub_ex(h, k, buf.reshape(-1))
        assert n == len(de) and np.array_equal(buf[:n], de)
        if len(se):
            ref = rm.sing_el(np.full(len(se), k, np.uint64), se)
            for o, r in zip(se, ref):
                assert L.hc_sing_el(h, k, np.ascontiguousarray(o)) == r
        ref = rm.doub_el(de)
        for o, r in zip(de[::7], ref[::7]):
            assert L.hc_doub_el(h, np.ascontiguousarray(o)) == r
        # parities + excited determinants
        for o in se[::3]:
            kk = C.c_uint64(k)
            kr = C.c_uint64(k)
            assert L.hc_bit_op(0, C.byref(kk), np.ascontiguousarray(o)) == reflib.lib().ref_sing_det_parity(
                C.byref(kr), np.ascontiguousarray(o))
            assert kk.value == kr.value
        for o in de[::11]:
            kk = C.c_uint64(k)
            kr = C.c_uint64(k)
            assert L.hc_bit_op(1, C.byref(kk), np.ascontiguousarray(o)) == reflib.lib().ref_doub_det_parity(
                C.byref(kr), np.ascontiguousarray(o))
            assert kk.value == kr.value
            assert L.hc_bit_op(3, C.byref(C.c_uint64(k)), np.ascontiguousarray(o)) == reflib.lib().ref_doub_parity(
                k, np.ascontiguousarray(o))


def test_hb_rows_and_weights(mols):
    sm, rm, h = mols
    L = hc.lib()
    rng = np.random.default_rng(5)
    keys = np.concatenate([[sm.hf], sm.random_dets(25, rng, None)]).astype(np.uint64)
    M, ne = sm.n_orb, sm.n_elec

    def both(which, k, a0=0, a1=0, a2=0):
        rr, rrow = rm.hb_row(which, k, a0, a1, a2)
        row = np.zeros(64)
        ln = C.c_int(0)
        r = L.hc_hb_row(h, which, k, a0, a1, a2, row, C.byref(ln))
        assert ln.value == len(rrow), (which, a0, a1, a2)
        assert r == rr or (np.isnan(r) and np.isnan(rr)), (which, r, rr)
        assert np.array_equal(row[:ln.value], rrow, equal_nan=True), (which, a0, a1, a2)

    for k in keys:
        k = int(k)
        occ = occ_of(k)
        for ex in (0, 1):
            both(0, k, ex)
        for o1 in range(ne):
            both(1, k, o1)
            if o1 >= 1:
                both(2, k, o1)
            for ex in (0, 1):
                both(3, k, occ[o1], ex)
        virt = [v for v in range(2 * M) if v not in occ]
        for _ in range(30):
            o1, o2 = rng.choice(ne, 2, replace=False)
            o1, o2 = int(max(o1, o2)), int(min(o1, o2))
            u1 = int(rng.choice([v for v in virt if v // M == occ[o1] // M]))
            both(4, k, occ[o1], occ[o2], u1)
            both(5, k, occ[o1], occ[o2], u1)
        de = rm.doub_ex(k)
        for o in de[:: max(1, len(de) // 60)]:
            o = np.ascontiguousarray(o)
            for nrm in (0, 1):
                # products/quotients: the reference build contracts a*b+c into FMAs, so allow a few ulp
                    assert L.hc_hb_wt(h, nrm, k, o) == pytest.approx(rm.hb_wt(nrm, k, o), rel=1e-14)
        for ch in range(ne):
            ei, nv = C.c_uint(0), C.c_uint(0)
            L.hc_sing_counts(h, k, ch, C.byref(ei), C.byref(nv))
        for spin in (0, 1):
            o8 = np.array(occ + [255], np.uint8)
            for n in range(M - ne // 2):
                assert L.hc_find_nth_virt(o8, spin, ne, M, n) == reflib.lib().ref_find_nth_virt(
                    np.array(occ, np.uint8), spin, ne, M, n)

"""CPU tier: the deterministic H.v of one parent (h_op_offdiag, molecule.cpp:448-665) as the kernels of csrc/iter.cu build it
-- the 32 lanes' shares of the excitations (`lane_excitations`) and the arithmetic of one connection (`hv_connection`: new
determinant, sign x matrix element x value x h_fac; csrc/hv_prov.cuh), compiled for the host -- against the oracle's
h_op_offdiag: the same multiset of (determinant, value), every connection exactly once."""
import numpy as np
import pytest

import oraclelib as ol
from fries_b200.synth import SynthMol
from hostcheck import hc


@pytest.mark.parametrize("case", [("ne", 2, True), ("ne", 2, False), ("h2o", 3, True), ("n2", 7, True),
                                  ((8, 6, 0, [0, 0, 1, 2, 3, 0, 1, 2]), 4, True)])
def test_hv_of_a_parent(case):
    sm = SynthMol(*case)
    om = ol.OracleMol(sm)
    t = om.hb_tables()
    L = hc.lib()
    hm = L.hc_mol_create(sm.n_orb, sm.n_elec_total, sm.n_frz, sm.hcore.reshape(-1), sm.eris_packed, sm.eris_packed.size,
                         sm.symm, t["d_diff"], t["d_same"], t["s_tens"], float(t["s_norm"][0]), t["exch_sqrt"],
                         t["diag_sqrt"], t["exch_norms"])
    try:
        rng = np.random.default_rng(4)
        parents = np.concatenate([[sm.hf], sm.random_dets(12, rng, None)]).astype(np.uint64)
        cap = 1 << 16
        for key in parents:
            val, h_fac = float(rng.normal()), -0.01
            gk, gv = np.zeros(cap, np.uint64), np.zeros(cap)
            n = L.hc_hv_parent(hm, int(key), val, h_fac, gk, gv, cap)
            assert n != 2**64 - 1 and n <= cap  # the counting pass and the writing pass agree
            k, v = om.h_apply(np.array([key], np.uint64), np.array([val]), 0.0, h_fac)
            off = k != key  # the oracle's list starts with the diagonal element
            assert off.sum() == len(k) - 1 and n == off.sum()
            o1, o2 = np.lexsort((v[off], k[off])), np.lexsort((gv[:n], gk[:n]))
            assert np.array_equal(k[off][o1], gk[:n][o2])
            assert len(np.unique(gk[:n])) == n  # every connected determinant once
            assert np.allclose(gv[:n][o2], v[off][o1], rtol=1e-12, atol=1e-300)
    finally:
        L.hc_mol_destroy(hm)

"""CPU tier: the Hubbard-Holstein arithmetic of the kernels (fries_b200/csrc/hh_prov.cuh: hub_diag, the neighbour masks,
phonon numbers, one state's term of calc_ref_ovlp), compiled for the host by tests/hostcheck, against the oracle (which
tests/test_oracle_hh.py pins to the compiled reference).  BASELINE configs[0] (frisys_hh) runs on these pieces."""
import ctypes as C

import numpy as np
import pytest

from hostcheck import hc
from test_oracle_hh import random_states
import oraclelib


def oracle():
    L = oraclelib.lib()
    u64, u, d = C.c_uint64, C.c_uint, C.c_double
    L.fo_hub_diag.restype = u; L.fo_hub_diag.argtypes = [u64, u]
    L.fo_gen_neel_det_1D.restype = u64; L.fo_gen_neel_det_1D.argtypes = [u, u]
    L.fo_hh_neighbors.argtypes = [u64, u, u, oraclelib.u8p]
    L.fo_hh_ref_ovlp.restype = d; L.fo_hh_ref_ovlp.argtypes = [oraclelib.u64p, oraclelib.f64p, C.c_size_t, u64, u, u, u, d]
    return L


@pytest.mark.parametrize("n_sites,n_elec", [(6, 6), (4, 4), (8, 6), (10, 10), (12, 8)])
def test_hh_arithmetic_on_the_host(n_sites, n_elec):
    L, H = oracle(), hc.lib()
    ph_bits = 3
    rng = np.random.default_rng(n_sites)
    neel = L.fo_gen_neel_det_1D(n_sites, n_elec)
    for with_ph in (False, True):
        keys = np.unique(np.concatenate([[neel], random_states(rng, 400, n_sites, n_elec, ph_bits, with_ph)]).astype(np.uint64))
        vals = rng.normal(size=keys.size)
        for k in keys:
            k = int(k)
            assert H.hc_hh_hub_diag(k, n_sites) == L.fo_hub_diag(k, n_sites)
            a = np.zeros(2 * (n_elec + 1), np.uint8)
            L.fo_hh_neighbors(k, n_sites, n_elec, a)
            p, m = C.c_uint64(0), C.c_uint64(0)
            H.hc_hh_neighbors(k, n_sites, C.byref(p), C.byref(m))
            assert [i for i in range(64) if p.value >> i & 1] == list(a[1:1 + a[0]])
            assert [i for i in range(64) if m.value >> i & 1] == list(a[n_elec + 2:n_elec + 2 + a[n_elec + 1]])
            ph = sum((k >> (2 * n_sites + s * ph_bits)) & ((1 << ph_bits) - 1) for s in range(n_sites))
            assert H.hc_hh_total_ph(k, n_sites, n_elec, ph_bits) == ph
        for g in (0.0, 0.7):
            got = H.hc_hh_ref_ovlp(keys, vals, keys.size, neel, n_elec, n_sites, ph_bits, g)
            want = L.fo_hh_ref_ovlp(keys, vals, keys.size, neel, n_elec, n_sites, ph_bits, g)
            assert got == pytest.approx(want, rel=1e-12, abs=1e-12)

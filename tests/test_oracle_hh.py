"""CPU tier: Hubbard-Holstein pieces of the oracle against the compiled reference (hub_diag, Neel state, neighbour lists,
hash with phonon numbers, calc_ref_ovlp)."""
import ctypes as C

import numpy as np
import pytest

import oraclelib
import reflib

pytestmark = pytest.mark.skipif(not reflib.available(), reason="oracle/_ref not built")


def libs():
    L, R = oraclelib.lib(), C.CDLL(reflib.REF_SO)
    u64, u, d = C.c_uint64, C.c_uint, C.c_double
    L.fo_hub_diag.restype = u; L.fo_hub_diag.argtypes = [u64, u]
    L.fo_gen_neel_det_1D.restype = u64; L.fo_gen_neel_det_1D.argtypes = [u, u]
    L.fo_hh_neighbors.argtypes = [u64, u, u, oraclelib.u8p]
    L.fo_hash_hh.restype = u64; L.fo_hash_hh.argtypes = [u64, oraclelib.u32p, u, u]
    L.fo_hh_ref_ovlp.restype = d; L.fo_hh_ref_ovlp.argtypes = [oraclelib.u64p, oraclelib.f64p, C.c_size_t, u64, u, u, u, d]
    R.ref_hub_diag.restype = u; R.ref_hub_diag.argtypes = [u64, u]
    R.ref_gen_neel_det_1D.restype = u64; R.ref_gen_neel_det_1D.argtypes = [u, u, u]
    R.ref_hh_neighbors.argtypes = [u64, u, u, u, oraclelib.u8p]
    R.ref_hh_hash.restype = u64; R.ref_hh_hash.argtypes = [u64, u, u, u, oraclelib.u32p]
    R.ref_hh_ref_ovlp.restype = d; R.ref_hh_ref_ovlp.argtypes = [oraclelib.u64p, oraclelib.f64p, C.c_size_t, u64, u, u, u, d]
    return L, R


def random_states(rng, n, n_sites, n_elec, ph_bits, with_ph):
    out = []
    for _ in range(n):
        up = rng.choice(n_sites, n_elec // 2, replace=False)
        dn = rng.choice(n_sites, n_elec // 2, replace=False)
        k = sum(1 << int(i) for i in up) | sum(1 << (int(i) + n_sites) for i in dn)
        if with_ph:
            for s in range(n_sites):
                k |= int(rng.integers(0, 3 if rng.random() < 0.3 else 1)) << (2 * n_sites + s * ph_bits)
        out.append(k)
    return np.array(out, np.uint64)


@pytest.mark.parametrize("n_sites,n_elec", [(6, 6), (4, 4), (8, 6), (10, 10)])
def test_hh_pieces(n_sites, n_elec):
    L, R = libs()
    ph_bits = 3
    rng = np.random.default_rng(n_sites)
    neel = L.fo_gen_neel_det_1D(n_sites, n_elec)
    assert neel == R.ref_gen_neel_det_1D(n_sites, n_elec, ph_bits)
    scr = rng.integers(0, 2**32, 2 * n_sites, dtype=np.uint64).astype(np.uint32)
    for with_ph in (False, True):
        keys = np.concatenate([[neel], random_states(rng, 300, n_sites, n_elec, ph_bits, with_ph)]).astype(np.uint64)
        for k in keys:
            k = int(k)
            assert L.fo_hub_diag(k, n_sites) == R.ref_hub_diag(k, n_sites)
            a = np.zeros(2 * (n_elec + 1), np.uint8)
            b = np.zeros(2 * (n_elec + 1), np.uint8)
            L.fo_hh_neighbors(k, n_sites, n_elec, a)
            R.ref_hh_neighbors(k, n_sites, ph_bits, n_elec, b)
            assert a[0] == b[0] and a[n_elec + 1] == b[n_elec + 1]
            assert np.array_equal(a[1:1 + a[0]], b[1:1 + b[0]])
            assert np.array_equal(a[n_elec + 2:n_elec + 2 + a[n_elec + 1]], b[n_elec + 2:n_elec + 2 + b[n_elec + 1]])
            assert L.fo_hash_hh(k, scr, n_sites, ph_bits) == R.ref_hh_hash(k, n_sites, ph_bits, n_elec, scr)
        vals = rng.normal(size=keys.size)
        for g in (0.0, 0.7):
            o = L.fo_hh_ref_ovlp(keys, vals, keys.size, neel, n_elec, n_sites, ph_bits, g)
            r = R.ref_hh_ref_ovlp(keys, vals, keys.size, neel, n_elec, n_sites, ph_bits, g)
            assert o == pytest.approx(r, rel=1e-13, abs=1e-13)
        # every single hop from the Neel state is seen by calc_ref_ovlp
        hops = []
        for o in range(2 * n_sites):
            for dlt in (1, -1):
                t = o + dlt
                if (neel >> o) & 1 and 0 <= t < 2 * n_sites and (t // n_sites) == (o // n_sites) and not (neel >> t) & 1:
                    hops.append((neel & ~(1 << o)) | (1 << t))
        hk = np.array(hops, np.uint64)
        assert L.fo_hh_ref_ovlp(hk, np.ones(hk.size), hk.size, neel, n_elec, n_sites, ph_bits, 0.0) == hk.size

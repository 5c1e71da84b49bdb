"""CPU tier: the oracle's alias method and multinomial row compression (setup_alias / sample_alias
FRIES/compress_utils.cpp:823-897, compress_vecs_multi FRIES/vec_utils.cpp:73-127) against the reference's own code
(oracle/_ref/libfries_ref.so) and against the reference's test of the alias tables (tests/test_compression.cpp:12-61)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oraclelib as ol
import reflib

pytestmark = pytest.mark.skipif(not reflib.available(), reason="oracle/_ref/libfries_ref.so not built")


@pytest.mark.parametrize("seed,n", [(0, 1), (1, 2), (2, 7), (3, 200), (4, 5000), (5, 65535)])
def test_setup_alias_matches_reference(seed, n):
    rng = np.random.default_rng(seed)
    p = rng.random(n) * np.exp(2 * rng.standard_normal(n))
    p[rng.random(n) < 0.1] = 0
    if p.sum() == 0:
        p[0] = 1
    p /= p.sum()
    al, ap = ol.setup_alias(p)
    ral, rap = np.zeros(n, np.uint32), np.zeros(n)
    reflib.lib().ref_setup_alias(p, ral, rap, n)
    assert np.array_equal(al, ral) and np.array_equal(ap, rap)
    # the table reproduces the distribution: p_i = (own share + what the columns aliased to i give up) / n
    back = np.minimum(ap, 1.0)
    np.add.at(back, al, np.where(al != np.arange(n), 1.0 - np.minimum(ap, 1.0), 0.0))
    assert np.allclose(back / n, p, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("seed,n,n_samp", [(0, 10, 1000), (7, 300, 5000), (11, 4000, 60000)])
def test_sample_alias_matches_reference(seed, n, n_samp):
    rng = np.random.default_rng(seed)
    p = rng.random(n)
    p /= p.sum()
    al, ap = ol.setup_alias(p)
    counts = ol.sample_alias(al, ap, n_samp, ol.mt19937(seed, 2 * n_samp))
    rc = np.zeros(n, np.uint16)
    reflib.lib().ref_sample_alias(al, ap, n, rc, n_samp, seed)
    assert np.array_equal(counts, rc) and int(counts.sum()) == n_samp


def test_compress_vecs_multi_matches_reference():
    """rows 0 and 1 of a two-row DistVec through the reference's compress_vecs_multi against the oracle row by row on the
    same std::mt19937 stream; deleted = zero in both rows"""
    rng = np.random.default_rng(5)
    n_orb, half, n, m = 20, 4, 3000, 900
    n_bits = 2 * n_orb
    scr = rng.integers(0, 2**32, n_bits, dtype=np.uint64).astype(np.uint32)
    keys = set()
    while len(keys) < n:
        a = rng.choice(n_orb, half, replace=False)
        b = rng.choice(n_orb, half, replace=False)
        keys.add(sum(1 << int(x) for x in a) | sum(1 << (int(x) + n_orb) for x in b))
    keys = np.array(sorted(keys), np.uint64)
    vals = rng.standard_normal((2, n)) * np.exp(1.5 * rng.standard_normal((2, n)))
    vals[1, rng.random(n) < 0.3] = 0
    L = reflib.lib()
    rvec = L.ref_vec_create(2 * n, n, n_bits, 2 * half, 2, scr, scr)
    try:
        for row in (0, 1):
            L.ref_vec_add(rvec, keys, np.ascontiguousarray(vals[row]), np.ones(n, np.uint8), n, row, row)
        cs = L.ref_vec_curr_size(rvec)
        rk, rv = np.zeros(cs, np.uint64), np.zeros((2, cs))
        L.ref_vec_dump(rvec, rk, rv.reshape(-1), 2)
        assert cs == n
        draws = ol.mt19937(31, 8 * m)
        exp, used = rv.copy(), 0
        for row in (0, 1):
            exp[row], u = ol.compress_multi_row(rv[row], m, draws[used:])
            used += u
        assert used == 8 * m
        L.ref_vec_compress_multi(rvec, 0, 2, m, 31)
        cs2 = L.ref_vec_curr_size(rvec)
        ak, av = np.zeros(cs2, np.uint64), np.zeros((2, cs2))
        L.ref_vec_dump(rvec, ak, av.reshape(-1), 2)
        lut = {int(k): i for i, k in enumerate(rk)}
        live = 0
        for j, k in enumerate(ak):
            if k == 0 and not av[:, j].any():
                continue  # a freed slot of the reference's storage
            i = lut[int(k)]
            assert np.array_equal(av[:, j], exp[:, i]), (j, av[:, j], exp[:, i])
            live += av[:, j].any()
        assert live == np.count_nonzero(exp.any(axis=0))
        for row in (0, 1):
            assert np.isclose(np.abs(exp[row]).sum(), np.abs(rv[row]).sum(), rtol=1e-12)
    finally:
        L.ref_vec_destroy(rvec)

"""GPU tier: the reproducible merge (csrc/vec_det.cu, fries_vec_set_deterministic / FRIES_DETERMINISTIC=1).
New determinants are appended and values added in batch order -- the order of the reference's sequential
DistVec::add_elements (FRIES/vec_utils.hpp:606-641) -- so the store must equal the reference's store position by position and
bit by bit, and two runs of a driver with one seed must write identical files."""
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import reflib
from driver_utils import OURS, read_col, write_fcidump

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import fries_b200
    c = fries_b200.Context(0)
    yield c
    c.close()


def rand_dets(rng, n, n_orb, half):
    out = np.zeros(n, np.uint64)
    for i in range(n):
        a = rng.choice(n_orb, half, replace=False)
        b = rng.choice(n_orb, half, replace=False)
        out[i] = sum(1 << int(x) for x in a) | sum(1 << (int(x) + n_orb) for x in b)
    return out


@pytest.mark.skipif(not reflib.available(), reason="oracle/_ref/libfries_ref.so not built")
def test_merge_in_batch_order_equals_reference_store(ctx):
    import fries_b200
    rng = np.random.default_rng(21)
    n_orb, half = 26, 5
    n_bits = 2 * n_orb
    ps = rng.integers(0, 2**32, n_bits, dtype=np.uint64).astype(np.uint32)
    vs = rng.integers(0, 2**32, n_bits, dtype=np.uint64).astype(np.uint32)
    pool = np.unique(rand_dets(rng, 20000, n_orb, half))
    L = reflib.lib()
    rvec = L.ref_vec_create(100000, 50000, n_bits, 2 * half, 2, ps, vs)
    gvec = fries_b200.Vec(ctx, 100000, n_bits, 2 * half, 2, ps, vs)
    gvec.set_deterministic(True)
    try:
        for rnd, (origin, dest) in enumerate([(0, 0), (0, 1), (0, 1), (1, 0), (1, 1)]):
            n = 40000
            keys = rng.choice(pool, n)
            keys[: n // 50] = pool[0]            # a hot determinant: 800 additions to one element, in batch order
            vals = rng.normal(size=n)
            vals[rng.random(n) < 0.02] = 0
            ini = (rng.random(n) < (1.0 if rnd == 0 else 0.5)).astype(np.uint8)
            L.ref_vec_add(rvec, keys, vals, ini, n, origin, dest)
            gvec.add(keys, vals, ini, origin, dest)
            cs = L.ref_vec_curr_size(rvec)
            rk, rv = np.zeros(cs, np.uint64), np.zeros((2, cs))
            L.ref_vec_dump(rvec, rk, rv.reshape(-1), 2)
            gk, gv = gvec.download()
            assert np.array_equal(gk, rk), (rnd, "storage order")          # position by position
            assert np.array_equal(gv, rv), (rnd, np.abs(gv - rv).max())     # bit by bit: the same order of additions
            assert gvec.nonini_occ_add() == L.ref_vec_nonini_occ_add(rvec)
    finally:
        L.ref_vec_destroy(rvec)
        gvec.close()


def test_frisys_mol_driver_is_reproducible(tmp_path):
    """two runs with one seed: identical output files (the default merge drifts apart after a few hundred iterations)"""
    from fries_b200.synth import SynthMol
    from test_gpu_drivers import TINY
    sm = SynthMol(*TINY)
    fd = str(tmp_path / "FCIDUMP")
    write_fcidump(fd, sm, "D2")
    outs = []
    for name in ("a", "b"):
        rd = str(tmp_path / name) + "/"
        os.makedirs(rd)
        r = subprocess.run([os.path.join(OURS, "frisys_mol"), "--fcidump_path", fd, "--distribution", "HB_unnorm", "--vec_nonz", "150",
                            "--mat_nonz", "300", "--max_dets", "20000", "--epsilon", "0.05", "--target", "500", "--max_iter", "1500",
                            "--result_dir", rd, "--point_group", "D2"], env=dict(os.environ, FRIES_SEED="11", FRIES_DETERMINISTIC="1"),
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
        assert r.returncode == 0 and "Exception" not in r.stderr, r.stderr[-500:]
        outs.append({f: open(rd + f).read() for f in ("projnum.txt", "projden.txt", "S.txt", "norm.txt", "nkept.txt")})
        outs[-1]["dets"] = open(rd + "dets0.dat", "rb").read()
        outs[-1]["vals"] = open(rd + "vals0.dat", "rb").read()
    for f in outs[0]:
        assert outs[0][f] == outs[1][f], f
    assert len(read_col(str(tmp_path / "a") + "/projnum.txt")) == 1500

"""CPU tier: fries_b200/checkpoint.py -- the reference's checkpoint format and re-sharding to another rank count, checked
with the reference itself as producer and consumer: a frisys_mol run saved on 4 ranks is re-sharded to 2 ranks (and to 1),
the reference restarts from it with --load_dir on that many ranks and carries on where the first run stopped."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "mpi_shim"))
import shimrun  # noqa: E402
import oraclelib  # noqa: E402
from driver_utils import REF, read_col, write_fcidump  # noqa: E402
from fries_b200 import checkpoint  # noqa: E402
from fries_b200.synth import SynthMol  # noqa: E402

have_ref = os.path.exists(os.path.join(REF, "frisys_mol"))


def test_det_hash_matches_the_oracle():
    rng = np.random.default_rng(3)
    for n_bits in (8, 44, 52):
        scr = rng.integers(0, 2**32, n_bits, dtype=np.uint64).astype(np.uint32)
        keys = rng.integers(0, 2**n_bits, 5000, dtype=np.uint64)
        h, o = oraclelib.hash_keys(keys, scr, 7)
        assert np.array_equal(checkpoint.det_hash(keys, scr), h)
        assert np.array_equal(checkpoint.owners(keys, scr, 7), o)


def run_ref(n, fd, rd, load=None, iters=300, seed=3):
    cmd = [os.path.join(REF, "frisys_mol"), "--fcidump_path", fd, "--distribution", "HB_unnorm", "--vec_nonz", "150",
           "--mat_nonz", "300", "--max_dets", "20000", "--epsilon", "0.05", "--target", "500", "--max_iter", str(iters),
           "--result_dir", rd, "--point_group", "D2"] + (["--load_dir", load] if load else [])
    err = open(rd + "stderr.txt", "w+")
    rc, _ = shimrun.run(n, cmd, 8 << 20, timeout=300, env=dict(os.environ, FRIES_SEED=str(seed)), stderr=err,
                        stamp=re.compile(r"^(\d+), en est: .* norm: (.*)$".replace("(.*)$", ".*$")))
    err.seek(0)
    assert rc == 0 and "Exception" not in err.read()
    return len(shimrun.run.stamps)


@pytest.mark.skipif(not have_ref, reason="oracle/_ref drivers not built")
def test_reshard_a_reference_checkpoint(tmp_path):
    sm = SynthMol((8, 6, 0, [0, 0, 1, 2, 3, 0, 1, 2]), 4)
    fd = str(tmp_path / "FCIDUMP")
    write_fcidump(fd, sm, "D2")
    first = str(tmp_path / "first") + "/"
    os.makedirs(first)
    assert run_ref(4, fd, first) == 300
    n_bits = 2 * sm.n_orb
    assert checkpoint.n_saved_ranks(first) == 4
    scr = np.fromfile(first + "hash.dat", np.uint32)
    before = {}
    for r in range(4):
        k, v = checkpoint.read_rank(first, r, n_bits, 2)
        assert np.all(checkpoint.owners(k, scr, 4) == r)        # the reference's own partition, reproduced by det_hash
        before.update({int(a): (b, c) for a, b, c in zip(k, v[0], v[1]) if b != 0 or c != 0})  # freed slots hold zeros
    norm_before = read_col(first + "norm.txt")[-1]
    for n_new in (2, 1):
        dst = str(tmp_path / f"to{n_new}") + "/"
        info = checkpoint.reshard(first, dst, n_bits, 2, n_new)
        assert info["ranks_in"] == 4 and info["determinants"] == len(before) and sum(info["per_rank"]) == len(before)
        after = {}
        for r in range(n_new):
            k, v = checkpoint.read_rank(dst, r, n_bits, 2)
            assert np.all(checkpoint.owners(k, scr, n_new) == r)
            after.update({int(a): (b, c) for a, b, c in zip(k, v[0], v[1])})
        assert after == before
        assert np.array_equal(np.fromfile(dst + "hash.dat", np.uint32), scr)
        # the reference restarts from the re-sharded files on n_new ranks and continues: the one-norm it reports after its
        # first ten iterations is the one the first run ended with, up to ten iterations of drift
        again = str(tmp_path / f"again{n_new}") + "/"
        os.makedirs(again)
        assert run_ref(n_new, fd, again, load=dst, iters=40, seed=9) == 40
        norm_after = read_col(again + "norm.txt")[0]
        assert norm_after == pytest.approx(norm_before, rel=0.1), (norm_before, norm_after)
        num, den = read_col(again + "projnum.txt"), read_col(again + "projden.txt")
        assert np.all(np.isfinite(num / den)) and abs(num[-1] / den[-1] + 0.3775) < 0.1

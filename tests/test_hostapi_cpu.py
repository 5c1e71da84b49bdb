"""CPU tier: the pure-host pieces of the C++ mirror of the reference's interface (fries_b200/host/fries_host.hpp) --
Matrix<T> (ndarr.hpp), find_bits (math_utils.c:62-98), HashTable::hash_fxn (det_hash.hpp:160-170) as used by
DistVec::idx_to_hash / idx_to_proc -- against the oracle.  The accessors that touch the device store are exercised on the
GPU by fries_b200/host/distvec_check.cpp (tests/test_zz_gpu_hostapi.py)."""
import os
import subprocess

import numpy as np

import oraclelib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_host_hash_and_bits(tmp_path):
    exe = str(tmp_path / "hostapi_check")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wno-unused-function", os.path.join(HERE, "hostcheck", "hostapi_check.cpp"),
                           "-o", exe, "-L" + os.path.join(ROOT, "fries_b200"), "-lfries_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "fries_b200")])
    rng = np.random.default_rng(5)
    for n_bits in (44, 52, 30):
        scr = rng.integers(0, 2**32, n_bits, dtype=np.uint64).astype(np.uint32)
        keys = rng.integers(0, 2**n_bits, 2000, dtype=np.uint64)
        txt = f"{n_bits} 8 {n_bits} " + " ".join(map(str, scr)) + f" {keys.size} " + " ".join(map(str, keys)) + "\n"
        r = subprocess.run([exe], input=txt, stdout=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.returncode
        got = np.array([[int(x) for x in ln.split()] for ln in r.stdout.splitlines()], dtype=object)
        for n_procs in (1, 3, 8):
            h, o = oraclelib.hash_keys(keys, scr, n_procs)
            assert [int(x) for x in got[:, 0]] == [int(x) for x in h]
            assert [int(x) % n_procs for x in got[:, 0]] == [int(x) for x in o]
        assert [int(x) for x in got[:, 1]] == [bin(int(k)).count("1") for k in keys]

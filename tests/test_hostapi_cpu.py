"""CPU tier: the pure-host pieces of the C++ mirror of the reference's interface (fries_b200/host/fries_host.hpp) --
Matrix<T> (ndarr.hpp), find_bits (math_utils.c:62-98), HashTable::hash_fxn (det_hash.hpp:160-170) as used by
DistVec::idx_to_hash / idx_to_proc -- against the oracle.  The accessors that touch the device store are exercised on the
GPU by fries_b200/host/distvec_check.cpp (tests/test_gpu_hostapi.py)."""
import os
import subprocess

import numpy as np

import oraclelib
import reflib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def build(tmp_path):
    exe = str(tmp_path / "hostapi_check")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wno-unused-function", os.path.join(HERE, "hostcheck", "hostapi_check.cpp"),
                           "-o", exe, "-L" + os.path.join(ROOT, "fries_b200"), "-lfries_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "fries_b200")])
    return exe


def test_host_hash_and_bits(tmp_path):
    exe = build(tmp_path)
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


def test_host_bit_string_functions(tmp_path):
    """det_store.h / math_utils.h / fci_utils.h twins on byte strings against the compiled reference (oracle/_ref) where it
    is built, else against the oracle: bits_between, excite_sign, sing/doub_det_parity, sing/doub_parity, sing/doub_det,
    flip_spins, read_bit, print_str, find_diff_bits"""
    exe = build(tmp_path)
    rng = np.random.default_rng(9)
    L = reflib.lib() if reflib.available() else None
    O = oraclelib.lib()
    for n_orb in (10, 22, 26, 31):
        recs = []
        for _ in range(400):
            k = int(rng.integers(0, 2**(2 * n_orb), dtype=np.uint64))
            occ = [i for i in range(2 * n_orb) if k >> i & 1]
            virt = [i for i in range(2 * n_orb) if not k >> i & 1]
            if len(occ) < 2 or len(virt) < 2:
                continue
            o = sorted(int(x) for x in rng.choice(occ, 2, replace=False))
            v = sorted(int(x) for x in rng.choice(virt, 2, replace=False))
            a, b = (int(x) for x in rng.choice(2 * n_orb, 2, replace=False))
            recs.append((k, o[0], o[1], v[0], v[1], a, b))
        txt = f"{n_orb} {len(recs)}\n" + "\n".join(" ".join(map(str, r)) for r in recs) + "\n"
        r = subprocess.run([exe, "bits"], input=txt, stdout=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.returncode
        rows = [[int(x) for x in ln.split()] for ln in r.stdout.splitlines()]
        assert len(rows) == len(recs)
        half = (1 << n_orb) - 1
        for (k, o0, o1, v2, v3, a, b), got in zip(recs, rows):
            nb = O.fo_bits_between(k, a, b)
            so, do = np.array([o0, v2], np.uint8), np.array([o0, o1, v2, v3], np.uint8)
            k_s = (k & ~(1 << o0)) | (1 << v2)
            k_d = (k & ~((1 << o0) | (1 << o1))) | (1 << v2) | (1 << v3)
            flipped = ((k & half) << n_orb) | (k >> n_orb & half)
            if L is not None:
                import ctypes as C
                assert nb == L.ref_bits_between(k, a, b)
                kk = C.c_uint64(k)
                s1 = L.ref_sing_det_parity(C.byref(kk), so)
                assert kk.value == k_s
                kk = C.c_uint64(k)
                s2 = L.ref_doub_det_parity(C.byref(kk), do)
                assert kk.value == k_d
                p1, p2 = L.ref_sing_parity(k, so), L.ref_doub_parity(k, do)
            else:
                inner = lambda key, x, y: bin(key & ((1 << max(x, y)) - 1) & ~((2 << min(x, y)) - 1)).count("1")
                s1 = p1 = -1 if inner(k & ~(1 << o0), o0, v2) & 1 else 1
                kz = k & ~((1 << o0) | (1 << o1))
                s2 = p2 = (-1 if inner(kz, v2, o0) & 1 else 1) * (-1 if inner(kz, v3, o1) & 1 else 1)
            assert got == [nb, s1, k_s, s2, k_d, p1, p2, flipped, -1 if nb & 1 else 1], (k, o0, o1, v2, v3, a, b, got)


def test_host_input_parsers(tmp_path):
    """parse_fcidump (io_utils.cpp:17-96 + convert_symm :189-239) and parse_hf_input (:98-187, legacy directory with frozen
    core) of the C++ mirror: the files the drivers read, written from a synthetic molecule, come back as the arrays the
    library is created from (hcore, SymmERIs-packed integrals, irreps) -- bit for bit"""
    from driver_utils import write_fcidump, write_hf_dir
    from fries_b200.synth import SynthMol
    exe = build(tmp_path)

    def parse(args):
        r = subprocess.run([exe, "parse"] + args, stdout=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stdout[-300:]
        head, symm, hcore, eris = r.stdout.splitlines()[:4]
        h = head.split()
        return ([int(x) for x in h[:3]], float.fromhex(h[3]), float.fromhex(h[4]), np.array([int(x) for x in symm.split()], np.uint8),
                np.array([float.fromhex(x) for x in hcore.split()]), np.array([float.fromhex(x) for x in eris.split()]))

    sm = SynthMol((8, 6, 0, [0, 0, 1, 2, 3, 0, 1, 2]), 4)
    fd = str(tmp_path / "FCIDUMP")
    write_fcidump(fd, sm, "D2")
    dims, _, _, symm, hcore, eris = parse(["fcidump", fd, "D2"])
    assert dims == [sm.n_orb, sm.n_elec_total, 0]
    assert np.array_equal(symm, sm.symm) and np.array_equal(hcore, sm.hcore.reshape(-1)) and np.array_equal(eris, sm.eris_packed)

    smf = SynthMol("ne", 2, True)  # two frozen electrons
    assert smf.n_frz == 2
    d = str(tmp_path / "hf") + "/"
    write_hf_dir(d, smf, 0.001, -128.5)
    dims, eps, hf_en, symm, hcore, eris = parse(["hf", d])
    assert dims == [smf.n_orb, smf.n_elec_total, smf.n_frz] and eps == 0.001 and hf_en == -128.5
    assert np.array_equal(symm, smf.symm) and np.array_equal(hcore, smf.hcore.reshape(-1)) and np.array_equal(eris, smf.eris_packed)


def test_host_hubbard_input(tmp_path):
    """parse_hh_input (io_utils.cpp:320-408; the key order of examples/hubbard_params.txt) and gen_neel_det_1D
    (hub_holstein.cpp:139-171) against the oracle; a file with a missing key is refused with the reference's message"""
    import ctypes
    exe = build(tmp_path)
    neel = oraclelib.lib().fo_gen_neel_det_1D
    neel.restype, neel.argtypes = ctypes.c_uint64, [ctypes.c_uint, ctypes.c_uint]
    f = str(tmp_path / "hubbard_params.txt")
    for n_elec, lat_len in ((6, 6), (4, 10), (10, 10), (2, 4)):
        open(f, "w").write(f"n_elec\n{n_elec}\nlat_len\n{lat_len}\nn_dim\n1\neps\n0.001\nU\n2\nomega\n0.5\ng\n0.25\ngs_energy\n-3.98791841486987\n")
        r = subprocess.run([exe, "parse", "hh", f], stdout=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stdout
        t = r.stdout.split()
        assert [int(x) for x in t[:3]] == [n_elec, lat_len, 1]
        assert [float.fromhex(x) for x in t[3:8]] == [0.001, 2.0, 0.5, 0.25, -3.98791841486987]
        assert int(t[8]) == neel(lat_len, n_elec)
    open(f, "w").write("n_elec\n6\nlat_len\n6\nn_dim\n1\neps\n0.001\nomega\n0\n")
    r = subprocess.run([exe, "parse", "hh", f], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 21 and "electron interaction parameter (U)" in r.stdout


def test_host_files_and_command_line(tmp_path):
    """load_vec_txt (io_utils.cpp:410-482), hash.dat (:589-619), load_last_line (:636-663) and the argparse semantics of the
    drivers (Ext_Libs/argparse.hpp:347-393: --key value, unknown flag = warning, missing required flag = exit(-1))"""
    exe = build(tmp_path)
    d = str(tmp_path) + "/"
    rng = np.random.default_rng(3)
    dets = rng.integers(1, 2**44, 50, dtype=np.uint64)
    vals = rng.normal(size=52)  # two values too many: truncated to the determinants, with a warning
    open(d + "ini_dets", "w").write("\n".join(str(int(x)) for x in dets) + "\n")
    open(d + "ini_vals", "w").write("\n".join(repr(float(x)) for x in vals) + "\n")
    scr = rng.integers(0, 2**32, 44, dtype=np.uint64).astype(np.uint32)
    scr.tofile(d + "hash.dat")
    open(d + "S.txt", "w").write("0.0\n-0.25\n-0.3125\n\n")
    r = subprocess.run([exe, "files", d], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert int(lines[0]) == 50 and "fewer determinants" in r.stderr
    got = [ln.split() for ln in lines[1:51]]
    assert [int(g[0]) for g in got] == [int(x) for x in dets]
    assert [float.fromhex(g[1]) for g in got] == [float(x) for x in vals[:50]]
    assert [int(x) for x in lines[51].split()] == [int(x) for x in scr]
    assert np.array_equal(np.fromfile(d + "copy_hash.dat", np.uint32), scr)
    assert lines[52].split()[0] == "1" and float.fromhex(lines[52].split()[1]) == -0.3125 and lines[53] == "0"
    # command line
    r = subprocess.run([exe, "args", "--fcidump_path", "x/FCIDUMP", "--vec_nonz", "1000", "--target=-5", "--bogus", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0 and r.stdout.split() == ["x/FCIDUMP", "1000", "-5", "./"]
    assert "unrecognised commandline argument: bogus" in r.stderr
    r = subprocess.run([exe, "args", "--vec_nonz", "1000"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 255 and "Argument missing: --fcidump_path" in r.stderr


def test_host_seed_sys(tmp_path):
    """seed_sys (compress_utils.cpp:107-127) of the C++ mirror against the oracle for every rank of 1, 3 and 8"""
    import ctypes
    exe = build(tmp_path)
    f = oraclelib.lib().fo_seed_sys
    f.restype = ctypes.c_double
    f.argtypes = [oraclelib.f64p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_uint]
    rng = np.random.default_rng(2)
    recs, want = [], []
    for n_procs in (1, 3, 8):
        for _ in range(20):
            norms = rng.lognormal(0, 2, n_procs)
            n_samp, u = int(rng.integers(1, 5000)), float(rng.random())
            for rank in range(n_procs):
                recs.append(f"{n_procs} {rank} {n_samp} {u!r} " + " ".join(repr(float(x)) for x in norms))
                rn = ctypes.c_double(u)
                lb = f(np.ascontiguousarray(norms), n_procs, rank, ctypes.byref(rn), n_samp)
                want.append((lb, rn.value))
    r = subprocess.run([exe, "seed"], input="\n".join(recs) + "\n", stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0
    got = [tuple(float.fromhex(x) for x in ln.split()) for ln in r.stdout.splitlines()]
    assert got == want

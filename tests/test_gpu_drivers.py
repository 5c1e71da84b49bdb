"""GPU tier: the C++ drivers (fries_b200/host/bin) against the reference's own drivers (oracle/_ref) on the same
input files: same command line, same output files; deterministic runs agree to the printed precision, stochastic
runs agree within combined error bars and with the exact ground state."""
import os

import numpy as np
import pytest

import oraclelib
from driver_utils import OURS, REF, exact_ground_state, read_col, run, write_fcidump, write_hf_dir, write_vec
from fries_b200.synth import SynthMol

pytestmark = pytest.mark.gpu
TINY = ((8, 6, 0, [0, 0, 1, 2, 3, 0, 1, 2]), 4)
have_ref = os.path.exists(os.path.join(REF, "frisys_mol"))


@pytest.fixture(scope="module")
def tiny():
    sm = SynthMol(*TINY)
    om = oraclelib.OracleMol(sm)
    e_corr, e_hf, n = exact_ground_state(sm, om)
    return sm, om, e_corr, e_hf, n


def blocked_ratio(num, den, burn, block=50):
    num, den = num[burn:], den[burn:]
    nb = len(num) // block
    r = np.array([num[i * block:(i + 1) * block].sum() / den[i * block:(i + 1) * block].sum() for i in range(nb)])
    return num.sum() / den.sum(), r.std(ddof=1) / np.sqrt(nb)


@pytest.mark.skipif(not have_ref, reason="oracle/_ref drivers not built")
def test_frifull_mol_matches_reference_when_compression_is_identity(tiny, tmp_path):
    sm, om, e_corr, e_hf, n = tiny
    d = str(tmp_path / "hf") + "/"
    write_hf_dir(d, sm, 0.05, float(e_hf))
    outs = {}
    for name, exe in (("ours", os.path.join(OURS, "frifull_mol")), ("ref", os.path.join(REF, "frifull_mol"))):
        rd = str(tmp_path / name) + "/"
        os.makedirs(rd)
        r = run(exe, ["--hf_path", d, "--vec_nonz", 4 * n, "--max_dets", 8 * n, "--max_iter", 150, "--target", 110,
                      "--result_dir", rd])
        assert "Exception" not in r.stderr, r.stderr[-500:]
        outs[name] = {f: read_col(rd + f) for f in ("projnum.txt", "projden.txt", "S.txt", "norm.txt")}
        assert os.path.exists(rd + "dets0.dat") and os.path.exists(rd + "vals0.dat") and os.path.exists(rd + "hash.dat")
    for f in ("projnum.txt", "projden.txt", "S.txt", "norm.txt"):
        a, b = outs["ours"][f], outs["ref"][f]
        assert a.shape == b.shape, f
        assert np.allclose(a, b, rtol=2e-5, atol=1e-6), (f, a[:5], b[:5])  # files carry 6 significant digits
    # deterministic power iteration: the projected energy approaches the exact correlation energy from above
    en = outs["ours"]["projnum.txt"][-1] / outs["ours"]["projden.txt"][-1]
    assert e_corr - 1e-9 < en < 0
    assert len(outs["ours"]["S.txt"]) == 15 and np.any(outs["ours"]["S.txt"] != 0)  # the shift engaged


def test_frisys_mol_driver_energy_and_files(tiny, tmp_path):
    sm, om, e_corr, e_hf, n = tiny
    fd = str(tmp_path / "FCIDUMP")
    write_fcidump(fd, sm, "D2")
    res = {}
    exes = [("ours", os.path.join(OURS, "frisys_mol"))]
    if have_ref:
        exes.append(("ref", os.path.join(REF, "frisys_mol")))
    n_it = 6000
    for name, exe in exes:
        rd = str(tmp_path / name) + "/"
        os.makedirs(rd)
        r = run(exe, ["--fcidump_path", fd, "--distribution", "HB_unnorm", "--vec_nonz", 150, "--mat_nonz", 300, "--max_dets",
                      20000, "--epsilon", 0.05, "--target", 500, "--max_iter", n_it, "--result_dir", rd, "--point_group",
                      "D2"], seed=3 if name == "ours" else 4)
        assert "Exception" not in r.stderr, r.stderr[-500:]
        files = sorted(os.listdir(rd))
        for f in ("S.txt", "dense.txt", "dets0.dat", "hash.dat", "nini.txt", "nkept.txt", "norm.txt", "params.txt",
                  "projden.txt", "projnum.txt", "vals0.dat"):
            assert f in files, (name, f, files)
        num, den = read_col(rd + "projnum.txt"), read_col(rd + "projden.txt")
        assert len(num) == n_it and len(read_col(rd + "S.txt")) == n_it // 10 and len(read_col(rd + "nkept.txt")) == n_it
        assert os.path.getsize(rd + "hash.dat") == 4 * 2 * sm.n_orb
        nd = os.path.getsize(rd + "dets0.dat") // ((2 * sm.n_orb + 7) // 8)
        assert os.path.getsize(rd + "vals0.dat") == nd * 2 * 8
        res[name] = blocked_ratio(num, den, burn=1000)
        assert r.stdout.splitlines()[0].startswith("seed on process 0 is")
        it_lines = [ln for ln in r.stdout.splitlines() if ", en est: " in ln]
        assert len(it_lines) == n_it and ", shift: " in it_lines[-1] and ", norm: " in it_lines[-1]
        assert r.stdout.splitlines()[-1].startswith("Total additions to nonzero:")
    e, s = res["ours"]
    print("exact", e_corr, "ours", res["ours"], "ref", res.get("ref"))
    assert abs(e - e_corr) < 5 * s + 2e-3 * abs(e_corr) + 2e-4, (e, s, e_corr)
    if "ref" in res:
        er, sr = res["ref"]
        assert abs(e - er) < 5 * (s + sr) + 2e-4, (e, s, er, sr)


def test_frisys_mol_legacy_hf_path(tiny, tmp_path):
    """the command line of the reference's examples/run_neon.sh (legacy --hf_path directory, no --distribution, no
    --epsilon) against the FCIDUMP form of the same molecule with the same seed: the same Hamiltonian, imaginary time step
    and distribution, hence the same first iterations and statistically the same energy (runs are not reproducible
    element by element: the storage order of merged determinants depends on the order the atomics resolve)"""
    sm, om, e_corr, e_hf, n = tiny
    fd = str(tmp_path / "FCIDUMP")
    write_fcidump(fd, sm, "D2")
    d = str(tmp_path / "hf") + "/"
    write_hf_dir(d, sm, 0.05, float(e_hf))
    n_it = 3000
    common = ["--vec_nonz", 150, "--mat_nonz", 300, "--max_dets", 20000, "--target", 500, "--max_iter", n_it]
    out = {}
    for name, extra in (("fcidump", ["--fcidump_path", fd, "--distribution", "HB_unnorm", "--epsilon", 0.05, "--point_group",
                                     "D2"]), ("legacy", ["--hf_path", d])):
        rd = str(tmp_path / name) + "/"
        os.makedirs(rd)
        r = run(os.path.join(OURS, "frisys_mol"), extra + common + ["--result_dir", rd], seed=11)
        assert "Exception" not in r.stderr and r.returncode == 0, r.stderr[-500:]
        out[name] = (read_col(rd + "projnum.txt"), read_col(rd + "projden.txt"))
        assert ("HF path: " if name == "legacy" else "FCIDUMP path: ") in open(rd + "params.txt").read()
    for a, b in zip(out["fcidump"], out["legacy"]):
        assert len(a) == n_it and len(b) == n_it
        assert np.allclose(a[:1], b[:1], rtol=1e-5, atol=1e-7)  # one parent, everything preserved: deterministic
    (e1, s1), (e2, s2) = blocked_ratio(*out["fcidump"], burn=800), blocked_ratio(*out["legacy"], burn=800)
    print("fcidump", e1, s1, "legacy", e2, s2, "exact", e_corr)
    assert abs(e1 - e2) < 5 * (s1 + s2) + 2e-4, (e1, s1, e2, s2)
    assert abs(e2 - e_corr) < 5 * s2 + 2e-3 * abs(e_corr) + 2e-4, (e2, s2, e_corr)


def test_frisys_mol_semistochastic_det_space(tiny, tmp_path):
    """--det_space (frisys_mol.cpp:233-252,347-401,479-485): the determinants of the file form a
    dense subspace whose columns of H are applied exactly.  Same files and statistics as the reference's driver."""
    sm, om, e_corr, e_hf, n = tiny
    fd = str(tmp_path / "FCIDUMP")
    write_fcidump(fd, sm, "D2")
    # dense subspace: the Hartree-Fock determinant and its 24 most strongly coupled connections
    hf = np.array([sm.hf], np.uint64)
    k, v = om.h_apply(hf, np.ones(1), 0.0, 1.0)
    order = np.argsort(-np.abs(v), kind="stable")
    dense = [int(sm.hf)] + [int(x) for x in k[order] if int(x) != int(sm.hf)][:24]
    dp = str(tmp_path / "dense_dets.txt")
    open(dp, "w").write("\n".join(str(d) for d in dense) + "\n")
    # start from HF plus all of its connections, so that the stochastic part is not empty in the first iteration (the
    # reference's driver does not survive an empty one)
    ip = str(tmp_path / "ini_")
    others = [(int(a), float(b)) for a, b in zip(k, v) if int(a) != int(sm.hf)]
    write_vec(ip, [int(sm.hf)] + [a for a, _ in others], [100.0] + [-20.0 * b for _, b in others])
    res = {}
    exes = [("ours", os.path.join(OURS, "frisys_mol"))]
    if have_ref:
        exes.append(("ref", os.path.join(REF, "frisys_mol")))
    n_it = 5000
    for name, exe in exes:
        rd = str(tmp_path / name) + "/"
        os.makedirs(rd)
        # mat_nonz = 3000: the reference sizes the scratch of its dense-H set-up by spawn_length = 4 mat_nonz
        r = run(exe, ["--fcidump_path", fd, "--distribution", "HB_unnorm", "--vec_nonz", 150, "--mat_nonz", 3000, "--max_dets",
                      20000, "--epsilon", 0.05, "--target", 500, "--max_iter", n_it, "--result_dir", rd, "--point_group",
                      "D2", "--det_space", dp, "--ini_vec", ip], seed=7 if name == "ours" else 8)
        assert "Exception" not in r.stderr, (name, r.stderr[-500:])
        assert open(rd + "dense.txt").read().split(",")[0].strip() == str(len(dense)), open(rd + "dense.txt").read()
        dense_h = [ln for ln in r.stdout.splitlines() if ln.startswith("Elements in dense H:")]
        assert dense_h and int(dense_h[0].split(":")[1]) > 0
        res[name + "_dense_h"] = int(dense_h[0].split(":")[1])
        num, den = read_col(rd + "projnum.txt"), read_col(rd + "projden.txt")
        assert len(num) == n_it, (name, r.stderr[-600:], r.stdout[-600:])
        res[name] = blocked_ratio(num, den, burn=1000)
        # the dense determinants are the first stored ones and survive every compression
        nb = (2 * sm.n_orb + 7) // 8
        raw = np.fromfile(rd + "dets0.dat", dtype=np.uint8).reshape(-1, nb)
        first = [int.from_bytes(bytes(row), "little") for row in raw[:len(dense)]]
        assert sorted(first) == sorted(dense), name
    e, s = res["ours"]
    print("semi-stochastic: exact", e_corr, "ours", res["ours"], "ref", res.get("ref"))
    assert abs(e - e_corr) < 5 * s + 2e-3 * abs(e_corr) + 2e-4, (e, s, e_corr)
    if "ref" in res:
        er, sr = res["ref"]
        assert abs(e - er) < 5 * (s + sr) + 2e-4, (e, s, er, sr)
        assert res["ours_dense_h"] == res["ref_dense_h"]


# ---- config[0]: Hubbard model, frisys_hh (examples/run_hubbard.sh sizes; parameter file with the keys HEAD's parser wants) ----
HUBBARD_PARAMS = "n_elec\n6\nlat_len\n6\nn_dim\n1\neps\n0.01\nU\n2\nomega\n0\ng\n0\ngs_energy\n-3.98791841486987\n"
HUBBARD_EXACT = -4.546313794436 + 3.98791841486987  # exact E0 (400 determinants) minus gs_energy, SURVEY.md section 6


def test_frisys_hh_driver_energy_and_files(tmp_path):
    pf = str(tmp_path / "hubbard_params.txt")
    open(pf, "w").write(HUBBARD_PARAMS)
    res = {}
    exes = [("ours", os.path.join(OURS, "frisys_hh"))]
    if os.path.exists(os.path.join(REF, "frisys_hh")):
        exes.append(("ref", os.path.join(REF, "frisys_hh")))
    n_it = 6000
    for name, exe in exes:
        rd = str(tmp_path / name) + "/"
        os.makedirs(rd)
        r = run(exe, ["--params_path", pf, "--target", 1000, "--max_dets", 1000, "--vec_nonz", 200, "--max_iter", n_it,
                      "--result_dir", rd], seed=5 if name == "ours" else 6)
        assert "Exception" not in r.stderr, r.stderr[-500:]
        files = sorted(os.listdir(rd))
        for f in ("S.txt", "dets0.dat", "hash.dat", "norm.txt", "params.txt", "projden.txt", "projnum.txt", "vals0.dat"):
            assert f in files, (name, f, files)
        num, den = read_col(rd + "projnum.txt"), read_col(rd + "projden.txt")
        assert len(num) == n_it and len(read_col(rd + "S.txt")) == n_it // 10
        assert os.path.getsize(rd + "hash.dat") == 4 * 12
        nd = os.path.getsize(rd + "dets0.dat") // 4  # 30 bits -> 4 bytes per state
        assert os.path.getsize(rd + "vals0.dat") == nd * 2 * 8 and 0 < nd <= 400
        it_lines = [ln for ln in r.stdout.splitlines() if ", en est: " in ln]
        assert len(it_lines) == n_it and it_lines[-1].split(",")[1].startswith(" norm: ") and ", n_neel: " in it_lines[-1]
        res[name] = blocked_ratio(num, den, burn=2000)
    e, s = res["ours"]
    print("hubbard exact", HUBBARD_EXACT, "ours", res["ours"], "ref", res.get("ref"))
    assert abs(e - HUBBARD_EXACT) < 5 * s + 3e-3, (e, s, HUBBARD_EXACT)
    if "ref" in res:
        er, sr = res["ref"]
        assert abs(e - er) < 5 * (s + sr) + 1e-3, (e, s, er, sr)


# ---- frifull_hh (FRIES_bin/frifull_hh.cpp): no matrix compression.  With a vector budget above the number of states that
# occur the run is a deterministic power iteration: our driver must reproduce the reference driver's files to their precision.
HOLSTEIN_PARAMS = "n_elec\n4\nlat_len\n4\nn_dim\n1\neps\n0.01\nU\n2\nomega\n1.0\ng\n0.5\ngs_energy\n-2.0\n"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "frifull_hh")), reason="oracle/_ref/frifull_hh not built")
@pytest.mark.parametrize("case", ["hubbard", "holstein"])
def test_frifull_hh_matches_reference(tmp_path, case):
    pf = str(tmp_path / "params.txt")
    if case == "hubbard":   # 400 determinants, 300 iterations with the shift engaged
        open(pf, "w").write(HUBBARD_PARAMS)
        opts, n_it = ["--target", 1000, "--max_dets", 2000, "--vec_nonz", 1000], 300
    else:                   # electron-phonon coupling: phonon creation / annihilation moves; the space grows every iteration
        open(pf, "w").write(HOLSTEIN_PARAMS)
        opts, n_it = ["--target", 0, "--max_dets", 400000, "--vec_nonz", 300000], 6
    outs = {}
    for name, exe in (("ours", os.path.join(OURS, "frifull_hh")), ("ref", os.path.join(REF, "frifull_hh"))):
        rd = str(tmp_path / name) + "/"
        os.makedirs(rd)
        r = run(exe, ["--params_path", pf] + opts + ["--max_iter", n_it, "--result_dir", rd], seed=3)
        assert "Exception" not in r.stderr, r.stderr[-500:]
        outs[name] = {f: read_col(rd + f) for f in ("projnum.txt", "projden.txt", "S.txt", "norm.txt")}
        for f in ("dets0.dat", "vals0.dat", "hash.dat", "params.txt"):
            assert os.path.exists(rd + f), (name, f)
        lines = [ln for ln in r.stdout.splitlines() if ", en est: " in ln]
        assert len(lines) == n_it and ", n_neel: " in lines[-1]
    for f in ("projnum.txt", "projden.txt", "S.txt", "norm.txt"):
        a, b = outs["ours"][f], outs["ref"][f]
        assert a.shape == b.shape, (f, a.shape, b.shape)
        assert np.allclose(a, b, rtol=2e-5, atol=1e-6), (f, a[:5], b[:5])
    if case == "hubbard":
        en = outs["ours"]["projnum.txt"][-1] / outs["ours"]["projden.txt"][-1]
        assert HUBBARD_EXACT - 1e-6 < en < -0.3   # power iteration from the Neel state: above E0, already most of the way

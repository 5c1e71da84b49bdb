"""GPU tier, >= 2 GPUs: tests/multi_gpu_check.py under torchrun (one rank per GPU): in-kernel exchange ping-pong,
distributed find_preserve / sys_comp against the single-rank oracle, distributed piv_comp_parallel against the oracle's chain, 30 frisys_mol iterations on both spawn routes with
the ownership invariant, routed H.v against the single-GPU H.v, multi-rank frifull_mol.  Skipped on a one-GPU box (the
driver's GPU tier); run by hand with gpurun --gpus 2 / 8 during the round."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_multi_gpu_check_under_torchrun():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs, found {n}")
    world = 2 if n < 8 else 8
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "multi_gpu_check.py")],
                       cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("[multi]")]
    assert r.returncode == 0, r.stdout[-3000:]
    assert any("== single-rank oracle" in ln for ln in lines), lines
    assert any("distributed piv_comp_parallel" in ln for ln in lines), lines
    assert any("route p2p" in ln for ln in lines) and any("route nccl" in ln for ln in lines), lines
    assert any("routed H.v == single-GPU H.v" in ln for ln in lines), lines


def test_frisys_mol_driver_on_two_gpus(tmp_path):
    """frisys_mol's command line on 2 GPUs, started by host/bin/fries_launch -n 2 (the reference: mpirun -n 2): C++ only --
    the CUDA IPC handshake, the per-rank files dets<r>.dat / vals<r>.dat (vec_utils.hpp:713-750) and hash.dat
    (io_utils.cpp:589-605).  Energy within error bars of the single-GPU run and of the exact ground state; every stored
    determinant is on the rank that owns it; the REFERENCE's driver restarts from the checkpoint on 2 ranks."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oraclelib
    from driver_utils import OURS, REF, exact_ground_state, read_col, write_fcidump
    from fries_b200.synth import SynthMol
    from test_gpu_drivers import TINY, blocked_ratio
    sm = SynthMol(*TINY)
    om = oraclelib.OracleMol(sm)
    e_corr, e_hf, n = exact_ground_state(sm, om)
    fd = str(tmp_path / "FCIDUMP")
    write_fcidump(fd, sm, "D2")
    n_it = 6000
    res = {}
    for name, pre in (("two", [os.path.join(OURS, "fries_launch"), "-n", "2"]), ("one", [])):
        rd = str(tmp_path / name) + "/"
        os.makedirs(rd)
        cmd = pre + [os.path.join(OURS, "frisys_mol"), "--fcidump_path", fd, "--distribution", "HB_unnorm", "--vec_nonz", "150",
                     "--mat_nonz", "300", "--max_dets", "20000", "--epsilon", "0.05", "--target", "500", "--max_iter", str(n_it),
                     "--result_dir", rd, "--point_group", "D2"]
        r = subprocess.run(cmd, env=dict(os.environ, FRIES_SEED="3" if name == "two" else "5"), stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True, timeout=600)
        assert r.returncode == 0 and "Exception" not in r.stderr, (name, r.stderr[-800:])
        num, den = read_col(rd + "projnum.txt"), read_col(rd + "projden.txt")
        assert len(num) == n_it, (name, len(num))
        res[name] = blocked_ratio(num, den, burn=1000)
        assert len([ln for ln in r.stdout.splitlines() if ", en est: " in ln]) == n_it  # one writer: the owner of HF
    (e2, s2), (e1, s1) = res["two"], res["one"]
    print("exact", e_corr, "2 GPUs", res["two"], "1 GPU", res["one"])
    assert abs(e2 - e1) < 5 * (s1 + s2) + 2e-4
    assert abs(e2 - e_corr) < 5 * s2 + 2e-3 * abs(e_corr) + 2e-4
    # per-rank files; ownership by the saved scrambler
    rd = str(tmp_path / "two") + "/"
    scr = np.fromfile(rd + "hash.dat", dtype=np.uint32)
    assert scr.size == 2 * sm.n_orb
    nb = (2 * sm.n_orb + 7) // 8
    import fries_b200
    ctx = fries_b200.Context(0)
    total = 0
    for rank in range(2):
        raw = np.fromfile(rd + f"dets{rank}.dat", dtype=np.uint8).reshape(-1, nb)
        keys = np.array([int.from_bytes(bytes(row), "little") for row in raw], np.uint64)
        assert os.path.getsize(rd + f"vals{rank}.dat") == keys.size * 2 * 8
        _, own = fries_b200.hash_owner(ctx, keys, scr, 2)
        assert np.all(own == rank)
        total += keys.size
    assert total > 0
    ctx.close()
    # the reference restarts from this checkpoint on 2 ranks (its MPI collectives over oracle/mpi_shim)
    ref = os.path.join(REF, "frisys_mol")
    if os.path.exists(ref):
        rr = str(tmp_path / "ref_restart") + "/"
        os.makedirs(rr)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "mpi_shim", "shimrun.py"), "-n", "2", ref,
                            "--fcidump_path", fd, "--distribution", "HB_unnorm", "--vec_nonz", "150", "--mat_nonz", "300",
                            "--max_dets", "20000", "--epsilon", "0.05", "--target", "500", "--max_iter", "2000", "--result_dir", rr,
                            "--point_group", "D2", "--load_dir", rd], env=dict(os.environ, FRIES_SEED="9"),
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
        assert "Exception" not in r.stderr, r.stderr[-800:]
        num, den = read_col(rr + "projnum.txt"), read_col(rr + "projden.txt")
        assert len(num) == 2000
        er, sr = blocked_ratio(num, den, burn=200)
        print("reference restarted from the 2-GPU checkpoint:", er, sr)
        assert abs(er - e_corr) < 5 * sr + 2e-3 * abs(e_corr) + 5e-4


def test_frifull_mol_driver_on_two_gpus(tmp_path):
    """frifull_mol's command line on 2 GPUs (fries_launch -n 2; the reference: mpirun -n 2 frifull_mol): with a compression
    budget above the size of the space the iteration is a deterministic power iteration, so the 2-GPU run must reproduce the
    1-GPU run's files to their printed precision; the determinants of every rank's checkpoint belong to that rank."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oraclelib
    from driver_utils import OURS, exact_ground_state, read_col, write_hf_dir
    from fries_b200.synth import SynthMol
    from test_gpu_drivers import TINY
    sm = SynthMol(*TINY)
    om = oraclelib.OracleMol(sm)
    e_corr, e_hf, n = exact_ground_state(sm, om)
    d = str(tmp_path / "hf") + "/"
    write_hf_dir(d, sm, 0.05, float(e_hf))
    outs = {}
    for name, pre in (("two", [os.path.join(OURS, "fries_launch"), "-n", "2"]), ("one", [])):
        rd = str(tmp_path / name) + "/"
        os.makedirs(rd)
        cmd = pre + [os.path.join(OURS, "frifull_mol"), "--hf_path", d, "--vec_nonz", str(4 * n), "--max_dets", str(8 * n),
                     "--max_iter", "150", "--target", "110", "--result_dir", rd]
        r = subprocess.run(cmd, env=dict(os.environ, FRIES_SEED="3"), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                           timeout=600)
        assert r.returncode == 0 and "Exception" not in r.stderr, (name, r.stderr[-800:])
        outs[name] = {f: read_col(rd + f) for f in ("projnum.txt", "projden.txt", "S.txt", "norm.txt")}
        assert len([ln for ln in r.stdout.splitlines() if ", en est: " in ln]) == 150  # one writer: the owner of HF
    for f in ("projnum.txt", "projden.txt", "S.txt", "norm.txt"):
        a, b = outs["two"][f], outs["one"][f]
        assert a.shape == b.shape and a.size > 0, (f, a.shape, b.shape)
        assert np.allclose(a, b, rtol=2e-5, atol=1e-6), (f, a[:5], b[:5])
    en = outs["two"]["projnum.txt"][-1] / outs["two"]["projden.txt"][-1]
    assert e_corr - 1e-9 < en < 0
    rd = str(tmp_path / "two") + "/"
    scr = np.fromfile(rd + "hash.dat", dtype=np.uint32)
    nb = (2 * sm.n_orb + 7) // 8
    import fries_b200
    ctx = fries_b200.Context(0)
    total = 0
    for rank in range(2):
        raw = np.fromfile(rd + f"dets{rank}.dat", dtype=np.uint8).reshape(-1, nb)
        keys = np.array([int.from_bytes(bytes(row), "little") for row in raw], np.uint64)
        _, own = fries_b200.hash_owner(ctx, keys, scr, 2)
        assert np.all(own == rank)
        total += keys.size
    assert total == n  # the whole space is populated after 150 applications of H
    ctx.close()


def test_frisys_mol_det_space_on_two_gpus(tmp_path):
    """Semi-stochastic frisys_mol on 2 GPUs (--det_space; DistVec::init_dense is collective, vec_utils.hpp:858-897; the
    dense multiplication frisys_mol.cpp:479-485 routes every rank's exact connections to their owners): dense.txt holds one
    size per rank, each rank's share of the subspace comes first in its checkpoint, the energy agrees with the exact one."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oraclelib
    from driver_utils import OURS, exact_ground_state, read_col, write_fcidump, write_vec
    from fries_b200.synth import SynthMol
    from test_gpu_drivers import TINY, blocked_ratio
    sm = SynthMol(*TINY)
    om = oraclelib.OracleMol(sm)
    e_corr, e_hf, n = exact_ground_state(sm, om)
    fd = str(tmp_path / "FCIDUMP")
    write_fcidump(fd, sm, "D2")
    hf = np.array([sm.hf], np.uint64)
    k, v = om.h_apply(hf, np.ones(1), 0.0, 1.0)
    order = np.argsort(-np.abs(v), kind="stable")
    dense = [int(sm.hf)] + [int(x) for x in k[order] if int(x) != int(sm.hf)][:24]
    dp = str(tmp_path / "dense_dets.txt")
    open(dp, "w").write("\n".join(str(d) for d in dense) + "\n")
    ip = str(tmp_path / "ini_")
    others = [(int(a), float(b)) for a, b in zip(k, v) if int(a) != int(sm.hf)]
    write_vec(ip, [int(sm.hf)] + [a for a, _ in others], [100.0] + [-20.0 * b for _, b in others])
    rd = str(tmp_path / "two") + "/"
    os.makedirs(rd)
    n_it = 5000
    cmd = [os.path.join(OURS, "fries_launch"), "-n", "2", os.path.join(OURS, "frisys_mol"), "--fcidump_path", fd, "--distribution",
           "HB_unnorm", "--vec_nonz", "150", "--mat_nonz", "3000", "--max_dets", "20000", "--epsilon", "0.05", "--target", "500",
           "--max_iter", str(n_it), "--result_dir", rd, "--point_group", "D2", "--det_space", dp, "--ini_vec", ip]
    r = subprocess.run(cmd, env=dict(os.environ, FRIES_SEED="7"), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=600)
    assert r.returncode == 0 and "Exception" not in r.stderr, r.stderr[-800:]
    sizes = [int(x) for x in open(rd + "dense.txt").read().replace("\n", "").split(",") if x.strip()]
    assert len(sizes) == 2 and sum(sizes) == len(dense), sizes
    dense_h = [ln for ln in r.stdout.splitlines() if ln.startswith("Elements in dense H:")]
    assert dense_h and int(dense_h[0].split(":")[1]) > 0
    scr = np.fromfile(rd + "hash.dat", dtype=np.uint32)
    nb = (2 * sm.n_orb + 7) // 8
    import fries_b200
    ctx = fries_b200.Context(0)
    seen = []
    for rank in range(2):
        raw = np.fromfile(rd + f"dets{rank}.dat", dtype=np.uint8).reshape(-1, nb)
        first = np.array([int.from_bytes(bytes(row), "little") for row in raw[:sizes[rank]]], np.uint64)
        assert set(int(x) for x in first) <= set(dense)
        if first.size:
            _, own = fries_b200.hash_owner(ctx, first, scr, 2)
            assert np.all(own == rank)
        seen += [int(x) for x in first]
    ctx.close()
    assert sorted(seen) == sorted(dense)
    num, den = read_col(rd + "projnum.txt"), read_col(rd + "projden.txt")
    assert len(num) == n_it
    e, s = blocked_ratio(num, den, burn=1000)
    print("semi-stochastic on 2 GPUs: exact", e_corr, "ours", e, s)
    assert abs(e - e_corr) < 5 * s + 2e-3 * abs(e_corr) + 2e-4, (e, s, e_corr)

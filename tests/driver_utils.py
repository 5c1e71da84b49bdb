"""Input-file writers and output readers shared by the driver tests (FCIDUMP, legacy --hf_path directory)."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "fries_b200", "host", "bin")
REF = os.path.join(ROOT, "oracle", "_ref")

INV_LABEL = {"D2h": {0: 1, 7: 2, 6: 3, 1: 4, 5: 5, 2: 6, 3: 7, 4: 8}, "C2v": {0: 1, 2: 2, 3: 3, 1: 4}, "D2": {0: 1, 3: 2, 2: 3, 1: 4}}


def write_fcidump(path, sm, point_group):
    inv = INV_LABEL[point_group]
    T = sm.tot_orb
    e = sm.eris_chem
    with open(path, "w") as f:
        f.write(f"&FCI NORB={T},NELEC={sm.n_elec_total},MS2=0,\n")
        f.write("ORBSYM=" + ",".join(str(inv[int(s)]) for s in sm.symm_all) + ",\n")
        f.write("ISYM=1,\n&END\n")
        for i in range(T):
            for j in range(i + 1):
                for k in range(i + 1):
                    for l in range(k + 1):
                        if k * (k + 1) // 2 + l > i * (i + 1) // 2 + j:
                            continue
                        if e[i, j, k, l] != 0.0:
                            f.write(f"{float(e[i, j, k, l])!r} {i + 1} {j + 1} {k + 1} {l + 1}\n")
        for i in range(T):
            for j in range(i + 1):
                if sm.hcore[i, j] != 0.0:
                    f.write(f"{float(sm.hcore[i, j])!r} {i + 1} {j + 1} 0 0\n")
        f.write("0.0 0 0 0 0\n")


def write_hf_dir(d, sm, eps, hf_en):
    """legacy input directory of frifull_mol (io_utils.cpp:98-187); eris.txt = dense <ij|ab> = (ia|jb)"""
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "sys_params.txt"), "w") as f:
        f.write(f"n_elec\n{sm.n_elec_total}\nn_frozen\n{sm.n_frz}\nn_orb\n{sm.n_orb}\neps\n{eps!r}\nhf_energy\n{hf_en!r}\n")
    with open(os.path.join(d, "symm.txt"), "w") as f:
        f.write("\n".join(str(int(s)) for s in sm.symm_all) + "\n")
    np.savetxt(os.path.join(d, "hcore.txt"), sm.hcore, delimiter=",", fmt="%.17g")
    T = sm.tot_orb
    phys = np.ascontiguousarray(sm.eris_chem.transpose(0, 2, 1, 3))  # <ij|ab> = (ia|jb)
    np.savetxt(os.path.join(d, "eris.txt"), phys.reshape(T * T, T * T), delimiter=",", fmt="%.17g")


def write_vec(prefix, keys, vals):
    with open(prefix + "dets", "w") as f:
        f.write("\n".join(str(int(k)) for k in keys) + "\n")
    with open(prefix + "vals", "w") as f:
        f.write("\n".join(repr(float(v)) for v in vals) + "\n")


def run(exe, args, seed=1, timeout=900):
    env = dict(os.environ, FRIES_SEED=str(seed))
    r = subprocess.run([exe] + [str(a) for a in args], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=timeout)
    return r


def read_col(path):
    with open(path) as f:
        return np.array([float(x) for x in f.read().split()])


def exact_ground_state(sm, om):
    """E0 - E_HF of the symmetry sector of the HF determinant, by dense diagonalisation of the oracle's H"""
    hf = np.array([sm.hf], np.uint64)
    space = {int(sm.hf)}
    frontier = hf
    while True:  # close the space under H
        k, _ = om.h_apply(frontier, np.ones(len(frontier)), 0.0, 1.0)
        new = [int(x) for x in k if int(x) not in space]
        if not new:
            break
        space.update(new)
        frontier = np.array(new, np.uint64)
    dets = np.array(sorted(space), np.uint64)
    idx = {int(d): i for i, d in enumerate(dets)}
    n = len(dets)
    H = np.zeros((n, n))
    for j, d in enumerate(dets):
        k, v = om.h_apply(np.array([d], np.uint64), np.ones(1), 0.0, 1.0)
        for kk, vv in zip(k, v):
            H[idx[int(kk)], j] = vv
    assert np.allclose(H, H.T, atol=1e-12)
    e = np.linalg.eigvalsh(H)
    return e[0] - H[idx[int(sm.hf)], idx[int(sm.hf)]], H[idx[int(sm.hf)], idx[int(sm.hf)]], n

"""Energy parity at size (north_star: "projected energies must match the reference's within combined error bars").

Runs the product's C++ driver (fries_b200/host/bin/frisys_mol, GPU) and the reference's own driver (oracle/_ref/frisys_mol,
the reference's MPI code on the host cores through oracle/mpi_shim) on the SAME FCIDUMP, start vector and command line of a
Ne aug-cc-pVDZ-sized calculation (BASELINE.json configs[1]: NORB 22, NELEC 8, D2h, HB_unnorm, vec_nonz 242000, mat_nonz
260000, initiator 1, synthetic integrals), with different seeds, and compares the projected energies with the blocking /
autocorrelation analysis of fries_b200/stats.py (= the reference's Benchmarks/calc_stats.py).  Prints one JSON line.

  python tests/tools/energy_parity.py --iters 20000 --burn 4000          # ~6 min of reference on 16 cores
The start vector needs no GPU (the oracle's H.v, as bench.py's reference arm)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402  (configs, FCIDUMP writer, start vector)
from driver_utils import OURS, REF, read_col, write_vec  # noqa: E402
from fries_b200.stats import trajectory_stats  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="ne")
    ap.add_argument("--iters", type=int, default=20000)
    ap.add_argument("--burn", type=int, default=4000)
    ap.add_argument("--scale", type=float, default=1.0, help="shrink vec_nonz / mat_nonz / target (quick runs)")
    ap.add_argument("--ref_ranks", type=int, default=0)
    a = ap.parse_args()
    cfg = dict(bench.CONFIGS[a.config])
    for k in ("vec_nonz", "mat_nonz", "max_dets"):
        cfg[k] = int(cfg[k] * a.scale)
    cfg["target"] = cfg["target"] * a.scale
    import importlib.util
    spec = importlib.util.spec_from_file_location("fries_synth_standalone", os.path.join(ROOT, "fries_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(synth)
    sm = synth.SynthMol(cfg["system"], cfg["seed"], frozen=False)
    keys, vals = bench.reference_start_vector(cfg, sm)
    wd = tempfile.mkdtemp(prefix="fries_parity_")
    fd = os.path.join(wd, "FCIDUMP")
    bench.write_fcidump(fd, sm, cfg["point_group"])
    write_vec(os.path.join(wd, "ini_"), keys, vals)
    common = ["--fcidump_path", fd, "--distribution", cfg["dist"], "--vec_nonz", str(cfg["vec_nonz"]), "--mat_nonz",
              str(cfg["mat_nonz"]), "--max_dets", str(cfg["max_dets"]), "--epsilon", str(cfg["eps"]), "--target", str(cfg["target"]),
              "--initiator", str(cfg["initiator"]), "--max_iter", str(a.iters), "--ini_vec", os.path.join(wd, "ini_"),
              "--point_group", cfg["point_group"]]
    ranks = a.ref_ranks or bench.reference_rank_count()
    out = {"config": cfg["workload"] + (f" x{a.scale}" if a.scale != 1 else ""), "iterations": a.iters, "burn_in": a.burn}
    runs = [("ours", [os.path.join(OURS, "frisys_mol")], "11"),
            ("reference", [sys.executable, os.path.join(ROOT, "oracle", "mpi_shim", "shimrun.py"), "-n", str(ranks),
                           os.path.join(REF, "frisys_mol")], "12")]
    for name, pre, seed in runs:
        rd = os.path.join(wd, name) + "/"
        os.makedirs(rd)
        t0 = time.time()
        r = subprocess.run(pre + common + ["--result_dir", rd], env=dict(os.environ, FRIES_SEED=seed), stdout=subprocess.DEVNULL,
                           stderr=subprocess.PIPE, text=True)
        sec = time.time() - t0
        if r.returncode != 0 or "Exception" in r.stderr:
            out[name] = {"error": r.stderr[-400:]}
            continue
        st = trajectory_stats(read_col(rd + "projnum.txt"), read_col(rd + "projden.txt"), a.burn)
        out[name] = {"energy": st["energy"], "std_err": st["std_err"], "iat": st["iat"], "n": st["n"], "wall_s": round(sec, 1),
                     "ranks_or_gpus": ranks if name == "reference" else 1}
    if "energy" in out.get("ours", {}) and "energy" in out.get("reference", {}):
        d = out["ours"]["energy"] - out["reference"]["energy"]
        s = float(np.hypot(out["ours"]["std_err"], out["reference"]["std_err"]))
        out["delta"] = d
        out["combined_sigma"] = s
        out["within_2_sigma"] = bool(abs(d) < 2 * s)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()

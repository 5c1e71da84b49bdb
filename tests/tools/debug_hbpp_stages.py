"""Diagnostics: apply_HBPP_sys stage by stage, CUDA path against the oracle (chunk 1), on the input of
tests/test_gpu_fullsize.py::test_apply_hbpp_sys_properties_h2o_1e6 (or a smaller one: --n_det / --n_samp).
Without a GPU (--oracle_only) it prints the oracle's per-stage list lengths."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oraclelib  # noqa: E402
from fries_b200.synth import SynthMol  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n_det", type=int, default=50000)
ap.add_argument("--n_samp", type=int, default=1000000)
ap.add_argument("--giant", type=float, default=50.0)
ap.add_argument("--oracle_only", action="store_true")
a = ap.parse_args()

sm = SynthMol("h2o", 3, True)
rng = np.random.default_rng(5)
n_det, n_samp = a.n_det, a.n_samp
keys = np.concatenate([[sm.hf], sm.random_dets(n_det - 1, rng, 0)]).astype(np.uint64)
vals = rng.lognormal(0, 2, n_det) * rng.choice([-1.0, 1.0], n_det)
if a.giant > 0:
    vals[0] = a.giant * np.abs(vals).max()
u5 = rng.random(5)
cap = 2 * n_samp + n_det
om = oraclelib.OracleMol(sm)
L = oraclelib.lib()


def oracle_stage(s):
    ov = np.zeros(cap)
    od = np.zeros(cap, np.uint64)
    oo = np.zeros((cap, 4), np.uint8)
    os_ = np.zeros(cap, np.uint32)
    n = L.fo_debug_hbpp_stage(om.h, keys, vals, len(keys), 0.98, 1, u5, n_samp, cap, s, ov, od, oo.reshape(-1), os_)
    return ov[:n], od[:n], oo[:n], os_[:n]


if not a.oracle_only:
    import fries_b200
    from fries_b200 import _capi
    ctx = fries_b200.Context(0)
    mol = fries_b200.Mol.from_synth(ctx, sm)
    lib = _capi.lib if hasattr(_capi, "lib") else fries_b200.lib

with oraclelib.keep_chunk(1):
    for s in range(5):
        ov, od, oo, osub = oracle_stage(s)
        line = f"stage {s}: oracle n={len(ov)} sum={ov.sum():.10g} max={ov.max():.6g} n_at_unit={(ov == np.median(ov)).sum()}"
        if not a.oracle_only:
            gv = np.zeros(cap)
            gd = np.zeros(cap, np.uint32)
            gp = np.zeros(cap, np.uint32)
            gs = np.zeros(cap, np.uint32)
            n = C.c_size_t(0)
            rc = lib.fries_debug_hbpp_stage(mol.h, keys.ctypes.data, vals.ctypes.data, len(keys), 0.98, 1, u5.ctypes.data, n_samp,
                                            cap, s, gv.ctypes.data, gd.ctypes.data, gp.ctypes.data, gs.ctypes.data, C.byref(n))
            m = n.value
            line += f" | gpu rc={rc} n={m} sum={gv[:m].sum():.10g}"
            k = min(m, len(ov))
            same = (gd[:k] == od[:k]) & (gs[:k] == osub[:k])
            first = int(np.argmin(same)) if not same.all() else -1
            line += f" first_diff={first} n_same_prefix={same.sum()}"
            if first >= 0:
                lo = max(0, first - 2)
                line += f"\n   oracle det/sub/val {list(zip(od[lo:first+4].tolist(), osub[lo:first+4].tolist(), ov[lo:first+4].tolist()))}"
                line += f"\n   gpu    det/sub/val {list(zip(gd[lo:first+4].tolist(), gs[lo:first+4].tolist(), gv[lo:first+4].tolist()))}"
        print(line, flush=True)

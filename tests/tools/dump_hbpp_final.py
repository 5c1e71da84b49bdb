import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from hbpp_cases import CASES, make_case
import fries_b200
ctx = fries_b200.Context(0)
out = {}
for ci, case in enumerate(CASES):
    sm, keys, vals, new_hb, n_samp, cap, uni = make_case(case)
    gm = fries_b200.Mol.from_synth(ctx, sm)
    v, d, o = gm.apply_hbpp_sys(keys, vals, 0.97, new_hb, uni, n_samp, cap)
    out[f"c{ci}_v"] = v; out[f"c{ci}_d"] = d; out[f"c{ci}_o"] = o
np.savez("gpurun_out/hbpp_final.npz", **out)

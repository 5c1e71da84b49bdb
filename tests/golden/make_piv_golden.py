"""Generate tests/golden/piv_golden.npz (pivotal compression family) from the COMPILED REFERENCE
(oracle/_ref/libfries_ref.so, built from /root/reference by `make -C oracle ref`).  Run in the build container only.

    python tests/golden/make_piv_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import reflib  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from fries_b200.synth import SynthMol  # noqa: E402
from golden_cases import (PIV_ADJUST_CASES, PIV_BUDGET_CASES, PIV_COMP_CASES, PIV_HBPP_CASES, PIV_SAMP_CASES, hbpp_inputs,  # noqa: E402
                          piv_adjust_inputs, piv_budget_inputs, piv_comp_inputs, piv_samp_inputs)

out = {"mt5489_10000": reflib.mt19937(5489, 10000)[-4:]}
for i, case in enumerate(PIV_SAMP_CASES):
    v, keep, norm = piv_samp_inputs(case)
    rv, rk, used = reflib.piv_samp_serial(v, norm, case[2], keep, case[0])
    out[f"ps{i}_v"], out[f"ps{i}_k"], out[f"ps{i}_used"] = rv, rk, used
for i, case in enumerate(PIV_BUDGET_CASES):
    b, used = reflib.piv_budget(piv_budget_inputs(case), case[2], case[0])
    out[f"pb{i}_b"], out[f"pb{i}_used"] = b, used
for i, case in enumerate(PIV_ADJUST_CASES):
    v, keep, n_loc, tot_norm = piv_adjust_inputs(case)
    rv, rk, rn, rnorm = reflib.adjust_probs(v, n_loc, case[3], case[2], tot_norm, keep)
    out[f"pa{i}_v"], out[f"pa{i}_k"], out[f"pa{i}_n"], out[f"pa{i}_norm"] = rv, rk, rn, rnorm
for i, case in enumerate(PIV_COMP_CASES):
    rv, rk, used = reflib.piv_comp_parallel(piv_comp_inputs(case), case[2], case[0])
    nz = np.flatnonzero(rv)
    out[f"pc{i}_idx"], out[f"pc{i}_val"], out[f"pc{i}_used"] = nz.astype(np.uint32), rv[nz], used
    assert np.array_equal(rk == 1, rv == 0)
for i, case in enumerate(PIV_HBPP_CASES):
    sm = SynthMol(*case[0])
    keys, vals = hbpp_inputs(sm, case)
    rv, rd, ro, used = reflib.RefMol(sm).apply_hbpp_piv(keys, vals, 0.97, case[3], case[4], case[2], 4 * case[2] + 4 * case[1])
    out[f"ph{i}_v"], out[f"ph{i}_d"], out[f"ph{i}_o"], out[f"ph{i}_used"] = rv, rd, ro, used
np.savez_compressed(os.path.join(HERE, "piv_golden.npz"), **out)
print("wrote", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "piv_golden.npz")) // 1024, "kB")

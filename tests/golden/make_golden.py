"""Generate tests/golden/fries_golden.npz from the COMPILED REFERENCE (oracle/_ref/libfries_ref.so, built from
/root/reference by `make -C oracle ref`).  Run in the build container only; the fixture is committed so that the
oracle and the CUDA path can be pinned on machines where /root/reference does not exist.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import reflib  # noqa: E402
from fries_b200.synth import SynthMol  # noqa: E402
from golden_cases import (COMP_SUB_CASES, HBPP_CASES, MOL_CASES, VEC_COMP_CASES, comp_sub_inputs, hbpp_inputs,  # noqa: E402
                          mol_keys, vec_values)

out = {}
L = reflib.lib()
# a1: hashes with the scramblers of tests/test_vector.cpp:192-224 and a random 52-bit one
rng = np.random.default_rng(0)
for tag, scr, nb in [("tv1", np.arange(1, 9, dtype=np.uint32), 8), ("tv2", np.arange(8, 0, -1, dtype=np.uint32), 8),
                     ("rnd", rng.integers(0, 2**32, 52, dtype=np.uint64).astype(np.uint32), 52)]:
    keys = rng.integers(0, 2**nb, 300, dtype=np.uint64)
    h = np.zeros(300, np.uint64)
    o = np.zeros(300, np.int32)
    L.ref_hash_keys(keys, 300, nb, scr, 8, h, o)
    out[f"hash_{tag}_scr"], out[f"hash_{tag}_keys"], out[f"hash_{tag}_h"], out[f"hash_{tag}_o8"] = scr, keys, h, o
# a4/a5
for i, case in enumerate(VEC_COMP_CASES):
    v = vec_values(case)
    loc, glob, left, keep = reflib.find_preserve(v, case[1])
    out[f"fp{i}_loc"], out[f"fp{i}_glob"], out[f"fp{i}_left"], out[f"fp{i}_keep"] = loc, glob, left, keep
    for j, rn in enumerate((0.0, 0.37, 0.999999)):
        rv, rk, rnorm = reflib.sys_comp(v, loc, left, keep, rn)
        out[f"sc{i}_{j}_v"], out[f"sc{i}_{j}_k"], out[f"sc{i}_{j}_norm"] = rv, rk, rnorm
# a6
for i, case in enumerate(COMP_SUB_CASES):
    v, nd, sw, ss = comp_sub_inputs(case)
    for j, rn in enumerate((0.123, 0.9)):
        rv, ri = reflib.comp_sub(v, nd, sw, ss, case[2], rn, 4 * max(case[2], case[0]) + 64)
        out[f"cs{i}_{j}_v"], out[f"cs{i}_{j}_i"] = rv, ri
# molecule
for i, case in enumerate(MOL_CASES):
    sm = SynthMol(*case)
    rm = reflib.RefMol(sm)
    for k, t in rm.hb_tables().items():
        out[f"mol{i}_{k}"] = t
    keys = mol_keys(sm)
    out[f"mol{i}_diag"] = rm.diag(keys)
    se = [rm.sing_ex(k) for k in keys[:6]]
    de = [rm.doub_ex(k) for k in keys[:6]]
    out[f"mol{i}_sing_off"] = np.cumsum([0] + [len(x) for x in se])
    out[f"mol{i}_doub_off"] = np.cumsum([0] + [len(x) for x in de])
    out[f"mol{i}_sing_ex"], out[f"mol{i}_doub_ex"] = np.concatenate(se), np.concatenate(de)
    out[f"mol{i}_sing_el"] = np.concatenate([rm.sing_el(np.full(len(x), k, np.uint64), x) for k, x in zip(keys[:6], se)])
    out[f"mol{i}_doub_el"] = np.concatenate([rm.doub_el(x) for x in de])
    out[f"mol{i}_wt0"] = np.concatenate([[rm.hb_wt(0, k, o) for o in x[::9]] for k, x in zip(keys[:6], de)])
    out[f"mol{i}_wt1"] = np.concatenate([[rm.hb_wt(1, k, o) for o in x[::9]] for k, x in zip(keys[:6], de)])
    scr = np.random.default_rng(3).integers(0, 2**32, sm.n_bits, dtype=np.uint64).astype(np.uint32)
    hk, hv = rm.h_apply(keys[:5], np.linspace(-1, 1, 5) + 0.1, 1.0, -0.01, 200000, scr, scr)
    o = np.argsort(hk)
    out[f"mol{i}_hv_keys"], out[f"mol{i}_hv_vals"] = hk[o], hv[o]
for i, case in enumerate(HBPP_CASES):
    sm = SynthMol(*case[0])
    rm = reflib.RefMol(sm)
    keys, vals = hbpp_inputs(sm, case)
    uni, rv, rd, ro = rm.apply_hbpp_sys(keys, vals, 0.97, case[3], case[4], case[2], 4 * case[2] + 4 * case[1])
    out[f"hb{i}_uni"], out[f"hb{i}_v"], out[f"hb{i}_d"], out[f"hb{i}_o"] = uni, rv, rd, ro
np.savez_compressed(os.path.join(HERE, "fries_golden.npz"), **out)
print("wrote", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "fries_golden.npz")) // 1024, "kB")

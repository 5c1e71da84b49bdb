import sys, os, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from fries_b200.synth import SynthMol
from hbpp_cases import CASES, make_case

def gpu_stage(gm, keys, vals, p_doub, new_hb, uni, n_samp, cap, stage):
    from fries_b200._capi import lib, check, ptr
    v = np.zeros(cap); d = np.zeros(cap, np.uint32); p = np.zeros(cap, np.uint32); s = np.zeros(cap, np.uint32)
    n = C.c_size_t(0)
    check(lib.fries_debug_hbpp_stage(gm.h, ptr(keys), ptr(vals), keys.size, p_doub, new_hb, ptr(uni), n_samp, cap, stage,
                                     ptr(v), ptr(d), ptr(p), ptr(s), C.byref(n)))
    n = n.value
    return v[:n], d[:n], p[:n], s[:n]

if __name__ == "__main__":
    import fries_b200
    ctx = fries_b200.Context(0)
    out = {}
    for ci, case in enumerate(CASES):
        sm, keys, vals, new_hb, n_samp, cap, uni = make_case(case)
        gm = fries_b200.Mol.from_synth(ctx, sm)
        for st in range(5):
            v, d, p, s = gpu_stage(gm, keys, vals, 0.97, new_hb, uni, n_samp, cap, st)
            out[f"c{ci}_s{st}_v"] = v; out[f"c{ci}_s{st}_d"] = d; out[f"c{ci}_s{st}_p"] = p; out[f"c{ci}_s{st}_s"] = s
        gm.close()
    np.savez("gpurun_out/hbpp_dump.npz", **out)

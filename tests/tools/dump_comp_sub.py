import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import fries_b200
ctx = fries_b200.Context(0)
out = {}
for (n, n_sub, budget, jagged) in [(40000, 18, 9000, True), (40000, 2, 60000, False)]:
    rng = np.random.default_rng(n * 7 + n_sub)
    v = rng.lognormal(0, 2, n)
    v[rng.random(n) < 0.03] = 0
    nd = np.where(rng.random(n) < 0.4, rng.integers(1, 30, n), 0).astype(np.uint32)
    sw = rng.random((n, n_sub)) ** 3
    sw[rng.random((n, n_sub)) < 0.1] = 0
    ss = None
    if jagged:
        ss = rng.integers(1, n_sub + 1, n).astype(np.uint16)
        for i in range(n):
            sw[i, ss[i]:] = 0
    tot = sw.sum(1, keepdims=True)
    tot[tot == 0] = 1
    sw = sw / tot
    cap = 4 * max(budget, n) + 64
    gv, gi, left, loc = fries_b200.comp_sub(ctx, v, nd, sw, ss, budget, 0.123, cap)
    out[f"gv_{n_sub}"] = gv; out[f"gi_{n_sub}"] = gi; out[f"left_{n_sub}"] = left; out[f"loc_{n_sub}"] = loc
np.savez("gpurun_out/comp_sub_dump.npz", **out)

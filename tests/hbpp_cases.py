import numpy as np
from fries_b200.synth import SynthMol

CASES = [("ne", 2, False, 1, 50, 1), ("h2o", 3, True, 500, 2000, 1), ("h2o", 3, True, 500, 2000, 0)]

def make_values(rng, n, kind):
    v = np.concatenate([rng.lognormal(6, 1, max(1, n // 100)), rng.lognormal(-3, 2, n - max(1, n // 100))])
    rng.shuffle(v)
    v *= rng.choice([-1.0, 1.0], n)
    v[rng.random(n) < 0.05] = 0.0
    return v

def make_case(case):
    name, seed, frozen, n_det, n_samp, new_hb = case
    sm = SynthMol(name, seed, frozen)
    rng = np.random.default_rng(n_det + new_hb)
    keys = np.concatenate([[sm.hf], sm.random_dets(n_det - 1, rng, 0)]).astype(np.uint64) if n_det > 1 else \
        np.array([sm.hf], np.uint64)
    vals = make_values(rng, n_det, "fri")
    vals[0] = 100.0
    cap = 4 * n_samp + 4 * n_det
    mt = np.random.RandomState(1)  # mt19937(1): same uniforms as the reference shim with seed 1
    uni = mt.randint(0, 2**32, 5, dtype=np.uint64) / (1.0 + 0xFFFFFFFF)
    return sm, keys, vals, new_hb, n_samp, cap, uni

"""Seeded inputs shared by tests/golden/make_golden.py (reference outputs -> fixture) and the tests that replay them."""
import numpy as np

VEC_COMP_CASES = [(1, 1, "uniform"), (7, 3, "lognormal"), (1000, 100, "lognormal"), (1000, 2000, "uniform"),
                  (20000, 2500, "fri")]
COMP_SUB_CASES = [(5, 2, 4, False), (200, 8, 50, False), (3000, 11, 700, True), (20000, 18, 6000, True)]
MOL_CASES = [("ne", 2, False), ("n2", 7, True), ((8, 6, 0, [0, 0, 1, 2, 3, 0, 1, 2]), 4, True)]
# (mol case, n_det, n_samp, new_hb, mt19937 seed)
HBPP_CASES = [(("ne", 2, False), 1, 50, 1, 1), (("ne", 2, False), 300, 1000, 0, 1), (("n2", 7, True), 3000, 5000, 1, 2),
              (("h2o", 3, True), 3000, 5000, 0, 2)]


def make_values(rng, n, kind):
    if kind == "lognormal":
        v = rng.lognormal(0, 2.5, n)
    elif kind == "uniform":
        v = rng.random(n)
    else:
        v = np.concatenate([rng.lognormal(6, 1, max(1, n // 100)), rng.lognormal(-3, 2, n - max(1, n // 100))])
        rng.shuffle(v)
    v *= rng.choice([-1.0, 1.0], n)
    v[rng.random(n) < 0.05] = 0.0
    return v


def vec_values(case):
    n, budget, kind = case
    return make_values(np.random.default_rng(n + budget), n, kind)


def comp_sub_inputs(case):
    n, n_sub, budget, jagged = case
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
    sw[:, 0] += (sw.sum(1) == 0)
    sw = sw / sw.sum(1, keepdims=True)
    return v, nd, sw, ss


def mol_keys(sm):
    rng = np.random.default_rng(5)
    return np.concatenate([[sm.hf], sm.random_dets(24, rng, None)]).astype(np.uint64)


def hbpp_inputs(sm, case):
    _, n_det, n_samp, new_hb, seed = case
    rng = np.random.default_rng(n_det + new_hb)
    keys = np.concatenate([[sm.hf], sm.random_dets(n_det - 1, rng, 0)]).astype(np.uint64) if n_det > 1 else \
        np.array([sm.hf], np.uint64)
    vals = make_values(rng, n_det, "fri")
    vals[0] = 100.0
    return keys, vals

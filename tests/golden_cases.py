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


# ---- pivotal family (compress_utils.cpp:354-681): fixture tests/golden/piv_golden.npz ------------------------------
# (seed, n, n_samp, fraction preserved, fraction of zeros); the draws are the first outputs of std::mt19937(seed)
PIV_SAMP_CASES = [(1, 200, 20, 0.0, 0.0), (2, 5000, 700, 0.1, 0.05), (3, 40000, 9000, 0.3, 0.0), (5, 3000, 1, 0.2, 0.0),
                  (6, 3000, 0, 0.2, 0.1)]
PIV_BUDGET_CASES = [(1, 1, 100), (2, 2, 7), (3, 4, 1000), (4, 8, 33), (5, 8, 100000), (6, 3, 2)]
# (seed, n, n_samp_tot, exp_nsamp_loc, budget rounded up?)
PIV_ADJUST_CASES = [(1, 500, 1000, 37.4, True), (2, 500, 1000, 37.4, False), (3, 3000, 9000, 411.9, True),
                    (4, 3000, 9000, 411.05, False)]
PIV_COMP_CASES = [(1, 300, 40), (2, 20000, 3000), (3, 20000, 19990)]


def piv_samp_inputs(case):
    """values whose non-preserved magnitudes are all below seg_norm / n_samp (piv_samp_serial's contract)"""
    seed, n, n_samp, frac_keep, zeros = case
    n_samp = max(n_samp, 1)
    rng = np.random.default_rng(seed)
    v = rng.random(n) ** 3 * np.where(rng.random(n) < 0.5, -1.0, 1.0)
    v[rng.random(n) < zeros] = 0
    keep = (rng.random(n) < frac_keep).astype(np.uint8)
    keep[v == 0] = 0
    for _ in range(100):
        norm = np.abs(v[keep == 0]).sum()
        big = (keep == 0) & (np.abs(v) >= norm / n_samp)
        if not big.any():
            break
        v[big] *= 0.5
    return v, keep, float(np.abs(v[keep == 0]).sum())


def piv_budget_inputs(case):
    seed, n_procs, n_samp = case
    return np.random.default_rng(100 + seed).random(n_procs) * 1000


def piv_adjust_inputs(case, hot_factor=0.9999):
    """one rank's residual: exp_loc sampling units of tot_norm / n_tot, five elements just below one unit"""
    seed, n, n_tot, exp_loc, up = case
    rng = np.random.default_rng(200 + seed)
    v = rng.random(n) * np.where(rng.random(n) < 0.5, -1.0, 1.0)
    keep = (rng.random(n) < 0.1).astype(np.uint8)
    tot_norm = 5000.0
    unit = tot_norm / n_tot
    free = np.flatnonzero(keep == 0)
    hot = rng.choice(free, 5, replace=False)
    rest = np.setdiff1d(free, hot)
    v[hot] = np.sign(v[hot]) * hot_factor * unit
    v[rest] *= (exp_loc - 5 * hot_factor) * unit / np.abs(v[rest]).sum()
    n_loc = int(np.ceil(exp_loc)) if up else int(exp_loc)
    return v, keep, n_loc, tot_norm


def piv_comp_inputs(case):
    seed, n, m = case
    rng = np.random.default_rng(300 + seed)
    return rng.standard_normal(n) * np.exp(3 * rng.standard_normal(n))


# apply_HBPP_piv: (mol case, n_det, n_samp, new_hb, mt19937 seed), inputs as hbpp_inputs
PIV_HBPP_CASES = [(("ne", 2, False), 1, 50, 1, 1), (("ne", 2, True), 300, 1000, 1, 2), (("h2o", 3, True), 1000, 1500, 0, 2)]

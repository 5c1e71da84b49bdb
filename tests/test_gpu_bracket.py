"""The bracketed threshold solve (csrc/compress.cuh: bracket_solve) against the plain rounds.

The plain rounds restate the reference's iteration (find_preserve compress_utils.cpp:52-92, find_keep_sub :153-265)
and are the ones pinned against the oracle in test_gpu_parity.py.  The bracketed solve starts from the fixed point of
a previous run of the same compression, so a one-shot call never uses it: `fries_debug_set_repeat(n)` runs a
standalone compression n times on the same inputs and returns the last run.  Both must give the same preserved set,
the same budget left and the same samples (up to the counted FP-boundary ties of the resampling line)."""
import ctypes as C

import numpy as np
import pytest

import oraclelib
from test_gpu_parity import comp_sub_inputs, hbpp_inputs, make_values

pytestmark = pytest.mark.gpu
REL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    import fries_b200
    c = fries_b200.Context(0)
    yield c
    c.close()


class repeat:
    """run standalone compressions n times; the warm-up runs see values perturbed by ~perturb (relative)"""

    def __init__(self, n, perturb=0.0):
        self.n, self.perturb = n, perturb

    def __enter__(self):
        from fries_b200._capi import check, lib
        check(lib.fries_debug_set_repeat(self.n))
        check(lib.fries_debug_set_perturb(self.perturb))

    def __exit__(self, *a):
        from fries_b200._capi import check, lib
        check(lib.fries_debug_set_repeat(1))
        check(lib.fries_debug_set_perturb(0.0))


def last_fast():
    from fries_b200._capi import check, lib
    n = C.c_int(0)
    check(lib.fries_debug_last_fast(C.byref(n)))
    return n.value


@pytest.mark.parametrize("n,budget,kind", [(1000, 100, "lognormal"), (50000, 5000, "fri"), (300000, 30000, "fri"),
                                            (300000, 250000, "lognormal"), (2000000, 700000, "lognormal"),
                                            # dense near the threshold: tens of thousands of candidates, i.e. the
                                            # grid-distributed candidate rounds
                                            (2000000, 500000, "uniform"), (3000000, 2500000, "uniform")])
@pytest.mark.parametrize("perturb", [0.0, 0.003, -0.004])
def test_find_preserve_bracket(ctx, n, budget, kind, perturb):
    import fries_b200
    rng = np.random.default_rng(n + budget)
    v = make_values(rng, n, kind)
    p_loc, p_glob, p_left, p_keep = fries_b200.find_preserve(ctx, v, budget)
    assert last_fast() == 0
    with repeat(3, perturb):
        b_loc, b_glob, b_left, b_keep = fries_b200.find_preserve(ctx, v, budget)
        assert last_fast() == 1, "the third run did not use the bracketed solve"
    assert np.array_equal(b_keep, p_keep), f"kept sets differ in {np.sum(b_keep != p_keep)} places"
    assert b_left == p_left and b_glob == p_glob
    assert b_loc == pytest.approx(p_loc, rel=REL, abs=1e-300)
    o_loc, o_glob, o_left, o_keep = oraclelib.find_preserve(v, budget)
    assert np.array_equal(b_keep, o_keep) and b_left == o_left


@pytest.mark.parametrize("n,n_sub,budget,jagged", [(3000, 11, 700, True), (40000, 18, 9000, True),
                                                    (40000, 2, 60000, False), (200000, 21, 150000, True)])
@pytest.mark.parametrize("perturb", [0.0, 0.003, -0.004])
def test_comp_sub_bracket(ctx, n, n_sub, budget, jagged, perturb):
    import fries_b200
    rng = np.random.default_rng(n * 7 + n_sub)
    v, nd, sw, ss = comp_sub_inputs(rng, n, n_sub, jagged)
    cap = 4 * max(budget, n) + 64
    n_fast = 0
    for rn in (0.123, 0.9):
        pv, pi, p_left, p_loc = fries_b200.comp_sub(ctx, v, nd, sw, ss, budget, rn, cap)
        with repeat(3, perturb):
            bv, bi, b_left, b_loc = fries_b200.comp_sub(ctx, v, nd, sw, ss, budget, rn, cap)
            n_fast += last_fast()
        assert b_left == p_left and b_loc == pytest.approx(p_loc, rel=REL)
        pset = {(int(a), int(b)): x for (a, b), x in zip(pi, pv)}
        bset = {(int(a), int(b)): x for (a, b), x in zip(bi, bv)}
        ties = set(pset) ^ set(bset)
        assert len(ties) <= max(1, n // 50000), f"{len(ties)} index mismatches of {len(pset)}"
        for k in set(pset) & set(bset):
            assert bset[k] == pytest.approx(pset[k], rel=1e-11)
        if not ties:
            assert np.array_equal(bi, pi)
        with oraclelib.keep_chunk(1):
            ov, oi, oleft, oloc = oraclelib.comp_sub(v, nd, sw, ss, budget, rn, cap)
        oset = {(int(a), int(b)) for a, b in oi}
        assert b_left == oleft and len(oset ^ set(bset)) <= max(1, n // 50000)
    assert n_fast == 2, "the bracketed solve was not used"


@pytest.fixture(scope="module", params=[("ne", 2, False), ("n2", 7, True)])
def mols(request, ctx):
    import fries_b200
    from fries_b200.synth import SynthMol
    system, seed, frozen = request.param
    sm = SynthMol(system, seed, frozen=frozen)
    gm = fries_b200.Mol.from_synth(ctx, sm)
    yield sm, gm
    gm.close()


@pytest.mark.parametrize("new_hb", [0, 1])
@pytest.mark.parametrize("n_det,n_samp", [(500, 2000), (20000, 30000)])
@pytest.mark.parametrize("perturb", [0.0, 0.003, -0.004])
def test_apply_hbpp_sys_bracket(mols, new_hb, n_det, n_samp, perturb):
    sm, gm = mols
    keys, vals = hbpp_inputs(sm, n_det, new_hb)
    uni = np.random.default_rng(n_det).random(5)
    cap = 4 * n_samp + 4 * n_det
    pv, pd, po = gm.apply_hbpp_sys(keys, vals, 0.97, new_hb, uni, n_samp, cap)
    with repeat(3, perturb):
        bv, bd, bo = gm.apply_hbpp_sys(keys, vals, 0.97, new_hb, uni, n_samp, cap)
        n_fast = last_fast()
    if n_det >= 20000:  # small cases: a per-element perturbation of 0.3 % moves the fixed point out of the bracket
        assert n_fast >= 4, f"only {n_fast} of 5 stages used the bracketed solve"
    pset = {(int(d), tuple(o)): v for d, o, v in zip(pd, po.tolist(), pv)}
    bset = {(int(d), tuple(o)): v for d, o, v in zip(bd, bo.tolist(), bv)}
    ties = set(pset) ^ set(bset)
    assert len(ties) <= max(0 if n_det < 1000 else 6, len(pset) // 5000), f"{len(ties)} of {len(pset)} differ"
    for k in set(pset) & set(bset):
        assert bset[k] == pytest.approx(pset[k], rel=1e-9)
    if not ties:
        assert np.array_equal(bd, pd) and np.array_equal(bo, po)


def test_iterate_bracket_on_off(ctx):
    """frisys_mol iterations with and without the bracketed solve, same uniforms.  New determinants are appended in
    the order the merge kernel's atomics resolve, so two runs of the SAME configuration already differ by a few
    resampled elements per iteration (a random walk of about one sample unit = 5e-5 of the one-norm per iteration); the
    comparison is therefore statistical (one-norm to 1e-3 after 8 iterations), and the point of the test is that the
    bracketed solve engages and leaves budgets and sizes intact."""
    import fries_b200
    from fries_b200._capi import FrisysParams, check, lib
    from fries_b200.synth import SynthMol
    sm = SynthMol("ne", 2, frozen=False)
    mol = fries_b200.Mol.from_synth(ctx, sm)
    rng = np.random.default_rng(11)
    scr = rng.integers(0, 2**32, sm.n_bits, dtype=np.uint64).astype(np.uint32)
    hf = np.array([sm.hf], np.uint64)
    hf_en = float(mol.diag(hf)[0])
    keys = np.concatenate([hf, sm.random_dets(30000, rng, 0)])
    keys = np.unique(keys)
    vals = rng.lognormal(0, 2, keys.size) * rng.choice([-1.0, 1.0], keys.size)
    vals *= 20000 / np.abs(vals).sum()
    uni = rng.random((8, 6))
    p = FrisysParams(eps=0.001, init_thresh=1.0, p_doub=0.97, new_hb=1, matr_samp=25000, target_nonz=20000, en_shift=0.0)
    out = {}
    for on in (0, 1):
        check(lib.fries_debug_set_bracket(on))
        try:
            vec = fries_b200.Vec(ctx, 200000, sm.n_bits, sm.n_elec, 2, scr, scr)
            vec.set_diag_mol(mol, hf_en)
            vec.upload(keys, np.stack([vals, np.zeros_like(vals)]))
            vec.frisys_setup(mol, 100000, hf, np.ones(1), hf, np.array([0.0]))
            rows = []
            for it in range(8):
                st = vec.frisys_iterate(p, uni[it])
                states = vec.states()
                rows.append((st.curr_size, st.glob_norm, st.numer, st.denom, states[:7, 10].copy(), states[:7, 4].copy()))
            out[on] = rows
            vec.close()
        finally:
            check(lib.fries_debug_set_bracket(1))
    mol.close()
    for it in range(8):
        a, b = out[0][it], out[1][it]
        print(it, a[0], b[0], a[1], b[1], a[2], b[2], b[4], b[5], a[5])
    for it in range(8):
        a, b = out[0][it], out[1][it]
        assert not a[4].any()
        assert abs(a[0] - b[0]) <= 3 and b[1] == pytest.approx(a[1], rel=1e-3)
        assert b[2] == pytest.approx(a[2], rel=1e-2, abs=1e-6) and b[3] == pytest.approx(a[3], rel=2e-3)
    fast = np.array([r[4] for r in out[1]])
    print("fast flags per iteration (stages 0-4, finalize, find_preserve):\n", fast)
    print("rounds:\n", np.array([r[5] for r in out[1]]))
    assert fast[3:, :5].mean() > 0.8, "the bracketed solve rarely engaged in the HB-PP stages"

"""GPU tier: stochastic outputs are unbiased (the design of the reference's tests/clt harness, re-implemented):
averaging many compressions of one fixed input, driven by independent uniforms, converges to the input at the CLT
rate; and the HB-PP compressed Hamiltonian column sums reproduce H.v in expectation."""
import numpy as np
import pytest

import oraclelib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import fries_b200
    c = fries_b200.Context(0)
    yield c
    c.close()


def test_vector_compression_unbiased(ctx):
    import fries_b200
    rng = np.random.default_rng(0)
    n, budget, reps = 4000, 600, 1500
    v = rng.lognormal(0, 1.5, n) * rng.choice([-1.0, 1.0], n)
    loc, glob, left, keep = fries_b200.find_preserve(ctx, v, budget)
    acc = np.zeros(n)
    for r in range(reps):
        out, _, _ = fries_b200.sys_comp(ctx, v, [loc], left, keep, rng.random())
        assert np.count_nonzero(out) <= budget
        acc += out
    mean = acc / reps
    kept = keep.astype(bool)
    assert np.allclose(mean[kept], v[kept], rtol=1e-12, atol=0)  # preserved exactly (up to the averaging round-off)
    # a resampled element is +-unit with probability |v| / unit: compare with the exact binomial standard error
    unit = loc / left
    prob = np.abs(v[~kept]) / unit
    assert prob.max() < 1.0
    se = unit * np.sqrt(prob * (1 - prob) / reps)
    z = (mean[~kept] - v[~kept]) / se
    well_sampled = prob * reps >= 20
    assert np.abs(z[well_sampled]).max() < 5.5
    assert abs(z[well_sampled].mean()) < 0.2 and 0.8 < z[well_sampled].std() < 1.2
    assert np.abs(mean - v).sum() / np.abs(v).sum() < 0.05
    # one-norm is conserved exactly by systematic resampling
    assert np.abs(out).sum() == pytest.approx(np.abs(v).sum(), rel=1e-9)


def test_hbpp_compression_unbiased(ctx):
    """E[ sum_samples value * e_(target) ] = H_offdiag . v for both factorizations (heat_bathPP.cpp:686-992)"""
    import fries_b200
    from fries_b200.synth import SynthMol
    sm = SynthMol((8, 6, 0, [0, 0, 1, 2, 3, 0, 1, 2]), 4)
    gm = fries_b200.Mol.from_synth(ctx, sm)
    om = oraclelib.OracleMol(sm)
    rng = np.random.default_rng(1)
    keys = np.concatenate([[sm.hf], sm.random_dets(11, rng, 0)]).astype(np.uint64)
    vals = rng.lognormal(0, 1, keys.size)  # apply_HBPP_sys compresses |v|; the spawn loop restores the sign
    # exact off-diagonal H.v
    ek, ev = om.h_apply(keys, vals, 0.0, 1.0)
    diag = om.diag(keys)
    exact = dict(zip(ek.tolist(), ev.tolist()))
    for k, v, d in zip(keys.tolist(), vals.tolist(), diag.tolist()):
        exact[k] -= d * v
    n_sing = int(gm.sing_ex(keys[:1])[0][-1])
    n_doub = int(gm.doub_ex(keys[:1])[0][-1])
    p_doub = n_doub / (n_sing + n_doub)
    reps, n_samp = 600, 400
    for new_hb in (1, 0):
        acc = {}
        sq = {}
        for r in range(reps):
            sv, sd, so = gm.apply_hbpp_sys(keys, vals, p_doub, new_hb, rng.random(5), n_samp, 8 * n_samp)
            contrib = {}
            for val, d, o in zip(sv, sd, so):
                k = int(keys[d])
                if o[2] == 0 and o[3] == 0:
                    nk = (k & ~(1 << int(o[0]))) | (1 << int(o[1]))
                else:
                    nk = (k & ~((1 << int(o[0])) | (1 << int(o[1])))) | (1 << int(o[2])) | (1 << int(o[3]))
                contrib[nk] = contrib.get(nk, 0.0) + val
            for k, x in contrib.items():
                acc[k] = acc.get(k, 0.0) + x
                sq[k] = sq.get(k, 0.0) + x * x
        tot = sum(abs(x) for x in exact.values())
        err = sum(abs(acc.get(k, 0.0) / reps - x) for k, x in exact.items())
        # relative L1 error of the averaged estimate: CLT-small, and no systematic offset on the large elements
        assert err / tot < 0.12, (new_hb, err / tot)
        big = sorted(exact, key=lambda k: -abs(exact[k]))[:20]
        for k in big:
            m = acc.get(k, 0.0) / reps
            s = np.sqrt(max(sq.get(k, 0.0) / reps - m * m, 0) / reps) + 1e-9
            assert abs(m - exact[k]) < 6 * s + 1e-3 * abs(exact[k]), (new_hb, k, m, exact[k], s)
    gm.close()

"""GPU tier: the fused vector kernel of the frisys_mol iteration (csrc/vecphase.cu) against the oracle.

The kernel runs find_preserve (compress_utils.cpp:29-105), sys_comp (:278-327) and the deletion of the zeroed elements
(frisys_mol.cpp:533-539, vec_utils.hpp:458-476) in one launch.  Its compression half is exposed by fries_debug_vec_phase and
must reproduce, on the stored vector in storage order:
  * find_preserve's preserved set and budget, bit for bit (a threshold test, no prefix sums involved);
  * sys_comp's resampled set up to counted FP-boundary ties (the prefix sums are associated per thread of four elements
    instead of running sequentially), magnitude residual / budget with the old sign;
  * the survivors in their old order (stable compaction), exact zeros kept as the reference keeps them.
Both solve paths: the first call of a scratch has no threshold bracket (plain rounds), the second one has the fixed point
of the first (bracketed solve on CTA 0)."""
import numpy as np
import pytest

import oraclelib
from fries_b200.synth import SynthMol

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.fixture(scope="module")
def ctx():
    import fries_b200
    c = fries_b200.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("n,budget,seed", [(5000, 1500, 0), (200_000, 60_000, 1), (1_000_000, 400_000, 2), (3000, 5000, 3)])
def test_vec_phase_matches_find_preserve_and_sys_comp(ctx, n, budget, seed):
    import fries_b200
    sm = SynthMol("ne", 2, True)
    mol = fries_b200.Mol.from_synth(ctx, sm)
    rng = np.random.default_rng(seed)
    keys = np.concatenate([[sm.hf], sm.random_dets(n - 1, rng, 0)]).astype(np.uint64)
    keys = np.unique(keys)  # distinct determinants; storage order = this order
    n = keys.size
    vals = rng.lognormal(0, 2, n) * rng.choice([-1.0, 1.0], n)
    vals[rng.integers(0, n, max(1, n // 500))] = 0.0  # exact zeros: neither sampled nor deleted (sys_comp :311)
    scr = rng.integers(0, 2**32, sm.n_bits, dtype=np.uint64).astype(np.uint32)
    vec = fries_b200.Vec(ctx, 2 * n + 64, sm.n_bits, sm.n_elec, 2, scr, scr)
    hf = np.array([sm.hf], np.uint64)
    vec.set_diag_mol(mol, 0.0)
    vec.frisys_setup(mol, 1024, hf, np.ones(1), hf, np.ones(1))
    for attempt, path in enumerate(["plain rounds", "bracketed solve"]):
        for rn in (0.37,):
            # upload keeps the zeros out of the store (fries_vec_upload drops all-zero elements), so compare on the stored set
            vec.upload(keys, np.stack([vals, np.zeros(n)]))
            sk, sv = vec.download()
            v0 = sv[0].copy()
            loc, glob, left, kept = vec.debug_vec_phase(budget, rn)
            o_loc, o_glob, o_left, o_keep = oraclelib.find_preserve(v0, budget)
            assert left == o_left and kept == int(o_keep.sum()), path
            assert glob == pytest.approx(o_glob, rel=1e-13) and loc == pytest.approx(o_loc, rel=1e-12)
            o_out, o_del = oraclelib.sys_comp(v0, [o_loc], o_left, o_keep.copy(), rn)[:2]
            gk, gv = vec.download()
            exp_keep = ~o_del.astype(bool)
            gset = dict(zip(gk.tolist(), gv[0].tolist()))
            oset = dict(zip(sk[exp_keep].tolist(), o_out[exp_keep].tolist()))
            ties = set(gset) ^ set(oset)
            assert len(ties) <= max(2, len(oset) // 50000), f"{path}: {len(ties)} of {len(oset)} survivors differ (FP-boundary ties)"
            both = [k for k in gset if k in oset]
            assert np.allclose([gset[k] for k in both], [oset[k] for k in both], rtol=1e-11, atol=0), path
            # preserved elements are untouched, bit for bit
            kd = dict(zip(sk[o_keep.astype(bool)].tolist(), v0[o_keep.astype(bool)].tolist()))
            assert all(gset[k] == x for k, x in kd.items()), path
            # stable compaction: the survivors keep their relative order; row 1 is zero; the index finds every survivor
            pos = {k: i for i, k in enumerate(sk.tolist())}
            order = [pos[k] for k in gk.tolist()]
            assert order == sorted(order), path
            assert not gv[1].any()
            assert vec.dot(gk, np.ones(gk.size)) == pytest.approx(gv[0].sum(), rel=1e-12, abs=1e-9)
            if not ties:
                assert np.array_equal(gk, sk[exp_keep])
            print(f"vec_phase n={n} budget={budget} {path}: kept {kept}, left {left}, {len(ties)} ties")
    vec.close()
    mol.close()

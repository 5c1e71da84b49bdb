"""The reference's own known-answer test of the un-normalised HB-PP factorisation (tests/test_hamiltonian.cpp:454-520,
"Complete test of compression of un-normalized HB-PP factorization"): with a budget larger than the number of
excitations nothing is resampled, apply_HBPP_sys returns every excitation of the Hartree-Fock determinant exactly once
and each value is (weight / total sampling probability) x matrix element -- i.e. +-1 with the reference's unit matrix
element shortcuts, the bare matrix element here.  CPU tier: the oracle; GPU tier: the CUDA path through the C-ABI."""
import numpy as np
import pytest

import oraclelib
from fries_b200.synth import SynthMol

CASES = [("ne", 2, True), ("h2o", 3, True), ((8, 6, 0, [0, 0, 1, 2, 3, 0, 1, 2]), 4, True)]


def column_of(sing, doub, sing_el, doub_el):
    exp = {}
    for o, e in zip(sing.tolist(), sing_el):
        exp[(o[0], o[1], 0, 0)] = e
    for o, e in zip(doub.tolist(), doub_el):
        exp[tuple(o)] = e
    return exp


def check_exact_limit(sm, out_val, out_det, out_orbs, exp, value):
    got = {}
    for v, d, o in zip(out_val, out_det, out_orbs.tolist()):
        assert d == 0
        k = tuple(o)
        assert k not in got, f"excitation {k} returned twice"
        got[k] = v
    dropped = {k for k, e in exp.items() if abs(e * value) < 1e-9}  # the 1e-9 cutoff of heat_bathPP.cpp:980
    assert set(got) | dropped == set(exp), (len(got), len(exp), list(set(exp) ^ set(got))[:5])
    for k, v in got.items():
        # the reference's margin: |v / matrix element| == 1 to 1e-7
        assert abs(v) == pytest.approx(abs(exp[k] * value), rel=1e-7), k


@pytest.mark.parametrize("case", CASES)
def test_oracle_hbpp_exact_limit(case):
    sm = SynthMol(*case)
    om = oraclelib.OracleMol(sm)
    hf = int(sm.hf)
    se, de = om.sing_ex(hf), om.doub_ex(hf)
    exp = column_of(se, de, om.sing_el([hf] * len(se), se), om.doub_el(de))
    n_ex = len(exp)
    p_doub = len(de) / n_ex
    for value in (1.0, -3.5):
        ov, od, oo = om.apply_hbpp_sys(np.array([hf], np.uint64), np.array([value]), p_doub, 1, np.full(5, 0.37),
                                       8 * n_ex, 16 * n_ex)
        check_exact_limit(sm, ov, od, oo, exp, value)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_hbpp_exact_limit(case):
    import fries_b200
    sm = SynthMol(*case)
    ctx = fries_b200.Context(0)
    gm = fries_b200.Mol.from_synth(ctx, sm)
    hf = np.array([sm.hf], np.uint64)
    se, de = gm.sing_ex(hf)[1], gm.doub_ex(hf)[1]
    exp = column_of(se, de, gm.sing_el(np.repeat(hf, len(se)), se), gm.doub_el(de))
    n_ex = len(exp)
    p_doub = len(de) / n_ex
    for value in (1.0, -3.5):
        gv, gd, go = gm.apply_hbpp_sys(hf, np.array([value]), p_doub, 1, np.full(5, 0.37), 8 * n_ex, 16 * n_ex)
        check_exact_limit(sm, gv, gd, go, exp, value)
    gm.close()
    ctx.close()

"""CPU tier: the spawn loop body of one sample (frisys_mol.cpp:436-461; `hbpp_spawn_element` of csrc/hbpp_prov.cuh, which the
finalize kernel calls in spawn mode), compiled for the host: new determinant by the reference's own sing_det / doub_det
(compiled reference where built, else bit arithmetic), element -eps x value x sign(parent), initiator flag from the parent's
magnitude against init_thresh."""
import ctypes as C

import numpy as np
import pytest

import reflib
from hostcheck import hc

INI = 1 << 63


@pytest.mark.parametrize("n_orb", [10, 22, 26, 31])
def test_spawn_element(n_orb):
    H = hc.lib()
    R = reflib.lib() if reflib.available() else None
    rng = np.random.default_rng(n_orb)
    for _ in range(500):
        k = int(rng.integers(0, 2**(2 * n_orb), dtype=np.uint64))
        occ = [i for i in range(2 * n_orb) if k >> i & 1]
        virt = [i for i in range(2 * n_orb) if not k >> i & 1]
        if len(occ) < 2 or len(virt) < 2:
            continue
        is_doub = bool(rng.random() < 0.7)
        if is_doub:
            o = sorted(int(x) for x in rng.choice(occ, 2, replace=False)) + sorted(int(x) for x in rng.choice(virt, 2, replace=False))
        else:
            o = [int(rng.choice(occ)), int(rng.choice(virt)), 0, 0]
            if o[1] == 0:
                continue
        orbs = np.array(o, np.uint8)
        el, cv = float(rng.normal()), float(rng.normal())
        eps, thresh = float(rng.random() * 0.01), float(rng.choice([0.0, 0.5, 1.0, 10.0]))
        if rng.random() < 0.1:
            cv = thresh  # exactly at the threshold: an initiator (>=)
        add = C.c_double(0)
        nk = H.hc_spawn_element(k, orbs, int(is_doub), el, cv, eps, thresh, C.byref(add))
        if R is not None:  # the reference's doub_det / sing_det through its parity twins
            kk = C.c_uint64(k)
            (R.ref_doub_det_parity if is_doub else R.ref_sing_det_parity)(C.byref(kk), orbs[:4] if is_doub else orbs[:2])
            new_det = kk.value
        elif is_doub:
            new_det = (k & ~((1 << o[0]) | (1 << o[1]))) | (1 << o[2]) | (1 << o[3])
        else:
            new_det = (k & ~(1 << o[0])) | (1 << o[1])
        want_add = -eps * el
        if cv < 0:
            want_add *= -1
        assert nk & ~INI == new_det
        assert bool(nk & INI) == (abs(cv) >= thresh)
        assert add.value == want_add

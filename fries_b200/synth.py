"""Synthetic molecular inputs for the configurations named in BASELINE.json / SURVEY.md section 8d.

The reference ships hcore/symm/CISD start vectors for Ne aug-cc-pVDZ, H2O cc-pVDZ and N2 cc-pVDZ but
every two-electron integral file is a stripped blob, so benchmarks and parity tests run on synthetic,
8-fold-symmetric, irrep-respecting integrals of the same dimensions.  The orbital irreps below are
the contents of Input_Data/<system>/symm.txt (data, not code); frozen-core orbitals come first.
"""
from __future__ import annotations

import numpy as np

# name -> (n_orb unfrozen, n_elec total, n_frozen electrons, irreps of ALL spatial orbitals)
SYSTEMS = {
    # Input_Data/Neon_augccpvdz/{sys_params,symm}.txt
    "ne": (22, 10, 2, [0, 0, 5, 6, 7, 0, 5, 6, 7, 0, 0, 1, 2, 3, 5, 6, 7, 0, 0, 0, 1, 2, 3]),
    # Input_Data/H2O_ccpvdz
    "h2o": (24, 10, 0, [0, 0, 3, 0, 2, 0, 3, 3, 0, 0, 2, 3, 1, 0, 2, 3, 0, 3, 0, 2, 1, 0, 0, 3]),
    # Input_Data/N2_ccpvdz
    "n2": (26, 14, 4, [0, 5, 0, 5, 0, 6, 7, 2, 3, 5, 0, 6, 7, 0, 2, 3, 5, 5, 0, 1, 6, 7, 4, 5, 0, 2, 3, 5]),
    # Input_Data/N2_str_ccpvdz
    "n2_str": (26, 14, 4, [0, 5, 0, 5, 0, 6, 7, 2, 3, 5, 6, 7, 5, 0, 2, 3, 0, 5, 0, 6, 7, 0, 1, 4, 5, 2, 3, 5]),
}


def tri_wdiag(i, j):
    """I_J_TO_TRI_WDIAG of FRIES/math_utils.h:17 for i <= j."""
    return j * (j + 1) // 2 + i


def make_integrals(n_tot: int, symm_all, seed: int):
    """hcore[n_tot, n_tot] and dense chemist-notation eris[n_tot]^4 = (ij|kl), 8-fold symmetric, zero
    unless the product of the four irreps is totally symmetric (SURVEY.md 8d recipe)."""
    rng = np.random.default_rng(seed)
    symm = np.asarray(symm_all, dtype=np.int64)
    T = n_tot
    ii, jj = np.triu_indices(T)  # pairs i <= j
    npair = ii.size
    g = rng.normal(0.0, 0.05, size=(npair, npair))
    g = np.triu(g) + np.triu(g, 1).T
    decay = np.exp(-0.3 * np.abs(ii - jj))
    g *= decay[:, None] * decay[None, :]
    psym = symm[ii] ^ symm[jj]
    g *= (psym[:, None] == psym[None, :])
    pair_of = np.zeros((T, T), dtype=np.int64)
    pair_of[ii, jj] = np.arange(npair)
    pair_of[jj, ii] = np.arange(npair)
    eris = g[pair_of[:, :, None, None], pair_of[None, None, :, :]].copy()
    idx = np.arange(T)
    eris[idx[:, None], idx[:, None], idx[None, :], idx[None, :]] += 0.3 / (1.0 + np.abs(idx[:, None] - idx[None, :]))
    h = rng.normal(0.0, 0.02, size=(T, T))
    h = (h + h.T) / 2
    h *= (symm[:, None] == symm[None, :])
    h[idx, idx] = -3.0 + 0.35 * idx
    return np.ascontiguousarray(h), np.ascontiguousarray(eris)


def pack_eris(eris_chem: np.ndarray) -> np.ndarray:
    """SymmERIs layout (FRIES/ndarr.hpp:206-244): data[TRI(p1, p2)], p = TRI(min, max) of an orbital pair."""
    T = eris_chem.shape[0]
    jj, ii = np.triu_indices(T)  # jj <= ii
    p = tri_wdiag(jj, ii)
    order = np.argsort(p)
    jj, ii = jj[order], ii[order]
    npair = ii.size
    a, b = np.triu_indices(npair)  # a <= b
    out = np.empty(npair * (npair + 1) // 2)
    out[tri_wdiag(a, b)] = eris_chem[jj[a], ii[a], jj[b], ii[b]]
    return out


def hf_det(n_orb: int, n_elec_unf: int) -> int:
    """gen_hf_bitstring FRIES/fci_utils.c:10-43: lowest n_elec/2 alpha and beta orbitals."""
    half = n_elec_unf // 2
    return ((1 << half) - 1) | (((1 << half) - 1) << n_orb)


class SynthMol:
    """Dimensions + synthetic integrals of one of SYSTEMS.  frozen=False drops the frozen core
    (FCIDUMP-style input of frisys_mol: n_frz = 0, NELEC = unfrozen electrons)."""

    def __init__(self, name, seed: int, frozen: bool = True):
        # name: a key of SYSTEMS, or a tuple (n_orb, n_elec_total, n_frozen, irreps) for a custom system
        n_orb, n_elec, n_frz, symm_all = SYSTEMS[name] if isinstance(name, str) else name
        symm_all = list(symm_all)
        if not frozen:
            symm_all = symm_all[n_frz // 2:]
            n_elec -= n_frz
            n_frz = 0
        self.name = name
        self.n_orb, self.n_elec_total, self.n_frz = n_orb, n_elec, n_frz
        self.n_elec = n_elec - n_frz
        self.tot_orb = n_orb + n_frz // 2
        self.symm_all = np.asarray(symm_all, dtype=np.uint8)
        self.symm = np.ascontiguousarray(self.symm_all[n_frz // 2:])
        self.hcore, self.eris_chem = make_integrals(self.tot_orb, symm_all, seed)
        self.eris_packed = pack_eris(self.eris_chem)
        self.hf = hf_det(n_orb, self.n_elec)
        self.n_bits = 2 * n_orb

    def random_dets(self, n: int, rng: np.random.Generator, irrep: int | None = 0) -> np.ndarray:
        """n distinct random determinants with n_elec/2 alpha + n_elec/2 beta electrons; irrep = the
        product irrep they must carry (None = any)."""
        M, h = self.n_orb, self.n_elec // 2
        out = set()
        symm = self.symm.astype(np.int64)
        while len(out) < n:
            m = max(1024, 2 * (n - len(out)))
            a = np.argsort(rng.random((m, M)), axis=1)[:, :h]
            b = np.argsort(rng.random((m, M)), axis=1)[:, :h]
            if irrep is not None:
                s = np.bitwise_xor.reduce(symm[a], axis=1) ^ np.bitwise_xor.reduce(symm[b], axis=1)
                ok = s == irrep
                a, b = a[ok], b[ok]
            keys = (1 << a.astype(np.uint64)).sum(axis=1) | ((1 << b.astype(np.uint64)).sum(axis=1) << np.uint64(M))
            for k in keys.tolist():
                out.add(k)
                if len(out) == n:
                    break
        return np.array(sorted(out), dtype=np.uint64)[rng.permutation(n)]

"""Checkpoint files of a FRI run: the reference's on-disk vector format and re-sharding to a
different number of ranks / GPUs.

Format (DistVec::save / load, FRIES/vec_utils.hpp:707-844; save_proc_hash / load_proc_hash, FRIES/io_utils.cpp): per rank r
  dets<r>.dat  curr_size x ceil(n_bits / 8) bytes, the determinants' bit strings (bit i = spin orbital i)
  vals<r>.dat  n_vecs rows of curr_size float64
and once  dense.txt  (sizes of the ranks' deterministic subspaces, comma separated; those determinants come first in their
rank's file) and  hash.dat  (n_bits uint32: the scrambler that decides ownership, owner = hash_fxn(occ; scrambler) % ranks,
det_hash.hpp:160-170, vec_utils.hpp:373-379).  The reference can only restart on the rank count it was saved with (every
rank reads its own file); `reshard` redistributes the determinants by the owner rule, so that a run saved on M ranks
restarts on N ranks or on N GPUs.  Host-side file tooling in numpy; nothing here is on the GPU hot path."""
from __future__ import annotations

import os
import shutil

import numpy as np

HASH_PRIME = np.uint64(1099511628211)


def det_hash(keys: np.ndarray, scrambler: np.ndarray) -> np.ndarray:
    """HashTable::hash_fxn over the occupied orbitals in ascending order (det_hash.hpp:160-170), vectorised over keys:
    h <- prime * h + uint32((i + 1) * scrambler[orbital]) for the i-th occupied orbital."""
    keys = np.ascontiguousarray(keys, np.uint64)
    h = np.zeros(keys.shape, np.uint64)
    i = np.zeros(keys.shape, np.uint64)
    with np.errstate(over="ignore"):
        for b, s in enumerate(np.asarray(scrambler, np.uint64)):
            occ = ((keys >> np.uint64(b)) & np.uint64(1)).astype(bool)
            if not occ.any():
                continue
            i = i + occ.astype(np.uint64)
            term = (i * s) & np.uint64(0xFFFFFFFF)
            h = np.where(occ, HASH_PRIME * h + term, h)
    return h


def owners(keys, scrambler, n_ranks: int) -> np.ndarray:
    return (det_hash(keys, scrambler) % np.uint64(n_ranks)).astype(np.int64)


def n_saved_ranks(path: str) -> int:
    n = 0
    while os.path.exists(f"{path}dets{n}.dat"):
        n += 1
    return n


def read_rank(path: str, rank: int, n_bits: int, n_vecs: int):
    nb = (n_bits + 7) // 8
    raw = np.fromfile(f"{path}dets{rank}.dat", np.uint8)
    n = raw.size // nb
    pad = np.zeros((n, 8), np.uint8)
    pad[:, :nb] = raw[:n * nb].reshape(n, nb)
    keys = pad.view("<u8").reshape(n)
    vals = np.fromfile(f"{path}vals{rank}.dat", np.float64)
    if vals.size != n * n_vecs:
        raise ValueError(f"{path}vals{rank}.dat holds {vals.size} values, expected {n_vecs} rows of {n}")
    return keys.copy(), vals.reshape(n_vecs, n).copy()


def read_dense_sizes(path: str, n_ranks: int):
    try:
        txt = open(f"{path}dense.txt").read().replace("\n", ",")
        sizes = [int(x) for x in txt.split(",") if x.strip()]
    except OSError:
        sizes = []
    return (sizes + [0] * n_ranks)[:n_ranks]


def write_rank(path: str, rank: int, keys, vals, n_bits: int):
    nb = (n_bits + 7) // 8
    k = np.ascontiguousarray(keys, "<u8")
    k.view(np.uint8).reshape(-1, 8)[:, :nb].tofile(f"{path}dets{rank}.dat")
    np.ascontiguousarray(vals, np.float64).tofile(f"{path}vals{rank}.dat")


def reshard(src: str, dst: str, n_bits: int, n_vecs: int, n_new: int) -> dict:
    """src / dst are the string prefixes the drivers take as --result_dir / --load_dir (ending in '/')."""
    n_old = n_saved_ranks(src)
    if n_old == 0:
        raise FileNotFoundError(f"no {src}dets0.dat")
    scr = np.fromfile(f"{src}hash.dat", np.uint32)
    if scr.size < n_bits:
        raise ValueError(f"{src}hash.dat holds {scr.size} entries, expected {n_bits}")
    scr = scr[:n_bits]
    dense_old = read_dense_sizes(src, n_old)
    keys, vals, dense = [], [], []
    for r in range(n_old):
        k, v = read_rank(src, r, n_bits, n_vecs)
        o = owners(k, scr, n_old)
        if k.size and not np.all(o == r):
            raise ValueError(f"{src}dets{r}.dat holds determinants that rank {r} of {n_old} does not own (wrong hash.dat?)")
        keys.append(k)
        vals.append(v)
        d = np.zeros(k.size, bool)
        d[:dense_old[r]] = True
        dense.append(d)
    keys = np.concatenate(keys)
    vals = np.concatenate(vals, axis=1)
    dense = np.concatenate(dense)
    # the reference's files include freed slots (stale key, every value 0; a live copy of the key may exist elsewhere):
    # they carry nothing -- DistVec::load skips them too (vec_utils.hpp:812-822)
    live = dense | np.any(vals != 0, axis=0)
    keys, vals, dense = keys[live], vals[:, live], dense[live]
    if np.unique(keys).size != keys.size:
        raise ValueError(f"{src}: a determinant is stored twice with nonzero values")
    own = owners(keys, scr, n_new)
    os.makedirs(os.path.dirname(dst) or ".", exist_ok=True)
    for f in os.listdir(os.path.dirname(src) or "."):  # shift, norms, energies, parameters: unchanged
        full = os.path.join(os.path.dirname(src) or ".", f)
        if os.path.isfile(full) and not (f.startswith("dets") or f.startswith("vals")) and f != "dense.txt":
            if os.path.abspath(full) != os.path.abspath(os.path.join(os.path.dirname(dst) or ".", f)):
                shutil.copy(full, os.path.join(os.path.dirname(dst) or ".", f))
    sizes, dense_new = [], []
    for r in range(n_new):
        sel = np.flatnonzero(own == r)
        sel = np.concatenate([sel[dense[sel]], sel[~dense[sel]]])  # the deterministic subspace stays in front
        write_rank(dst, r, keys[sel], vals[:, sel], n_bits)
        sizes.append(int(sel.size))
        dense_new.append(int(dense[sel].sum()))
    with open(f"{dst}dense.txt", "w") as f:
        f.write(",".join(str(x) for x in dense_new) + "\n")
    return {"ranks_in": n_old, "ranks_out": n_new, "determinants": int(keys.size), "per_rank": sizes, "dense": dense_new}


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="re-shard a FRI checkpoint to another number of ranks / GPUs")
    ap.add_argument("src", help="prefix the run was saved with (--result_dir), e.g. run1/")
    ap.add_argument("dst", help="prefix to write (--load_dir of the restart), e.g. run1_8gpu/")
    ap.add_argument("--n_orb", type=int, required=True, help="unfrozen spatial orbitals (determinants have 2 n_orb bits)")
    ap.add_argument("--n_vecs", type=int, default=2, help="value rows saved (2 for frisys_mol / frifull_mol)")
    ap.add_argument("--ranks", type=int, required=True)
    a = ap.parse_args(argv)
    print(reshard(a.src, a.dst, 2 * a.n_orb, a.n_vecs, a.ranks))


if __name__ == "__main__":
    main()

"""One rank of tests/test_mpi_shim.py::test_reference_compression_on_several_ranks (started by oracle/mpi_shim/shimrun.py):
the REFERENCE's find_preserve + sys_comp and piv_comp_parallel on this rank's contiguous shard of a seeded vector, with
its collectives running between the ranks.  Writes rank<r>.npz into the directory given as argv[1]."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import reflib  # noqa: E402


def shard_bounds(n, world):
    return np.linspace(0, n, world + 1).astype(int)


def make_vector(n):
    rng = np.random.default_rng(5)
    v = np.concatenate([rng.lognormal(6, 1, n // 100), rng.lognormal(-3, 2, n - n // 100)])
    rng.shuffle(v)
    return v * rng.choice([-1.0, 1.0], n)


def main():
    out_dir, n, budget, rn = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
    L = reflib.lib()
    world = L.ref_mpi_init()
    rank = L.ref_mpi_rank()
    b = shard_bounds(n, world)
    v = make_vector(n)[b[rank]:b[rank + 1]].copy()
    # find_preserve (collective: sum_mpi per round) + sys_comp (loc_norms all-gathered by the caller, frisys_mol.cpp:532)
    keep = np.zeros(v.size, np.uint8)
    ns, gn = C.c_uint(budget), C.c_double(0)
    loc = L.ref_find_preserve(v, v.size, C.byref(ns), C.byref(gn), keep)
    # the drivers all-gather the residual norms before sys_comp (frisys_mol.cpp:532); sys_comp overwrites its copy with the
    # one-norms after resampling (compress_utils.cpp:322-326)
    norms = np.zeros(world)
    norms[rank] = loc
    L.ref_allgather_doubles(norms, 1)
    sv, sk, after = v.copy(), keep.copy(), norms.copy()
    L.ref_sys_comp(sv, sv.size, after.ctypes.data_as(C.POINTER(C.c_double)), ns.value, sk, rn)
    # piv_comp_parallel on a fresh copy; every rank seeds its own generator
    pv, pk = v.copy(), np.zeros(v.size, np.uint8)
    used = L.ref_piv_comp_parallel(pv, pv.size, budget, pk, 100 + rank)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), keep=keep, loc=loc, glob=gn.value, left=ns.value, norms=norms, norms_after=after, sv=sv, sk=sk,
             pv=pv, pk=pk, used=used)


if __name__ == "__main__":
    main()

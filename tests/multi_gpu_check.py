"""2+ GPU checks, launched as:  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu_check.py

(1) distributed find_preserve + sys_comp on a contiguous split of one array must give the single-rank result of the
    oracle on the concatenation (kept set exact; sampled set up to counted FP-boundary ties);
(2) a few frisys_mol iterations over N ranks: every stored determinant lives on its hash owner, the projected
    energy is finite and the global one-norm matches a single-GPU run of the same configuration statistically."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import fries_b200
    import oraclelib
    from fries_b200._capi import FrisysParams, check, lib
    from fries_b200.multi import Comm, MultiGpuFrisys
    from fries_b200.synth import SynthMol

    ctx = fries_b200.Context(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    check(lib.fries_ctx_set_stream(ctx.h, stream.cuda_stream))
    dev = torch.device("cuda", local)
    comm = Comm(ctx, dist, rank, world, dev)
    check(lib.fries_ctx_set_comm(ctx.h, comm.h))

    # ---- (0) cost of one in-kernel exchange ----
    for ctas, n in ((1, 1), (1, 12), (0, 1), (0, 12)):
        us = C.c_double(0)
        check(lib.fries_comm_pingpong(comm.h, ctas, 200, n, C.byref(us)))
        if rank == 0:
            print(f"[multi] exchange of {n} doubles + grid barrier, {'1 CTA' if ctas else 'full grid'}: {us.value:.2f} us", flush=True)

    # ---- (1) distributed vector compression ----
    rng = np.random.default_rng(5)
    n, budget = 400000, 60000
    v = np.concatenate([rng.lognormal(6, 1, n // 100), rng.lognormal(-3, 2, n - n // 100)])
    rng.shuffle(v)
    v *= rng.choice([-1.0, 1.0], n)
    bounds = np.linspace(0, n, world + 1).astype(int)
    lo, hi = bounds[rank], bounds[rank + 1]
    o_loc, o_glob, o_left, o_keep = oraclelib.find_preserve(v, budget)
    for rn in (0.31, 0.87):
        ov, ok, _ = oraclelib.sys_comp(v, [o_loc], o_left, o_keep, rn)
        d_vals = torch.tensor(v[lo:hi], device=dev)
        d_keep = torch.zeros(hi - lo, dtype=torch.uint8, device=dev)
        d_r4 = torch.zeros(4, dtype=torch.float64, device=dev)
        d_nn = torch.zeros(1, dtype=torch.float64, device=dev)
        check(lib.fries_find_preserve_dev(ctx.h, d_vals.data_ptr(), hi - lo, budget, d_keep.data_ptr(), d_r4.data_ptr()))
        ctx.sync()
        keep = d_keep.cpu().numpy()
        r4 = d_r4.cpu().numpy()
        assert np.array_equal(keep, o_keep[lo:hi]), f"rank {rank}: kept set differs in {np.sum(keep != o_keep[lo:hi])}"
        assert int(r4[2]) == o_left and abs(r4[1] - o_glob) <= 1e-12 * o_glob
        check(lib.fries_sys_comp_dev(ctx.h, d_vals.data_ptr(), hi - lo, d_r4.data_ptr(), d_keep.data_ptr(), rn,
                                     d_nn.data_ptr()))
        ctx.sync()
        dele = d_keep.cpu().numpy()
        ties = int(np.sum(dele != ok[lo:hi]))
        t = torch.tensor([ties], device=dev)
        dist.all_reduce(t)
        assert int(t.item()) <= 2, f"{int(t.item())} sampled-set mismatches over {world} ranks"
        same = dele == ok[lo:hi]
        assert np.allclose(d_vals.cpu().numpy()[same], ov[lo:hi][same], rtol=1e-12, atol=0)
    if rank == 0:
        print(f"[multi] distributed find_preserve/sys_comp over {world} ranks == single-rank oracle", flush=True)
    # ---- (1b) distributed pivotal compression: every rank's shard against the oracle's chain piv_budget ->
    #      adjust_probs -> piv_samp_serial (each pinned against the reference) on the same norms and draws ----
    from fries_b200.multi import piv_comp_parallel
    from test_hostcheck_piv import same_up_to_closing_unit
    shared = oraclelib.mt19937(1234, 2 * world + 8)
    local = oraclelib.mt19937(77 + rank, 2 * (budget + world) + 8)  # the C-ABI wants room for 2 * (compress_size + n_ranks)
    gv, gk, n_drawn, (keep, norms, n_left, used_b), new_norm = piv_comp_parallel(ctx, dist, rank, world, v[lo:hi], budget,
                                                                                shared, local, dev)
    assert np.array_equal(keep, o_keep[lo:hi]) and n_left == o_left
    mine = np.abs(v[lo:hi][keep == 0]).sum()
    assert abs(norms[rank] - mine) <= 1e-12 * mine
    budgets, ub = oraclelib.piv_budget(norms, n_left, shared)
    assert ub == used_b and int(budgets.sum()) == n_left
    glob = float(np.sum(norms))
    exp_loc = n_left * norms[rank] / glob
    av, ak, a_n, a_norm = oraclelib.adjust_probs(v[lo:hi], int(budgets[rank]), exp_loc, n_left, glob, keep)
    ov2, ok2, o_used = oraclelib.piv_samp_serial(av, a_norm, a_n, ak, local)
    assert n_drawn == a_n == o_used // 2, (n_drawn, a_n, o_used)
    # elements that adjust_probs made exact are preserved for the sampler: they are not candidates of a sampling unit
    same_up_to_closing_unit(av, ak, a_norm, a_n, gv, gk, ov2, ok2, rtol=1e-10)
    tot = torch.tensor([float((gv != 0).sum())], device=dev, dtype=torch.float64)
    dist.all_reduce(tot)
    assert int(tot.item()) == budget, (int(tot.item()), budget)
    if rank == 0:
        print(f"[multi] distributed piv_comp_parallel over {world} ranks: budgets {budgets.tolist()} == oracle chain, "
              f"{budget} elements in total", flush=True)
    if os.environ.get("FRIES_MULTI_ONLY") == "compress":
        check(lib.fries_ctx_set_comm(ctx.h, None))
        comm.close()
        dist.barrier()
        dist.destroy_process_group()
        return
    check(lib.fries_ctx_set_comm(ctx.h, None))
    comm.close()

    # ---- (2) frisys_mol iterations ----
    sm = SynthMol("ne", 2, frozen=False)
    mol = fries_b200.Mol.from_synth(ctx, sm)
    rs = np.random.RandomState(0)
    proc_scr = rs.randint(0, 2**32, sm.n_bits, dtype=np.uint64).astype(np.uint32)
    vec_scr = rs.randint(0, 2**32, sm.n_bits, dtype=np.uint64).astype(np.uint32)
    hf = np.array([sm.hf], np.uint64)
    hf_en = float(mol.diag(hf)[0])
    tmp = fries_b200.Vec(ctx, 1 << 16, sm.n_bits, sm.n_elec, 2, proc_scr, vec_scr)
    tmp.set_diag_mol(mol, hf_en)
    tmp.add(hf, np.ones(1), np.ones(1, np.uint8))
    tmp.h_apply(mol, 0, 1, 1.0, -0.5)
    k1, v1 = tmp.download()
    tmp.h_apply(mol, 0, 1, 0.0, 1.0)
    hk, hv = tmp.download()
    tmp.close()
    p_doub = int(mol.doub_ex(hf)[0][-1]) / (int(mol.doub_ex(hf)[0][-1]) + int(mol.sing_ex(hf)[0][-1]))
    keys, vals = k1, v1[1] * (20000.0 / np.abs(v1[1]).sum())
    _, owner = fries_b200.hash_owner(ctx, keys, proc_scr, world)
    mat_nonz, vec_nonz = 40000, 30000
    norms = {}
    for route in ("p2p", "nccl"):
        eng = MultiGpuFrisys(ctx, dist, rank, world, mol, sm, 200000, 4 * mat_nonz // world,
                             2 * mat_nonz // (world * world) + 4096, proc_scr, vec_scr, hf_en, (hf, np.ones(1)),
                             (hk, hv[1]), route=route)
        eng.load(keys, vals, owner)
        params = FrisysParams(eps=0.001, init_thresh=1.0, p_doub=p_doub, new_hb=1, matr_samp=mat_nonz,
                              target_nonz=vec_nonz, en_shift=0.0)
        uni = np.random.RandomState(1).random_sample((30, 6))
        for it in range(30):
            st = eng.iterate(params, uni[it])
        lk, lv = eng.vec.download()
        _, own = fries_b200.hash_owner(ctx, lk, proc_scr, world)
        assert np.all(own == rank), f"rank {rank} stores {np.sum(own != rank)} determinants it does not own"
        assert eng.comm.error_epoch() == 0
        nloc = torch.tensor([lk.size], device=dev)
        dist.all_reduce(nloc)
        assert int(nloc.item()) == st.curr_size, (int(nloc.item()), st.curr_size)
        en = st.numer / st.denom
        assert np.isfinite(en) and -1.0 < en < 0.0, en
        assert st.n_spawned > 0.9 * mat_nonz, st.n_spawned
        norms[route] = st.glob_norm
        if rank == 0:
            print(f"[multi] {world} ranks, route {route}, 30 iterations: stored={st.curr_size} norm={st.glob_norm:.3f} "
                  f"energy={en:.6f} spawned/iter={st.n_spawned}", flush=True)
        eng.close()
    # same uniforms, same elements delivered: the two routes differ only in arrival (= storage) order
    assert abs(norms["p2p"] - norms["nccl"]) <= 2e-3 * norms["nccl"], norms
    # ---- (3) deterministic H.v and frifull_mol iterations on the partitioned vector (direct route) ----
    from fries_b200._capi import FrifullParams
    single = fries_b200.Vec(ctx, 1 << 21, sm.n_bits, sm.n_elec, 2, proc_scr, vec_scr)
    single.set_diag_mol(mol, hf_en)
    par = np.argsort(-np.abs(vals), kind="stable")[:3000]
    pk, pv = np.ascontiguousarray(keys[par]), np.ascontiguousarray(vals[par])
    single.upload(pk, np.stack([pv, np.zeros_like(pv)]))
    n_single = single.h_apply(mol, 0, 1, 1.0, -0.01)
    sk, sv = single.download()
    single.close()
    # segments far smaller than the spawn count: several routed windows per H.v
    eng = MultiGpuFrisys(ctx, dist, rank, world, mol, sm, 1 << 21, 1 << 16, 60000, proc_scr, vec_scr, hf_en,
                         (hf, np.ones(1)), (hk, hv[1]), route="p2p")
    _, own = fries_b200.hash_owner(ctx, pk, proc_scr, world)
    eng.load(pk, pv, own)
    n_loc = eng.h_apply(0, 1, 1.0, -0.01)
    n_all = torch.tensor([n_loc], device=dev)
    dist.all_reduce(n_all)
    assert int(n_all.item()) == n_single, (int(n_all.item()), n_single)
    lk, lv = eng.vec.download()
    _, own = fries_b200.hash_owner(ctx, lk, proc_scr, world)
    assert np.all(own == rank)
    ref = dict(zip(sk.tolist(), sv[1].tolist()))
    mine = np.array([ref[k] for k in lk.tolist()])
    assert np.allclose(lv[1], mine, rtol=1e-12, atol=1e-13), np.abs(lv[1] - mine).max()
    n_glob = torch.tensor([lk.size], device=dev)
    dist.all_reduce(n_glob)
    assert int(n_glob.item()) == sk.size, (int(n_glob.item()), sk.size)
    # a few frifull_mol iterations: energy estimator finite, sizes consistent
    fp = FrifullParams(eps=0.005, target_nonz=2000, en_shift=0.0, adjust_shift=0, damp_factor=0.05, target_norm=0.0,
                       last_one_norm=0.0)
    for it in range(4):
        st = eng.frifull_iterate(fp, float(uni[it, 0]))
    lk, lv = eng.vec.download()
    n_glob = torch.tensor([lk.size], device=dev)
    dist.all_reduce(n_glob)
    assert int(n_glob.item()) == st.curr_size and st.n_spawned > 0 and np.isfinite(st.numer / st.denom)
    assert eng.comm.error_epoch() == 0
    if rank == 0:
        print(f"[multi] {world} ranks: routed H.v == single-GPU H.v ({n_single} connections, {sk.size} determinants); "
              f"frifull_mol 4 iterations: stored={st.curr_size} spawned={st.n_spawned} energy={st.numer / st.denom:.6f}",
              flush=True)
    eng.close()
    # ---- (3b) the same from ONE stored determinant: every rank but the owner starts empty, and must still take part in
    # every collective of the routed H.v (round count, publish / wait / merge, barrier) -- frifull_mol's normal start ----
    single = fries_b200.Vec(ctx, 1 << 21, sm.n_bits, sm.n_elec, 2, proc_scr, vec_scr)
    single.set_diag_mol(mol, hf_en)
    single.upload(hf, np.array([[1.0], [0.0]]))
    n_single = single.h_apply(mol, 0, 1, 1.0, -0.01)
    sk, sv = single.download()
    single.close()
    eng = MultiGpuFrisys(ctx, dist, rank, world, mol, sm, 1 << 21, 1 << 16, 60000, proc_scr, vec_scr, hf_en,
                         (hf, np.ones(1)), (hk, hv[1]), route="p2p")
    _, own = fries_b200.hash_owner(ctx, hf, proc_scr, world)
    eng.load(hf, np.ones(1), own)
    n_loc = eng.h_apply(0, 1, 1.0, -0.01)
    n_all = torch.tensor([n_loc], device=dev)
    dist.all_reduce(n_all)
    assert int(n_all.item()) == n_single, (int(n_all.item()), n_single)
    lk, lv = eng.vec.download()
    ref = dict(zip(sk.tolist(), sv[1].tolist()))
    mine = np.array([ref[k] for k in lk.tolist()])
    assert np.allclose(lv[1], mine, rtol=1e-12, atol=1e-13)
    n_glob = torch.tensor([lk.size], device=dev)
    dist.all_reduce(n_glob)
    assert int(n_glob.item()) == sk.size, (int(n_glob.item()), sk.size)
    assert eng.comm.error_epoch() == 0
    if rank == 0:
        print(f"[multi] {world} ranks: routed H.v from one stored determinant == single-GPU H.v ({n_single} connections)",
              flush=True)
    eng.close()
    mol.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

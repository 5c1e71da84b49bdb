"""CPU tier (gloo, world_size 2): the host-side plumbing of the multi-GPU path -- handle all-gather, owner
partitioning, and the counts + fixed-capacity-segment all-to-all of the Router -- with CPU tensors."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oraclelib
        from fries_b200.multi import Router, gather_bytes, owned_slice
        # (a) fixed-size byte all-gather (IPC handles)
        mine = bytes([(rank * 7 + i) % 256 for i in range(64)])
        got = gather_bytes(dist, mine, world, torch.device("cpu"))
        assert got[rank] == mine and all(got[r] == bytes([(r * 7 + i) % 256 for i in range(64)]) for r in range(world))
        # (b) route spawned elements to their hash owner with the Router's layout
        rng = np.random.default_rng(100 + rank)
        scr = np.random.default_rng(0).integers(0, 2**32, 52, dtype=np.uint64).astype(np.uint32)
        keys = rng.integers(0, 2**52, 5000, dtype=np.uint64)
        vals = rng.normal(size=5000)
        _, owner = oraclelib.hash_keys(keys, scr, world)  # the checker plays the device's owner function here
        seg_cap = 4096
        r = Router(dist, world, seg_cap, torch.device("cpu"))
        sb = r.send_buf.numpy().view(np.uint64)
        for p in range(world):
            idx = owned_slice(owner, p)
            r.send_counts[p] = idx.size
            sb[p, : idx.size] = keys[idx]
            sb[p, seg_cap: seg_cap + idx.size] = vals[idx].view(np.uint64)
        r.exchange()
        rb = r.recv_buf.numpy().view(np.uint64)
        total = 0.0
        for p in range(world):
            n = int(r.recv_counts[p])
            rk = rb[p, :n]
            _, own = oraclelib.hash_keys(np.ascontiguousarray(rk), scr, world)
            assert np.all(own == rank)
            total += rb[p, seg_cap: seg_cap + n].view(np.float64).sum()
        # conservation: the sum of all routed values equals the sum of all generated values
        t = torch.tensor([total, vals.sum()], dtype=torch.float64)
        dist.all_reduce(t)
        assert abs(t[0] - t[1]) < 1e-9
        # (c) the budget step of the distributed pivotal compression (multi.piv_comp_parallel): the residual norms are
        # all-gathered, every rank evaluates rank 0's arithmetic (compress_utils.cpp:560-608) on the SAME shared draws
        # with the library's host function, and all ranks must arrive at the same budgets and the same draw count
        import fries_b200
        shared = oraclelib.mt19937(77, 2 * world + 8)
        for n_left, my_norm in ((1000, 3.0 + 2.5 * rank), (7, 0.125 * (rank + 1)), (5, 0.0)):
            g = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(g, torch.tensor([my_norm], dtype=torch.float64))
            norms = np.array([float(x.item()) for x in g])
            budgets, used = fries_b200.piv_budget(norms, n_left, shared)
            ob, oused = oraclelib.piv_budget(norms, n_left, shared)
            assert used == oused and np.abs(budgets.astype(int) - ob.astype(int)).sum() <= 2
            if norms.sum() > 0:
                assert int(budgets.sum()) == n_left
                assert np.all(np.abs(budgets - norms / norms.sum() * n_left) < 1 + 1e-9)
            else:
                assert not budgets.any() and used == 2 * min(world, n_left)
            t = torch.tensor(np.concatenate([budgets.astype(np.float64), [float(used)]]))
            lo, hi = t.clone(), t.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            assert torch.equal(lo, hi)  # identical on every rank
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_router_and_handle_gather_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res

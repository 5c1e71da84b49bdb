"""Multi-GPU host plumbing: one process per GPU, `torch.distributed` for the rendezvous and the all-to-all.

What runs where:
  * owner partition  hash_fxn(occ; proc_scrambler) % n_ranks        -- device (finalize kernel packs by owner)
  * all-to-all of spawned (key | ini, value) pairs                   -- NCCL `all_to_all_single`, issued here
  * every global scalar reduction (sum_mpi, loc_norms Allgather)     -- inside the kernels, over peer-mapped
                                                                        inboxes (csrc/comm.cuh); no NCCL call
This module is plumbing only: buffers are torch tensors, all arithmetic is in libfries_b200.so.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from ._capi import FrisysParams, IterStats, arr, check, lib, ptr


def gather_bytes(dist, payload: bytes, world: int, device) -> list[bytes]:
    """all-gather a fixed-size byte string (IPC handles) with whatever backend the group uses"""
    t = torch.tensor(list(payload), dtype=torch.uint8, device=device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [bytes(o.cpu().tolist()) for o in out]


def bind_stream(ctx):
    """One stream for the library's kernels and torch's own work (NCCL collectives, fills, copies): the context is bound to
    torch's CURRENT stream.  A fries_ctx otherwise runs on a private non-blocking stream (csrc/ctx.cu), and an
    all_to_all_single or a torch.zeros issued on torch's stream would not be ordered against the kernels around it."""
    check(lib.fries_ctx_set_stream(ctx.h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))


class Comm:
    """peer-mapped inboxes for the in-kernel cross-rank reductions (fries_comm)"""

    def __init__(self, ctx, dist, rank: int, world: int, device):
        self.ctx, self.rank, self.world = ctx, rank, world
        bind_stream(ctx)
        h = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        check(lib.fries_comm_create(ctx.h, world, rank, C.byref(h), handle))
        self.h = h
        handles = gather_bytes(dist, bytes(handle), world, device)
        blob = b"".join(handles)
        check(lib.fries_comm_connect(self.h, blob))
        dist.barrier()

    def open_route(self, dist, seg_cap: int, device):
        """receive window of the direct spawn route (fries_comm_route_create / _connect)"""
        handle = (C.c_uint8 * 64)()
        check(lib.fries_comm_route_create(self.h, seg_cap, handle))
        blob = b"".join(gather_bytes(dist, bytes(handle), self.world, device))
        check(lib.fries_comm_route_connect(self.h, blob))
        dist.barrier()

    def error_epoch(self) -> int:
        e = C.c_uint64(0)
        check(lib.fries_comm_error(self.h, C.byref(e)))
        return e.value

    def close(self):
        if self.h:
            lib.fries_comm_destroy(self.h)
            self.h = None


def piv_comp_parallel(ctx, dist, rank: int, world: int, values_local, compress_size: int, shared_draws, local_draws,
                      device):
    """piv_comp_parallel (compress_utils.cpp:354-387) of a vector partitioned over the ranks.

    ctx must carry the ranks' inboxes (fries_ctx_set_comm): find_preserve is collective with its reductions inside the
    kernel.  The residual norms are all-gathered (the reference's MPI_Allgather, :365), every rank evaluates rank 0's
    budget arithmetic on the SAME shared_draws (the reference draws on rank 0 and scatters, :560-608), then adjusts and
    samples its own shard with local_draws.  Returns (values, delete flags, budget of this rank, norms before, one-norm
    after); plumbing only, the arithmetic is in libfries_b200.so."""
    from .api import piv_budget
    bind_stream(ctx)  # the fills below run on torch's stream, the compression on the context's: make them one
    v = np.array(values_local, np.float64)  # a copy: the compression works in place
    d_vals = torch.tensor(v, device=device)
    d_keep = torch.zeros(max(v.size, 1), dtype=torch.uint8, device=device)
    d_r4 = torch.zeros(4, dtype=torch.float64, device=device)
    check(lib.fries_find_preserve_dev(ctx.h, d_vals.data_ptr(), v.size, compress_size, d_keep.data_ptr(), d_r4.data_ptr()))
    ctx.sync()
    r4 = d_r4.cpu().numpy()
    n_left = int(r4[2])
    norms = [torch.zeros(1, dtype=torch.float64, device=device) for _ in range(world)]
    dist.all_gather(norms, d_r4[0:1].clone())
    norms = np.array([float(t.item()) for t in norms])
    keep = d_keep.cpu().numpy()[:v.size]
    shared = arr(shared_draws, np.uint32)
    used_b = 0
    if n_left:
        _, used_b = piv_budget(norms, n_left, shared)
    draws = np.concatenate([shared[:used_b], arr(local_draws, np.uint32)])
    if draws.size < 2 * (compress_size + world):
        raise ValueError("piv_comp_parallel needs 2 * compress_size local draws")
    ln = norms.copy()
    k = keep.copy()
    used = C.c_size_t(0)
    check(lib.fries_piv_comp(ctx.h, ptr(v), v.size, compress_size, ptr(k), ptr(draws), C.byref(used), ptr(ln), world, rank,
                             1, n_left))
    n_drawn = (used.value - used_b) // 2
    return v, k, n_drawn, (keep, norms, n_left, used_b), float(ln[rank])


class Router:
    """the Adder's exchange (vec_utils.hpp:991-1019): counts, then fixed-capacity segments"""

    def __init__(self, dist, world: int, seg_cap: int, device):
        self.dist, self.world, self.seg_cap = dist, world, seg_cap
        self.send_buf = torch.zeros((world, 2 * seg_cap), dtype=torch.int64, device=device)
        self.recv_buf = torch.zeros((world, 2 * seg_cap), dtype=torch.int64, device=device)
        self.send_counts = torch.zeros(world + 1, dtype=torch.int64, device=device)  # last = overflow
        self.recv_counts = torch.zeros(world, dtype=torch.int64, device=device)

    def exchange(self):
        self.dist.all_to_all_single(self.recv_counts, self.send_counts[: self.world].contiguous())
        self.dist.all_to_all_single(self.recv_buf, self.send_buf)

    @property
    def bytes_sent_per_exchange(self) -> int:
        return self.send_buf.numel() * 8 + self.world * 8


def owned_slice(owner: np.ndarray, rank: int):
    """indices of the elements whose owner is `rank` (DistVec keeps only those)"""
    return np.nonzero(owner == rank)[0]


class MultiGpuFrisys:
    """frisys_mol loop body over n_ranks GPUs.  route = "p2p" (default): the spawn kernel stores every element straight
    into its owner's receive window over NVLink and _finish waits on epoch flags -- no collective, no host round trip;
    route = "nccl": spawn -> all_to_all_single (counts, then fixed-capacity segments) -> finish."""

    def __init__(self, ctx, dist, rank, world, mol, sm, max_dets_local, spawn_cap_local, seg_cap, proc_scr, vec_scr,
                 hf_en, trial, htrial, route="p2p"):
        from .api import Vec
        self.ctx, self.dist, self.rank, self.world, self.route = ctx, dist, rank, world, route
        device = torch.device("cuda", torch.cuda.current_device())
        self.comm = Comm(ctx, dist, rank, world, device)
        self.seg_cap = seg_cap
        self.vec = Vec(ctx, max_dets_local, sm.n_bits, sm.n_elec, 2, proc_scr, vec_scr, world, rank)
        self.vec.set_diag_mol(mol, hf_en)
        self.vec.frisys_setup(mol, spawn_cap_local, trial[0], trial[1], htrial[0], htrial[1])
        self.mol = mol
        if route == "p2p":
            self.router = None
            self.comm.open_route(dist, seg_cap, device)
            check(lib.fries_hbpp_set_route_p2p(self.vec.hb, self.comm.h))
        else:
            self.router = r = Router(dist, world, seg_cap, device)
            check(lib.fries_hbpp_set_route(self.vec.hb, self.comm.h, r.send_buf.data_ptr(), r.recv_buf.data_ptr(),
                                           r.send_counts.data_ptr(), seg_cap))

    def load(self, keys, vals, owner):
        idx = owned_slice(owner, self.rank)
        k = np.ascontiguousarray(keys[idx])
        v = np.ascontiguousarray(vals[idx])
        self.vec.upload(k, np.stack([v, np.zeros_like(v)]))
        return idx.size

    def h_apply(self, src, dest, id_fac, h_fac) -> int:
        """deterministic H.v on the partitioned vector (direct route only); returns this rank's spawn count"""
        n = C.c_uint64(0)
        check(lib.fries_h_apply_routed(self.vec.h, self.mol.h, self.vec.hb, src, dest, id_fac, h_fac, C.byref(n)))
        return n.value

    def frifull_iterate(self, params, uniform: float) -> IterStats:
        """frifull_mol.cpp:256-320 over the ranks (direct route only)"""
        st = IterStats()
        check(lib.fries_frifull_mol_iterate(self.vec.h, self.mol.h, self.vec.hb, C.byref(params), uniform, C.byref(st)))
        return st

    def iterate(self, params: FrisysParams, uniforms6) -> IterStats:
        u = arr(uniforms6, np.float64)
        check(lib.fries_frisys_mol_spawn(self.vec.h, self.mol.h, self.vec.hb, C.byref(params), ptr(u)))
        st = IterStats()
        if self.router is None:
            check(lib.fries_frisys_mol_finish(self.vec.h, self.mol.h, self.vec.hb, C.byref(params), ptr(u), None,
                                              C.byref(st)))
        else:
            self.router.exchange()
            check(lib.fries_frisys_mol_finish(self.vec.h, self.mol.h, self.vec.hb, C.byref(params), ptr(u),
                                              self.router.recv_counts.data_ptr(), C.byref(st)))
        return st

    @property
    def route_description(self):
        if self.router is None:
            return {"collective": "none: the spawn kernel stores into the owners' peer-mapped receive windows over "
                                  "NVLink (csrc/comm.cuh RouteView), epoch flags instead of a barrier",
                    "bytes_per_rank_per_step": "16 B per spawned element (only what was spawned)"}
        return {"collective": "nccl all_to_all_single (counts + fixed-capacity segments)",
                "bytes_per_rank_per_step": self.router.bytes_sent_per_exchange}

    def close(self):
        self.vec.close()
        self.comm.close()


def run_multi_gpu_bench(args, cfg, ctx, dist, rank, world, local, prepare_workload, ClockSampler, cpu_baseline_block):
    """bench.py body for --gpus N > 1, one JSON line on rank 0.  Strong scaling for the molecular configurations (the
    vector and the sample budget are those of the configuration at every N: BASELINE.json configs[2] is quoted "at 1/2/4/8
    B200"); the synthetic configuration (configs[4], `synthetic_vector`) is weak-scaled: its sizes are per GPU."""
    import json

    from .api import hash_owner

    stream = torch.cuda.current_stream()
    gcfg = dict(cfg)
    synthetic = bool(cfg.get("synthetic_vector"))
    weak = synthetic
    if weak:
        gcfg["vec_nonz"] = cfg["vec_nonz"] * world
        gcfg["mat_nonz"] = cfg["mat_nonz"] * world
        gcfg["target"] = cfg["target"] * world
        gcfg["max_dets"] = cfg["max_dets"] * world
    # capacity of one rank's store: its share plus head-room for the imbalance of the owner hash
    max_dets_local = cfg["max_dets"] if weak else int(1.5 * cfg["max_dets"] / world) + 65536
    if synthetic:
        gcfg["skip_vector"] = True
    wl = prepare_workload(gcfg, ctx)  # every rank prepares the same global start vector (deterministic)
    sm, mol = wl["sm"], wl["mol"]
    if not synthetic:
        _, owner = hash_owner(ctx, wl["keys"], wl["proc_scr"], world)
    spawn_cap_local = 4 * gcfg["mat_nonz"] // world          # spawn_length = matr_samp * 4 / n_procs
    seg_cap = 2 * gcfg["mat_nonz"] // (world * world) + 8192
    eng = MultiGpuFrisys(ctx, dist, rank, world, mol, sm, max_dets_local, spawn_cap_local, seg_cap, wl["proc_scr"],
                         wl["vec_scr"], wl["hf_en"], (wl["hf"], np.ones(1)), (wl["htrial_keys"], wl["htrial_vals"]),
                         route=os.environ.get("FRIES_ROUTE", "p2p"))
    if synthetic:
        # every rank draws its own share of random determinants, the shares are routed to their hash owners with one
        # all-to-all (untimed set-up), duplicates across ranks merge at the owner
        from bench import synthetic_vector
        k_r, v_r = synthetic_vector(sm, cfg["vec_nonz"], 1.0, seed=12345 + rank)
        if rank:
            k_r, v_r = k_r[1:], v_r[1:]  # the Hartree-Fock determinant comes from rank 0 only
        _, own_r = hash_owner(ctx, k_r, wl["proc_scr"], world)
        order = np.argsort(own_r, kind="stable")
        counts = np.bincount(own_r, minlength=world).astype(np.int64)
        dev = torch.device("cuda", torch.cuda.current_device())
        send_c = torch.from_numpy(counts).to(dev)
        recv_c = torch.empty_like(send_c)
        dist.all_to_all_single(recv_c, send_c)
        n_recv = int(recv_c.sum().item())
        send_k = torch.from_numpy(k_r[order].view(np.int64)).to(dev)
        send_v = torch.from_numpy(v_r[order]).to(dev)
        recv_k = torch.empty(n_recv, dtype=torch.int64, device=dev)
        recv_v = torch.empty(n_recv, dtype=torch.float64, device=dev)
        dist.all_to_all_single(recv_k, send_k, recv_c.tolist(), counts.tolist())
        dist.all_to_all_single(recv_v, send_v, recv_c.tolist(), counts.tolist())
        lk, first = np.unique(recv_k.cpu().numpy().view(np.uint64), return_index=True)
        lv = recv_v.cpu().numpy()[first]
        perm = np.random.default_rng(7 + rank).permutation(lk.size)
        lk, lv = np.ascontiguousarray(lk[perm]), np.ascontiguousarray(lv[perm])
        nrm = torch.tensor([np.abs(lv).sum()], dtype=torch.float64, device=dev)
        dist.all_reduce(nrm)
        lv *= gcfg["target"] / float(nrm.item())
        del send_k, send_v, recv_k, recv_v
        eng.vec.upload(lk, np.stack([lv, np.zeros_like(lv)]))
    else:
        eng.load(wl["keys"], wl["vals"], owner)
    params = FrisysParams(eps=cfg["eps"], init_thresh=cfg["initiator"], p_doub=wl["p_doub"],
                          new_hb=1 if cfg["dist"] == "HB_unnorm" else 0, matr_samp=gcfg["mat_nonz"],
                          target_nonz=gcfg["vec_nonz"], en_shift=0.0)
    rs = np.random.RandomState(1)
    uni = (rs.randint(0, 2**32, 6 * (args.warmup + 2 * args.steps + 16), dtype=np.uint64) / (1.0 + 0xFFFFFFFF)).reshape(-1, 6)
    ui = 0
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    last = None
    for _ in range(args.warmup):
        last = eng.iterate(params, uni[ui]); ui += 1
    clocks = ClockSampler(local)
    clocks.start()
    dist.barrier()
    torch.cuda.synchronize()
    launches0 = ctx.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    spawned = 0
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record(stream)
        last = eng.iterate(params, uni[ui]); ui += 1
        ev[k][1].record(stream)
        spawned += last.n_spawned
    dist.barrier()
    torch.cuda.synchronize()
    clk = clocks.stop()
    ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)  # max over ranks
    ms = float(ms.item())
    ms_per_step = ms / args.steps
    # e2e, the same definition as at N = 1: every step uploads this rank's shard of the vector from pinned host memory,
    # iterates, and downloads the shard again; timed on the host between barriers, max over ranks
    import time
    cap_l = eng.vec.capacity
    hk = torch.empty(cap_l, dtype=torch.int64).pin_memory().numpy().view(np.uint64)
    hv = torch.empty(2 * cap_l, dtype=torch.float64).pin_memory().numpy()
    n_now = eng.vec.download_into(hk, hv)
    h2d = d2h = 0
    dist.barrier()
    torch.cuda.synchronize()
    t_e2e = 0.0
    for k in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        check(lib.fries_vec_upload(eng.vec.h, hk.ctypes.data, hv.ctypes.data, n_now))
        h2d += n_now * 8 * 3 + 48
        last = eng.iterate(params, uni[ui]); ui += 1
        n_now = eng.vec.download_into(hk, hv)
        d2h += n_now * 8 * 3 + 128
        torch.cuda.synchronize()
        t_e2e += time.perf_counter() - t0
    te = torch.tensor([t_e2e, float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
    tmax = te.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(te)  # bytes: summed over the ranks
    e2e_value = args.steps / float(tmax[0].item())
    h2d_tot, d2h_tot = float(te[1].item()), float(te[2].item())
    # per-kernel times (event pairs around every launch; kernels with in-kernel exchanges include the wait for peers)
    ctx.set_profile(2)
    for _ in range(5):
        flush.zero_()
        last = eng.iterate(params, uni[ui]); ui += 1
    ctx.set_profile(0)
    kern = {}
    for nm in ["hbpp_stage0", "hbpp_stage1", "hbpp_stage2", "hbpp_stage3", "hbpp_stage4", "hbpp_finalize", "merge_insert",
               "merge_accum", "vec_phase", "death_axpy", "find_preserve", "sys_comp", "compact"]:
        t, n = ctx.kernel_ms(nm)
        if n:
            kern[nm] = round(t / n, 4)
    # roofline of the slowest kernel on rank 0 (an HB-PP stage: 40 B per matrix sample; one GPU's share)
    roofline = None
    if kern:
        top = max(kern, key=kern.get)
        try:
            peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                     "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
            src = "MEASURED_PEAKS.json (burst copy)"
        except OSError:
            peak, src = 6650.0, "fallback B200_PROFILING.md"
        share = 1 if weak else world  # one GPU's share of the configuration's samples / elements
        bytes_alg = {"hbpp_finalize": 56, "merge_insert": 28 if "merge_accum" in kern else 56, "merge_accum": 28}.get(top, 40) * cfg["mat_nonz"] // share
        if top in ("death_axpy", "find_preserve", "sys_comp", "compact"):
            bytes_alg = {"death_axpy": 32, "find_preserve": 24, "sys_comp": 16, "compact": 8}[top] * cfg["vec_nonz"] // share
        ach = bytes_alg / (kern[top] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": top, "achieved": round(ach, 2), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 5), "traffic": None, "peak_source": src, "ms_per_launch": kern[top],
                    "algorithmic_bytes_per_launch": bytes_alg, "what": "rank 0, one GPU's share of the samples"}
    states = eng.vec.states()
    rts = np.zeros(16)
    check(lib.fries_hbpp_round_stamps(eng.vec.hb, 3, ptr(rts)))
    err = eng.comm.error_epoch()
    if rank == 0:
        # spawn route over NVLink: 16 B (determinant | flag, value) per spawned element that leaves its GPU, (N - 1) / N of
        # them under a uniform owner hash; per GPU and direction, against the measured peer-copy rate (B200_PROFILING.md)
        nvl_bytes_per_gpu = 16.0 * (spawned / args.steps) * (world - 1) / world / world
        nvl_gbps = nvl_bytes_per_gpu / (ms_per_step * 1e-3) / 1e9
        out = {
            "metric": "fri_iterations_per_sec", "value": round(1000.0 / ms_per_step, 3), "unit": "iter/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": cfg["workload"] + (f" x{world} (weak scaling: vec_nonz, mat_nonz, target x n_gpus)" if weak else ""),
                       "vec_nonz": gcfg["vec_nonz"], "mat_nonz": gcfg["mat_nonz"],
                       "l2": "flushed between iterations (512 MB write)", "stored_dets": int(last.curr_size),
                       "engine": "compress2 stage kernels; round-1 vector kernels (multi-rank)"},
            "spawned_elements_per_sec": round(spawned / (ms * 1e-3), 1),
            "determinant_updates_per_sec": round(gcfg["vec_nonz"] * args.steps / (ms * 1e-3), 1),
            "gpu_launches": int(ctx.launch_count - launches0), "clocks": clk,
            "route": dict(eng.route_description,
                          scalar_reductions="in-kernel, peer-mapped inboxes over NVLink (csrc/comm.cuh)",
                          nvlink_bytes_per_gpu_per_step=round(nvl_bytes_per_gpu, 1), nvlink_GBps_per_gpu=round(nvl_gbps, 3),
                          nvlink_peak_GBps=770.0, nvlink_frac=round(nvl_gbps / 770.0, 5),
                          nvlink_note="averaged over the whole step; the stores are issued by the finalize kernel only"),
            "e2e": {"value": round(e2e_value, 3), "unit": "iter/s", "h2d_bytes_per_step": int(h2d_tot / args.steps),
                    "d2h_bytes_per_step": int(d2h_tot / args.steps),
                    "what": "per step and rank: fries_vec_upload of the rank's shard (pinned host memory) + "
                            "fries_frisys_mol_spawn + route + fries_frisys_mol_finish + fries_vec_download; bytes summed "
                            "over the ranks, time = max over ranks"},
            "roofline": roofline,
            "comm_error_epoch": err,
            "stage3_round_stamps_us": [round(x / 1e3, 1) for x in rts],
            "kernels_ms_rank0": kern,
            "rounds": {"hbpp_stages": [int(x) for x in states[:5, 4]], "find_preserve": int(states[6, 4]),
                       "fast": [int(x) for x in states[:7, 10]]},
            "stage_phase_us": {"what": "in-kernel %globaltimer of CTA 0 on rank 0: prep, preserved set, line scan, count, emit",
                               "us": [[round((states[s, 13 + k] - states[s, 12 + k]) / 1e3, 1) for k in range(5)]
                                      for s in range(5)],
                               "candidate_rounds_us": [round((states[s, 18] - states[s, 13]) / 1e3, 1) for s in range(5)],
                               "apply_cut_us": [round((states[s, 19] - states[s, 18]) / 1e3, 1) for s in range(5)]},
            "energy_est": last.numer / last.denom if last.denom else None,
        }
        print(json.dumps(out), flush=True)
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


def run_multi_gpu_frifull(args, cfg, ctx, dist, rank, world, local, prepare_workload, ClockSampler):
    """bench.py --config n2full --gpus N > 1 (BASELINE.json configs[3]: frifull_mol on a partitioned vector).  Weak scaling:
    vec_nonz parents PER GPU, i.e. N times the single-GPU line's compression budget; every connection is routed to its owner
    over NVLink from inside hv_fill_kernel.  value = spawned H.v elements per second over all GPUs."""
    import json
    import time

    from ._capi import FrifullParams
    from .api import hash_owner

    stream = torch.cuda.current_stream()
    gcfg = dict(cfg, mat_nonz=1, max_dets=4_000_000, vec_nonz=cfg["vec_nonz"] * world, target=cfg["target"] * world)
    wl = prepare_workload(gcfg, ctx)  # the same global start vector on every rank (deterministic)
    sm, mol = wl["sm"], wl["mol"]
    _, owner = hash_owner(ctx, wl["keys"], wl["proc_scr"], world)
    seg_cap = 1 << 22  # connections per routed window and rank
    eng = MultiGpuFrisys(ctx, dist, rank, world, mol, sm, cfg["max_dets"], 1 << 16, seg_cap, wl["proc_scr"], wl["vec_scr"],
                         wl["hf_en"], (wl["hf"], np.ones(1)), (wl["htrial_keys"], wl["htrial_vals"]), route="p2p")
    eng.load(wl["keys"], wl["vals"], owner)
    fp = FrifullParams(eps=cfg["eps"], target_nonz=gcfg["vec_nonz"], en_shift=0.0, adjust_shift=0, damp_factor=0.05,
                       target_norm=0.0, last_one_norm=0.0)
    rs = np.random.RandomState(1)
    uni = rs.randint(0, 2**32, args.warmup + 2 * args.steps + 16, dtype=np.uint64) / (1.0 + 0xFFFFFFFF)
    ui = 0
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    last = None
    for _ in range(args.warmup):
        last = eng.frifull_iterate(fp, float(uni[ui])); ui += 1
    clocks = ClockSampler(local)
    clocks.start()
    dist.barrier()
    torch.cuda.synchronize()
    launches0 = ctx.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    spawned = 0
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record(stream)
        last = eng.frifull_iterate(fp, float(uni[ui])); ui += 1
        ev[k][1].record(stream)
        spawned += last.n_spawned  # global (summed over the ranks inside the call)
    dist.barrier()
    torch.cuda.synchronize()
    clk = clocks.stop()
    ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    ms_per_step = ms / args.steps
    launches = ctx.launch_count - launches0
    # e2e: every step uploads this rank's shard from pinned host memory, iterates, downloads the shard
    cap_l = eng.vec.capacity
    hk = torch.empty(cap_l, dtype=torch.int64).pin_memory().numpy().view(np.uint64)
    hv = torch.empty(2 * cap_l, dtype=torch.float64).pin_memory().numpy()
    n_now = eng.vec.download_into(hk, hv)
    n_e2e = max(2, args.steps // 4)
    t_e2e, h2d, d2h, sp_e2e = 0.0, 0, 0, 0
    for k in range(n_e2e):
        flush.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        check(lib.fries_vec_upload(eng.vec.h, hk.ctypes.data, hv.ctypes.data, n_now))
        h2d += n_now * 24 + 48
        st = eng.frifull_iterate(fp, float(uni[ui])); ui += 1
        n_now = eng.vec.download_into(hk, hv)
        d2h += n_now * 24 + 128
        torch.cuda.synchronize()
        t_e2e += time.perf_counter() - t0
        sp_e2e += st.n_spawned
    te = torch.tensor([t_e2e, float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
    tmax = te.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(te)
    ctx.set_profile(2)
    for _ in range(2):
        flush.zero_()
        eng.frifull_iterate(fp, float(uni[ui])); ui += 1
    ctx.set_profile(0)
    kern = {}
    for nm in ["h_diag", "hv_count", "hv_scan", "hv_fill", "merge_insert", "merge_accum", "find_preserve", "sys_comp", "compact"]:
        t, n = ctx.kernel_ms(nm)
        if n:
            kern[nm] = round(t / 2, 4)
    err = eng.comm.error_epoch()
    if rank == 0:
        try:
            peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                     "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
        except OSError:
            peak = 6650.0
        per_iter = spawned / args.steps
        top = max(kern, key=kern.get) if kern else None
        bytes_top = {"hv_fill": 16, "merge_insert": 28 if "merge_accum" in kern else 56, "merge_accum": 28}.get(top, 56) * per_iter / world
        nvl_bytes = 16.0 * per_iter * (world - 1) / world / world  # per GPU and direction
        nvl_gbps = nvl_bytes / (ms_per_step * 1e-3) / 1e9
        out = {
            "metric": "spawned_hv_elements_per_sec", "value": round(spawned / (ms * 1e-3), 1), "unit": "elements/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["workload"] + f" x{world} (weak scaling: vec_nonz per GPU)", "vec_nonz": gcfg["vec_nonz"],
                       "l2": "flushed between iterations (512 MB write)", "stored_dets": int(last.curr_size),
                       "spawned_per_iteration": int(per_iter)},
            "fri_iterations_per_sec": round(1000.0 / ms_per_step, 3),
            "gpu_launches": int(launches), "clocks": clk,
            "e2e": {"value": round(sp_e2e / float(tmax[0].item()), 1), "unit": "elements/s",
                    "h2d_bytes_per_step": int(float(te[1].item()) / n_e2e), "d2h_bytes_per_step": int(float(te[2].item()) / n_e2e),
                    "what": "per step and rank: fries_vec_upload of the rank's shard (pinned host memory) + "
                            "fries_frifull_mol_iterate + fries_vec_download; bytes summed over the ranks, time = max over ranks"},
            "roofline": {"bound": "hbm", "kernel": top, "achieved": round(bytes_top / (kern[top] * 1e-3) / 1e9, 2) if top else None,
                         "peak": peak, "unit": "GB/s", "frac": round(bytes_top / (kern[top] * 1e-3) / 1e9 / peak, 5) if top else None,
                         "traffic": None, "algorithmic_bytes_per_launch": int(bytes_top), "kernels_ms_per_iteration_rank0": kern,
                         "what": "rank 0, one GPU's share; a kernel's time includes the waits of its in-kernel exchanges"},
            "route": {"collective": "none: hv_fill_kernel stores every connection into its owner's peer-mapped window "
                                    "(csrc/comm.cuh RouteView), windows of 2^22 connections per rank",
                      "nvlink_bytes_per_gpu_per_step": round(nvl_bytes, 1), "nvlink_GBps_per_gpu": round(nvl_gbps, 3),
                      "nvlink_peak_GBps": 770.0, "nvlink_frac": round(nvl_gbps / 770.0, 5)},
            "comm_error_epoch": err,
            "energy_est": last.numer / last.denom if last and last.denom else None,
            "cpu_baseline": {"value": None, "unit": "elements/s", "cores": 1, "kind": "reference",
                             "sample": "not run in this line (see the N = 1 line and bench.py --impl reference)"},
        }
        print(json.dumps(out), flush=True)
    eng.close()

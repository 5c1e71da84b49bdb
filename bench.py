#!/usr/bin/env python
"""bench.py -- FRI iterations/s and spawned H.v elements/s of the frisys_mol hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config ne|h2o]

A "step" is one frisys_mol iteration (FRIES_bin/frisys_mol.cpp:405-552: HB-PP matrix compression ->
spawn -> (route) -> hash-merge -> death/cloning -> find_preserve -> projected energy -> sys_comp -> delete)
on the Ne aug-cc-pVDZ-sized configuration of examples/run_neon.sh (BASELINE.json configs[1]; synthetic
integrals, see fries_b200/synth.py).  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # examples/run_neon.sh:12 sizes; FCIDUMP-style input (frozen core dropped), D2h
    "ne": dict(system="ne", seed=2, point_group="D2h", eps=0.001, target=250000.0, vec_nonz=242000, mat_nonz=260000,
               max_dets=500000, initiator=1.0, dist="HB_unnorm",
               workload="Ne aug-cc-pVDZ-sized frisys_mol (NORB=22 NELEC=8 D2h, HB_unnorm, vec_nonz 242000, mat_nonz 260000)"),
    # Benchmarks/Results.tex:43 sizes, H2O cc-pVDZ, C2v
    "h2o": dict(system="h2o", seed=3, point_group="C2v", eps=0.001, target=1000000.0, vec_nonz=1000000,
                mat_nonz=1000000, max_dets=2000000, initiator=1.0, dist="HB_unnorm",
                workload="H2O cc-pVDZ-sized frisys_mol (NORB=24 NELEC=10 C2v, HB_unnorm, vec_nonz 1e6, mat_nonz 1e6)"),
    # BASELINE.json configs[4] / SURVEY.md 8d C5, the share of ONE GPU: 1.25e7 distinct random determinants (5 alpha +
    # 5 beta in 26 orbitals, default_rng(12345)), values sign * 10^(-4u), N2 cc-pVDZ-sized integrals (frozen core)
    "c5": dict(system="n2", seed=7, point_group="D2h", eps=0.001, target=1.25e7, vec_nonz=12500000, mat_nonz=12500000,
               max_dets=25000000, initiator=1.0, dist="HB_unnorm", synthetic_vector=True, frozen=True,
               workload="synthetic 1.25e7-determinant vector per GPU, N2 cc-pVDZ-sized integrals (NORB=26 NELEC=10, "
                        "HB_unnorm, vec_nonz = mat_nonz = 1.25e7 per GPU): spawn -> merge -> compress"),
    # BASELINE.json configs[3]: N2 stretched cc-pVDZ-sized frifull_mol -- full deterministic H.v + systematic vector
    # compression (FRIES_bin/frifull_mol.cpp:256-320).  One GPU's share of the 8-GPU configuration: 2.5e4 parents per
    # iteration, ~2.1e3 connections each, i.e. ~5e7 spawned H.v elements per iteration.
    "n2full": dict(system="n2", seed=7, point_group="D2h", eps=0.001, target=2.5e4, vec_nonz=25000, mat_nonz=0,
                   max_dets=40000000, initiator=0.0, dist="full", frozen=True, full_hv=True,
                   workload="N2 stretched cc-pVDZ-sized frifull_mol (NORB=26 NELEC=10 frozen core, full deterministic H.v + "
                            "systematic compression to vec_nonz 25000)"),
}

# SURVEY.md section 8d algorithmic bytes
B_PER_SAMPLE_STAGE = 40   # per matrix sample per HB-PP stage
B_PER_SPAWN = 56          # per spawned element (write, read at merge, probe, value RMW)
START_PARENTS = (1, 1000, 1500)  # parents kept before each H application (x vec_nonz / 242000 for the last)


def start_parents(cfg):
    return START_PARENTS[:2] + (max(START_PARENTS[2], START_PARENTS[2] * cfg["vec_nonz"] // 242000),)
# per stored vector element; vec_phase = the four fused (death/cloning, find_preserve, sys_comp, delete + compact)
B_PER_VEC_EL = {"death_axpy": 32, "find_preserve": 24, "sys_comp": 16, "compact": 8, "vec_phase": 80}
# roofline.traffic (dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel) cannot be measured inside this
# program: captured_traffic() reads it from the committed ncu --set full capture of this command under profiles/ when there is
# one for the dominant kernel at this configuration, and the line says which file; otherwise null


def captured_traffic(config_name, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of `kernel` from the committed `ncu --set full` capture of this command
    (profiles/r02c_ncu_full_<kernel>_<config>_raw.csv, exported with `ncu -i ... --page raw --csv`) -> (bytes, file) or
    (None, reason).  Read from the capture, never typed in: a kernel without a capture of its own reports null."""
    import csv
    tag = {"hbpp_stage4": "stage4", "vec_phase": "vecphase"}.get(kernel)
    path = os.path.join(ROOT, "profiles", f"r02c_ncu_full_{tag}_{config_name}_raw.csv") if tag else None
    if not path or not os.path.exists(path):
        return None, "no ncu --set full capture of this kernel at this configuration under profiles/ (captures: r02c_ncu_full_*)"
    rows = list(csv.reader(open(path)))
    h, u, v = rows[0], rows[1], rows[2]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = h.index(key)
        tot += float(v[i]) * scale[u[i]]
    return int(tot), os.path.relpath(path, ROOT)


def mt_uniforms(seed, n):
    """std::mt19937(seed)() / (1. + UINT32_MAX), the reference's uniform (compress_utils.cpp:24)"""
    rs = np.random.RandomState(seed)
    return rs.randint(0, 2**32, n, dtype=np.uint64) / (1.0 + 0xFFFFFFFF)


def mt_u32(rs, n):
    return rs.randint(0, 2**32, n, dtype=np.uint64).astype(np.uint32)


# ---------------------------------------------------------------------------------------------------
# workload preparation (untimed)
# ---------------------------------------------------------------------------------------------------
def prepare_workload(cfg, ctx, n_ranks=1, rank=0):
    import fries_b200
    from fries_b200.synth import SynthMol

    sm = SynthMol(cfg["system"], cfg["seed"], frozen=bool(cfg.get("frozen", False)))
    mol = fries_b200.Mol.from_synth(ctx, sm)
    rs = np.random.RandomState(0)  # mt19937(0): proc scrambler then vec scrambler (frisys_mol.cpp:132-144)
    proc_scr, vec_scr = mt_u32(rs, sm.n_bits), mt_u32(rs, sm.n_bits)
    hf = np.array([sm.hf], np.uint64)
    hf_en = float(mol.diag(hf)[0])
    # trial vector = HF, H*trial with the diagonal shifted by hf_en (frisys_mol.cpp:155-214)
    tmp = fries_b200.Vec(ctx, 1 << 16, sm.n_bits, sm.n_elec, 2, proc_scr, vec_scr)
    tmp.set_diag_mol(mol, hf_en)
    tmp.add(hf, np.ones(1), np.ones(1, np.uint8))
    tmp.h_apply(mol, 0, 1, 0.0, 1.0)
    hk, hv = tmp.download()
    htrial_keys, htrial_vals = hk.copy(), hv[1].copy()
    n_sing = int(mol.sing_ex(hf)[0][-1])
    n_doub = int(mol.doub_ex(hf)[0][-1])
    p_doub = n_doub / (n_sing + n_doub)  # frisys_mol.cpp:217-220
    tmp.close()
    if cfg.get("synthetic_vector"):
        if cfg.get("skip_vector"):  # multi-GPU: every rank draws and routes its own share (fries_b200/multi.py)
            keys, vals = hf, np.ones(1)
        else:
            keys, vals = synthetic_vector(sm, cfg["vec_nonz"], cfg["target"])
        return dict(sm=sm, mol=mol, proc_scr=proc_scr, vec_scr=vec_scr, hf=hf, hf_en=hf_en, htrial_keys=htrial_keys,
                    htrial_vals=htrial_vals, p_doub=p_doub, keys=keys, vals=vals)
    # starting vector: (1 - 0.5 (H - E_HF))^3 HF restricted to the PARENTS[k] largest elements before each
    # application -- a realistic population (HF, singles/doubles, up to hextuples), truncated to the vec_nonz
    # largest elements and scaled to the target one-norm
    big = fries_b200.Vec(ctx, 8 * cfg["max_dets"], sm.n_bits, sm.n_elec, 2, proc_scr, vec_scr)
    big.set_diag_mol(mol, hf_en)
    k2, cur = hf, np.ones(1)
    for npar in start_parents(cfg):
        order = np.argsort(-np.abs(cur), kind="stable")[:npar]
        big.upload(k2[order], np.stack([cur[order], np.zeros(order.size)]))
        big.h_apply(mol, 0, 1, 1.0, -0.5)
        k2, v2 = big.download()
        cur = v2[1]
    big.close()
    vals = cur
    order = np.argsort(-np.abs(vals), kind="stable")[: cfg["vec_nonz"]]
    keys, vals = k2[order], vals[order]
    vals = vals * (cfg["target"] / np.abs(vals).sum())
    perm = np.random.default_rng(7).permutation(keys.size)  # storage order is not sorted by magnitude in a real run
    keys, vals = np.ascontiguousarray(keys[perm]), np.ascontiguousarray(vals[perm])
    return dict(sm=sm, mol=mol, proc_scr=proc_scr, vec_scr=vec_scr, hf=hf, hf_en=hf_en, htrial_keys=htrial_keys,
                htrial_vals=htrial_vals, p_doub=p_doub, keys=keys, vals=vals)


def synthetic_vector(sm, n, one_norm, seed=12345):
    """n distinct random determinants (HF first) with values sign * 10^(-4u), u ~ U(0,1), scaled to the one-norm"""
    rng = np.random.default_rng(seed)
    M, h = sm.n_orb, sm.n_elec // 2
    parts, have = [np.array([sm.hf], np.uint64)], 1
    while have < n:
        m = min(2_000_000, int(1.1 * (n - have)) + 1024)
        a = np.argpartition(rng.random((m, M), dtype=np.float32), h, axis=1)[:, :h].astype(np.uint64)
        b = np.argpartition(rng.random((m, M), dtype=np.float32), h, axis=1)[:, :h].astype(np.uint64)
        k = (np.uint64(1) << a).sum(axis=1, dtype=np.uint64) | ((np.uint64(1) << b).sum(axis=1, dtype=np.uint64) << np.uint64(M))
        parts.append(k)
        keys = np.unique(np.concatenate(parts))
        parts, have = [keys], keys.size
    keys = parts[0]
    keys = keys[keys != np.uint64(sm.hf)]
    keys = np.concatenate([np.array([sm.hf], np.uint64), keys[rng.permutation(keys.size)[: n - 1]]])
    vals = rng.choice([-1.0, 1.0], n) * 10.0 ** (-4.0 * rng.random(n))
    vals[0] = 1.0
    vals *= one_norm / np.abs(vals).sum()
    return np.ascontiguousarray(keys), np.ascontiguousarray(vals)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_frifull(args, cfg, ctx, stream, local):
    """--config n2full: frifull_mol iterations (compress the iterate to vec_nonz, then the full deterministic H.v into the
    other row).  value = spawned H.v elements per second; the H.v kernels gather packed integrals from L2 (0.5 MB table),
    so the roofline line is the spawned-element byte contract (56 B each) against HBM."""
    import torch

    import fries_b200
    from fries_b200._capi import FrifullParams
    wcfg = dict(cfg, mat_nonz=1, max_dets=4_000_000)  # the start vector's construction needs a far smaller store
    wl = prepare_workload(wcfg, ctx)
    sm, mol = wl["sm"], wl["mol"]
    vec = fries_b200.Vec(ctx, cfg["max_dets"], sm.n_bits, sm.n_elec, 2, wl["proc_scr"], wl["vec_scr"])
    vec.set_diag_mol(mol, wl["hf_en"])
    vec.upload(wl["keys"], np.stack([wl["vals"], np.zeros_like(wl["vals"])]))
    vec.frisys_setup(mol, 1 << 24, wl["hf"], np.ones(1), wl["htrial_keys"], wl["htrial_vals"])
    fp = FrifullParams(eps=cfg["eps"], target_nonz=cfg["vec_nonz"], en_shift=0.0, adjust_shift=0, damp_factor=0.05,
                       target_norm=0.0, last_one_norm=0.0)
    uni = mt_uniforms(1, args.warmup + 2 * args.steps + 16)
    ui = 0
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    last = None
    for _ in range(args.warmup):
        last = vec.frifull_iterate(fp, float(uni[ui])); ui += 1
    clocks = ClockSampler(local)
    clocks.start()
    torch.cuda.synchronize()
    launches0 = ctx.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    spawned = 0
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record(stream)
        last = vec.frifull_iterate(fp, float(uni[ui])); ui += 1
        ev[k][1].record(stream)
        spawned += last.n_spawned
    torch.cuda.synchronize()
    clk = clocks.stop()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    ms_per_step = ms / args.steps
    launches = ctx.launch_count - launches0
    # e2e: the compressed iterate goes to the host and comes back every step (upload -> iterate -> download of the result)
    hk = torch.empty(cfg["max_dets"], dtype=torch.int64).pin_memory().numpy().view(np.uint64)
    hv = torch.empty(2 * cfg["max_dets"], dtype=torch.float64).pin_memory().numpy()
    n_now = vec.download_into(hk, hv)
    t_e2e, h2d, d2h = 0.0, 0, 0
    from fries_b200._capi import check, lib
    for k in range(max(2, args.steps // 4)):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        check(lib.fries_vec_upload(vec.h, hk.ctypes.data, hv.ctypes.data, n_now))
        h2d += n_now * 24 + 48
        st = vec.frifull_iterate(fp, float(uni[ui])); ui += 1
        n_now = vec.download_into(hk, hv)
        d2h += n_now * 24 + 128
        torch.cuda.synchronize()
        t_e2e += time.perf_counter() - t0
    n_e2e = max(2, args.steps // 4)
    ctx.set_profile(2)
    for _ in range(3):
        flush.zero_()
        vec.frifull_iterate(fp, float(uni[ui])); ui += 1
    ctx.set_profile(0)
    kern = {}
    for nm in ["h_diag", "hv_count", "hv_scan", "hv_fill", "merge_insert", "merge_accum", "find_preserve", "sys_comp", "compact"]:
        t, n = ctx.kernel_ms(nm)
        if n:
            kern[nm] = round(t / 3, 4)  # per iteration (a kernel may run several windows)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    per_iter = spawned / args.steps
    top = max(kern, key=kern.get) if kern else None
    bytes_top = {"hv_fill": 16, "merge_insert": 28 if "merge_accum" in kern else 56, "merge_accum": 28}.get(top, 56) * per_iter
    out = {
        "metric": "spawned_hv_elements_per_sec", "value": round(spawned / (ms * 1e-3), 1), "unit": "elements/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["workload"], "vec_nonz": cfg["vec_nonz"], "l2": "flushed between iterations (512 MB write)",
                   "stored_dets": int(last.curr_size), "spawned_per_iteration": int(per_iter)},
        "fri_iterations_per_sec": round(1000.0 / ms_per_step, 3),
        "gpu_launches": int(launches), "clocks": clk,
        "e2e": {"value": round(per_iter * n_e2e / t_e2e, 1), "unit": "elements/s", "h2d_bytes_per_step": int(h2d / n_e2e),
                "d2h_bytes_per_step": int(d2h / n_e2e),
                "what": "fries_vec_upload (pinned host vector) + fries_frifull_mol_iterate + fries_vec_download per step"},
        "roofline": {"bound": "hbm", "kernel": top, "achieved": round(bytes_top / (kern[top] * 1e-3) / 1e9, 2) if top else None,
                     "peak": peak, "unit": "GB/s", "frac": round(bytes_top / (kern[top] * 1e-3) / 1e9 / peak, 5) if top else None,
                     "traffic": None, "algorithmic_bytes_per_launch": int(bytes_top), "kernels_ms_per_iteration": kern,
                     "iter_algorithmic_GBps": round((96 * cfg["vec_nonz"] + 56 * per_iter) / (ms_per_step * 1e-3) / 1e9, 2),
                     "note": "B_iter(frifull) = 96 N_v + 56 N_s (SURVEY 8d); the connection arithmetic gathers integrals from L2"},
        "energy_est": last.numer / last.denom if last and last.denom else None,
        "cpu_baseline": {"value": None, "unit": "elements/s", "cores": 1, "kind": "reference",
                         "sample": "not run in this line: one reference frifull_mol iteration at this size takes ~40 s per core "
                                   "(SURVEY 6: 1.3e6 spawned elements/s/core); see profiles/ for the recorded comparison"},
    }
    out["roofline"]["iter_frac"] = round(out["roofline"]["iter_algorithmic_GBps"] / peak, 5)
    print(json.dumps(out), flush=True)
    vec.close()
    mol.close()
    ctx.close()


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist

    import fries_b200
    from fries_b200._capi import FrisysParams, check, lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = fries_b200.Context(local)
    # one explicit (non-default) stream for the library's kernels, the L2 flush and the timing events
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    check(lib.fries_ctx_set_stream(ctx.h, stream.cuda_stream))
    if os.environ.get("FRIES_NO_BRACKET"):  # diagnostics: the reference's plain rounds in every compression
        check(lib.fries_debug_set_bracket(0))

    if world > 1 and cfg.get("full_hv"):
        from fries_b200.multi import run_multi_gpu_frifull
        return run_multi_gpu_frifull(args, cfg, ctx, dist, rank, world, local, prepare_workload, ClockSampler)
    if world > 1:
        from fries_b200.multi import run_multi_gpu_bench
        return run_multi_gpu_bench(args, cfg, ctx, dist, rank, world, local, prepare_workload, ClockSampler,
                                   cpu_baseline_block)

    if cfg.get("full_hv"):
        return run_frifull(args, cfg, ctx, stream, local)
    wl = prepare_workload(cfg, ctx)
    sm, mol = wl["sm"], wl["mol"]
    spawn_cap = 4 * cfg["mat_nonz"]  # spawn_length = matr_samp * 4 / n_procs (frisys_mol.cpp:109)
    vec = fries_b200.Vec(ctx, cfg["max_dets"], sm.n_bits, sm.n_elec, 2, wl["proc_scr"], wl["vec_scr"])
    vec.set_diag_mol(mol, wl["hf_en"])
    vec.upload(wl["keys"], np.stack([wl["vals"], np.zeros_like(wl["vals"])]))
    vec.frisys_setup(mol, spawn_cap, wl["hf"], np.ones(1), wl["htrial_keys"], wl["htrial_vals"])
    params = FrisysParams(eps=cfg["eps"], init_thresh=cfg["initiator"], p_doub=wl["p_doub"],
                          new_hb=1 if cfg["dist"] == "HB_unnorm" else 0, matr_samp=cfg["mat_nonz"],
                          target_nonz=cfg["vec_nonz"], en_shift=0.0)
    uni = mt_uniforms(1, 6 * (args.warmup + 2 * args.steps + 128)).reshape(-1, 6)
    ui = 0
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    if os.environ.get("FRIES_BENCH_NOFLUSH"):  # diagnostic only (how much of an iteration is cold-cache cost); says so in config
        class _NoFlush:
            def zero_(self):
                pass
        flush = _NoFlush()

    last = None
    for _ in range(args.warmup):
        last = vec.frisys_iterate(params, uni[ui]); ui += 1

    # ---- timed region: K iterations, state resident in HBM, L2 flushed between iterations ----
    clocks = ClockSampler(local)
    clocks.start()
    torch.cuda.synchronize()
    launches0 = ctx.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    spawned = samples = 0
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record(stream)
        last = vec.frisys_iterate(params, uni[ui]); ui += 1
        ev[k][1].record(stream)
        spawned += last.n_spawned
        samples += last.n_matrix_samples
    torch.cuda.synchronize()
    clk = clocks.stop()
    launches = ctx.launch_count - launches0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    ms_per_step = ms / args.steps
    value = 1000.0 / ms_per_step

    # ---- e2e: the same iteration with the vector in HOST (pinned) memory: upload -> iterate -> download ----
    n_now = vec.curr_size()
    hk = torch.empty(cfg["max_dets"], dtype=torch.int64).pin_memory().numpy().view(np.uint64)
    hv = torch.empty(2 * cfg["max_dets"], dtype=torch.float64).pin_memory().numpy()
    n_now = vec.download_into(hk, hv)
    h2d = d2h = 0
    t_e2e = 0.0
    for k in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        check(lib.fries_vec_upload(vec.h, hk.ctypes.data, hv.ctypes.data, n_now))
        h2d += n_now * 8 * 3 + 48
        st = vec.frisys_iterate(params, uni[ui]); ui += 1
        n_now = vec.download_into(hk, hv)
        d2h += n_now * 8 * 3 + 128
        torch.cuda.synchronize()
        t_e2e += time.perf_counter() - t0
    e2e_value = args.steps / t_e2e

    # ---- per-kernel times (separate pass with event pairs around every launch) ----
    ctx.set_profile(2)
    prof_iters = 5
    for _ in range(prof_iters):
        flush.zero_()
        stp = vec.frisys_iterate(params, uni[ui]); ui += 1
    ctx.set_profile(0)
    states = vec.states()
    # how often the threshold brackets of the previous iteration held (a miss falls back to the plain Newton rounds)
    hit_iters = 40
    hits = np.zeros(7)
    fp_rounds = []
    for _ in range(hit_iters):
        vec.frisys_iterate(params, uni[ui]); ui += 1
        st_ = vec.states()
        hits += st_[:7, 10]
        fp_rounds.append(int(st_[6, 4]))
    bracket_hits = {"iterations": hit_iters, "hbpp_stages": [int(x) for x in hits[:5]], "find_preserve": int(hits[6]),
                    "find_preserve_rounds_max": max(fp_rounds)}
    names = ["hbpp_stage0", "hbpp_stage1", "hbpp_stage2", "hbpp_stage3", "hbpp_stage4", "hbpp_finalize", "merge_insert",
             "merge_accum", "vec_phase", "death_axpy", "find_preserve", "sys_comp", "compact"]
    kern = {}
    for nm in names:
        t, n = ctx.kernel_ms(nm)
        if n:
            kern[nm] = t / n
    n_vec = stp.curr_size
    units = {nm: cfg["mat_nonz"] * B_PER_SAMPLE_STAGE for nm in names if nm.startswith("hbpp_stage")}
    units["hbpp_finalize"] = stp.n_matrix_samples * (B_PER_SAMPLE_STAGE + 16)
    units["merge_insert"] = stp.n_spawned * (28 if "merge_accum" in kern else 56)  # one pass since the end of round 2: both halves
    units["merge_accum"] = stp.n_spawned * 28
    for nm, b in B_PER_VEC_EL.items():
        units[nm] = n_vec * b
    top = max(kern, key=kern.get)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = units[top] / (kern[top] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": top, "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 5),
                "peak_nominal": 8000.0, "frac_nominal": round(achieved / 8000.0, 5),  # report against both
                "traffic": captured_traffic(args.config, top)[0],
                "traffic_source": captured_traffic(args.config, top)[1],
                "kernel_table": {k: {"ms": round(kern[k], 4), "algorithmic_MB": round(units[k] / 1e6, 2),
                                     "GBps": round(units[k] / (kern[k] * 1e-3) / 1e9, 1),
                                     "frac": round(units[k] / (kern[k] * 1e-3) / 1e9 / peak, 4)} for k in kern if k in units},
                "algorithmic_bytes_per_launch": units[top],
                "peak_source": "MEASURED_PEAKS.json (burst copy)" if peaks else "fallback B200_PROFILING.md",
                "ms_per_launch": round(kern[top], 4),
                "kernels_ms": {k: round(v, 4) for k, v in kern.items()},
                "rounds": {"hbpp_stages": [int(x) for x in states[:5, 4]], "find_preserve": int(states[6, 4])},
                "stage_items_in": [int(x) for x in states[:5, 7]], "stage_items_out": [int(x) for x in states[:5, 6]],
                "stage_bracket": {"fast": [int(x) for x in states[:5, 10]], "candidates": [int(x) for x in states[:5, 9]]},
                "find_preserve_bracket": {"fast": int(states[6, 10]), "candidates": int(states[6, 9]),
                                          "rounds": int(states[6, 4])},
                "bracket_hits": bracket_hits,
                "stage_phase_us": {"what": "in-kernel %globaltimer of CTA 0: prep, preserved set, line scan, count, emit",
                                   "us": [[round((states[s, 13 + k] - states[s, 12 + k]) / 1e3, 1) for k in range(5)]
                                          for s in range(5)],
                                   "candidate_rounds_us": [round((states[s, 18] - states[s, 13]) / 1e3, 1) for s in range(5)],
                                   "apply_cut_us": [round((states[s, 19] - states[s, 18]) / 1e3, 1) for s in range(5)]},
                "iter_algorithmic_GBps": round((96 * n_vec + 200 * cfg["mat_nonz"] + 56 * stp.n_spawned) / (ms_per_step * 1e-3) / 1e9, 2)}
    if os.environ.get("FRIES_TIMELINE"):
        names = ["start", "A loop", "A cta-sum", "A grid-sum", "solve", "fixup", "set decided", "B loads", "B cta-sum", "B grid-sum",
                 "grid seeded", "C loads", "C scan", "C tests", "C compact", "C rows", "C tile end", "C done", "C grid-sum",
                 "D loads", "D scan", "D coded", "D heavy", "D tile end", "end", "A posted", "A reduced", "A solved", "A published"]
        for s_ in range(5):
            tl = vec.timeline(s_)
            print(f"timeline stage {s_} (SM cycles, thread 0 of CTA 0): " +
                  ", ".join(f"{nm} {int(tl[k])}" for k, nm in enumerate(names)), file=sys.stderr)
    if os.environ.get("FRIES_CTA_MARKS"):
        for s_ in (0, 3, 4, 5, 6):
            try:
                m = vec.cta_marks(s_)
            except Exception as ex:  # the fused vector kernel is not part of every configuration
                print(f"cta marks {s_}: {ex}", file=sys.stderr)
                continue
            labels = (["start", "A done", "set decided", "B done", "C done", "end"] if s_ < 5 else
                      ["start", "P1 done", "set decided", "P2 done", "line seeded", "P3 done", "offsets known", "P4 done"] if s_ == 5
                      else ["P4 tile0 scanned", "P4 tile0 stored", "P4 tile0 indexed", "index cleared"])
            for k, nm in enumerate(labels):
                r = m[k]
                order = r.argsort()
                print(f"cta marks stage {s_} {nm}: min {r.min() / 1e3:.1f} us (cta {order[0]}), median {float(sorted(r)[len(r) // 2]) / 1e3:.1f}, "
                      f"p90 {float(sorted(r)[int(len(r) * 0.9)]) / 1e3:.1f}, max {r.max() / 1e3:.1f} (cta {order[-1]}, {order[-2]}, {order[-3]})", file=sys.stderr)
    # the whole iteration against the roofline, by SURVEY 8d's byte contract B_iter = 96 N_v + 200 N_m + 56 N_s
    roofline["iter_frac"] = round(roofline["iter_algorithmic_GBps"] / peak, 5)

    out = {
        "metric": "fri_iterations_per_sec", "value": round(value, 3), "unit": "iter/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["workload"], "vec_nonz": cfg["vec_nonz"], "mat_nonz": cfg["mat_nonz"],
                   "l2": ("NOT flushed (diagnostic run, FRIES_BENCH_NOFLUSH)" if os.environ.get("FRIES_BENCH_NOFLUSH") else "flushed between iterations (512 MB write)"), "stored_dets": int(n_vec),
                   "engine": "compress2 + vecphase" if not os.environ.get("FRIES_ENGINE") else "round-1 kernels (FRIES_ENGINE=1)"},
        "spawned_elements_per_sec": round(spawned / (ms * 1e-3), 1),
        "matrix_samples_per_sec": round(samples / (ms * 1e-3), 1),
        "gpu_launches": int(launches), "clocks": clk,
        "e2e": {"value": round(e2e_value, 3), "unit": "iter/s", "h2d_bytes_per_step": int(h2d / args.steps),
                "d2h_bytes_per_step": int(d2h / args.steps),
                "what": "fries_vec_upload (pinned host vector) + fries_frisys_mol_iterate + fries_vec_download per step"},
        "roofline": roofline,
        "energy_est": last.numer / last.denom if last and last.denom else None,
    }
    # CPU baseline: the reference driver started from the SAME warmed-up vector (downloaded from the GPU)
    wk, wv = vec.download()
    wl["keys"], wl["vals"] = wk, wv[0]
    if cfg.get("synthetic_vector"):
        out["cpu_baseline"] = {"value": None, "unit": "iter/s", "cores": 1, "kind": "reference",
                               "sample": "not run at this size (one reference iteration takes > 10 s); see the default config"}
    else:
        try:
            out["cpu_baseline"] = cpu_baseline_block(cfg, wl, n_iter=int(os.environ.get("FRIES_BENCH_CPU_ITERS", "12")))
        except Exception as exc:  # the GPU numbers above stand on their own: never lose the line to the CPU leg
            out["cpu_baseline"] = {"value": None, "unit": "iter/s", "cores": 1, "kind": "reference",
                                   "sample": f"reference run failed: {type(exc).__name__}: {exc}"[:300]}
    print(json.dumps(out), flush=True)
    vec.close()
    mol.close()
    ctx.close()


# ---------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own frisys_mol (oracle/_ref/frisys_mol, built from
# /root/reference by oracle/Makefile with a single-rank MPI stand-in), same integrals, same start vector.
# ---------------------------------------------------------------------------------------------------
def write_fcidump(path, sm, point_group):
    inv = {"D2h": {0: 1, 7: 2, 6: 3, 1: 4, 5: 5, 2: 6, 3: 7, 4: 8}, "C2v": {0: 1, 2: 2, 3: 3, 1: 4}}[point_group]
    T = sm.tot_orb
    with open(path, "w") as f:
        f.write(f"&FCI NORB={T},NELEC={sm.n_elec_total},MS2=0,\n")
        f.write("ORBSYM=" + ",".join(str(inv[int(s)]) for s in sm.symm_all) + ",\n")
        f.write("ISYM=1,\n&END\n")
        e = sm.eris_chem
        lines = []
        for i in range(T):
            for j in range(i + 1):
                for k in range(i + 1):
                    for l in range(k + 1):
                        if k * (k + 1) // 2 + l > i * (i + 1) // 2 + j:
                            continue
                        v = e[i, j, k, l]
                        if v != 0.0:
                            lines.append(f"{float(v)!r} {i + 1} {j + 1} {k + 1} {l + 1}\n")
        f.writelines(lines)
        for i in range(T):
            for j in range(i + 1):
                if sm.hcore[i, j] != 0.0:
                    f.write(f"{float(sm.hcore[i, j])!r} {i + 1} {j + 1} 0 0\n")
        f.write("0.0 0 0 0 0\n")


def reference_rank_count():
    """ranks for the reference's MPI code on this host: the largest power of two that fits the cores this process may use
    (capped at 64).  There is no MPI installation here; oracle/mpi_shim runs the reference's own collectives between
    processes over a shared-memory file (oracle/mpi_shim/mpi.h, shimrun.py)."""
    if os.environ.get("FRIES_REF_RANKS"):
        return max(1, int(os.environ["FRIES_REF_RANKS"]))
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    try:  # a container may be limited to fewer CPUs than it can see (cgroup v2 cpu.max = "<quota> <period>" or "max ...")
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            cores = max(1, min(cores, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    try:  # cgroup v1
        quota = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if quota > 0 and period > 0:
            cores = max(1, min(cores, quota // period))
    except (OSError, ValueError):
        pass
    n = 1
    while 2 * n <= min(cores, 64):
        n *= 2
    return n


def reference_iter_seconds(cfg, sm, keys, vals, it_a, it_b, workdir, n_ranks=1):
    exe = os.path.join(ROOT, "oracle", "_ref", "frisys_mol")
    if not os.path.exists(exe):
        return None, "oracle/_ref/frisys_mol not built"
    sys.path.insert(0, os.path.join(ROOT, "oracle", "mpi_shim"))
    import shimrun
    fd = os.path.join(workdir, "FCIDUMP")
    write_fcidump(fd, sm, cfg["point_group"])
    with open(os.path.join(workdir, "ini_dets"), "w") as f:
        f.write("\n".join(str(int(k)) for k in keys) + "\n")
    with open(os.path.join(workdir, "ini_vals"), "w") as f:
        f.write("\n".join(repr(float(v)) for v in vals) + "\n")
    # one run of it_b iterations; the ranks' stdout goes through a pseudo-terminal and the arrival of the per-iteration
    # lines ("<iteration>, en est: ...", frisys_mol.cpp:545) is time-stamped: loop time of the last it_b - it_a iterations
    import re
    rd = os.path.join(workdir, "res") + "/"
    os.makedirs(rd, exist_ok=True)
    cmd = [exe, "--fcidump_path", fd, "--distribution", cfg["dist"], "--vec_nonz", str(cfg["vec_nonz"]), "--mat_nonz",
           str(cfg["mat_nonz"]), "--max_dets", str(cfg["max_dets"]), "--epsilon", str(cfg["eps"]), "--target",
           str(cfg["target"]), "--initiator", str(cfg["initiator"]), "--max_iter", str(it_b), "--result_dir", rd,
           "--ini_vec", os.path.join(workdir, "ini_"), "--point_group", cfg["point_group"]]
    env = dict(os.environ, FRIES_SEED="1")
    err_path = os.path.join(workdir, "stderr.txt")
    def failure(rc):
        stderr = open(err_path, errors="replace").read()
        if rc != 0 or "Exception" in stderr:
            lines = [ln for ln in stderr.strip().splitlines() if ln.strip() and not ln.startswith("Please send")]
            return f"reference frisys_mol failed on {n_ranks} rank(s) (rc {rc}): " + (lines or ["?"])[-1][:200]
        return None

    try:
        with open(err_path, "w") as ef:
            rc, _ = shimrun.run(n_ranks, cmd, 256 << 20, timeout=300, env=env, stderr=ef,
                                stamp=re.compile(r"^(\d+), en est: "))
    except OSError:
        # no pseudo-terminal on this host: wall-clock difference of two runs with different --max_iter (noisier)
        walls = []
        for n_it in (it_a, it_b):
            cmd[cmd.index("--max_iter") + 1] = str(n_it)
            with open(err_path, "w") as ef:
                rc, sec = shimrun.run(n_ranks, cmd, 256 << 20, timeout=300, env=env, stdout=subprocess.DEVNULL, stderr=ef)
            if failure(rc):
                return None, failure(rc)
            walls.append(sec)
        return (walls[1] - walls[0]) / (it_b - it_a), None
    if failure(rc):
        return None, failure(rc)
    t = {int(k): v for k, v in shimrun.run.stamps}
    if it_a - 1 not in t or it_b - 1 not in t:
        return None, f"reference frisys_mol on {n_ranks} rank(s): iteration lines {it_a - 1} / {it_b - 1} not seen on stdout"
    return (t[it_b - 1] - t[it_a - 1]) / (it_b - it_a), None


def reference_iter_seconds_best(cfg, sm, keys, vals, it_a, it_b):
    try:
        return _reference_iter_seconds_best(cfg, sm, keys, vals, it_a, it_b)
    except Exception as exc:
        return None, 0, f"reference run failed: {type(exc).__name__}: {exc}"[:300]


def _reference_iter_seconds_best(cfg, sm, keys, vals, it_a, it_b):
    """the reference on all the host cores it can use (reference_rank_count ranks), falling back to one rank if the
    multi-process run fails; returns (seconds per iteration, ranks used, error)"""
    tried, best = [], None
    n_max = reference_rank_count()
    # all cores, and a quarter of them when there are many (ranks with a few thousand samples each spend their time in the
    # collectives); one rank only if the multi-process runs fail
    for n in dict.fromkeys([n_max] + ([n_max // 4] if n_max >= 16 else []) + [1]):
        if n == 1 and best is not None:
            break
        wd = tempfile.mkdtemp(prefix="fries_ref_")
        try:
            sec, err = reference_iter_seconds(cfg, sm, keys, vals, it_a, it_b, wd, n)
        finally:
            shutil.rmtree(wd, ignore_errors=True)
        if sec is not None and sec > 0:
            if best is None or sec < best[0]:
                best = (sec, n)
        else:
            tried.append(err or "non-positive time")
    if best is not None:
        return best[0], best[1], None
    return None, 0, "; ".join(tried)


def cpu_baseline_block(cfg, wl, n_iter):
    sec, ranks, err = reference_iter_seconds_best(cfg, wl["sm"], wl["keys"], wl["vals"], 2, 2 + n_iter)
    if sec is None:
        return {"value": None, "unit": "iter/s", "cores": 1, "kind": "reference", "sample": err}
    return {"value": round(1.0 / sec, 4), "unit": "iter/s", "cores": ranks, "kind": "reference",
            "sample": f"{n_iter} iterations of the reference's frisys_mol on {ranks} rank(s) = host cores (its own MPI code; "
                      "the collectives run over shared memory, oracle/mpi_shim) on the same FCIDUMP and start vector; "
                      f"loop time from the time-stamped iteration lines on stdout"
                      f" ({n_iter} iterations after 2)"}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the synthetic-integral generator is plain numpy: load the module by path, NOT through the fries_b200 package (whose
    # __init__ loads libfries_b200.so) -- nothing of the product runs in this arm
    import importlib.util
    spec = importlib.util.spec_from_file_location("fries_synth_standalone", os.path.join(ROOT, "fries_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(synth)
    SynthMol = synth.SynthMol
    # strong scaling: the workload is the same at every N, so this arm times the same thing whatever --gpus says
    sm = SynthMol(cfg["system"], cfg["seed"], frozen=False)
    keys, vals = reference_start_vector(cfg, sm)
    sec, ranks, err = reference_iter_seconds_best(cfg, sm, keys, vals, args.warmup, args.warmup + args.steps)
    sample = (f"{args.steps} iterations of oracle/_ref/frisys_mol after {args.warmup} warm-up iterations on {ranks} "
              "rank(s) = host cores (the reference's own MPI code; no MPI installation here: its collectives run "
              "between processes over shared memory, oracle/mpi_shim)")
    if sec is None:
        print(json.dumps({"impl": "reference", "unavailable": err}), flush=True)
        return
    v = round(1.0 / sec, 4)
    out = {"impl": "reference", "metric": "fri_iterations_per_sec", "value": v, "unit": "iter/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 3), "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "spawned_elements_per_sec": round(cfg["mat_nonz"] * v, 1),
           "config": {"workload": cfg["workload"], "vec_nonz": cfg["vec_nonz"], "mat_nonz": cfg["mat_nonz"]},
           "cpu_baseline": {"value": v, "unit": "iter/s", "cores": ranks, "kind": "reference", "sample": sample},
           "e2e": {"value": v, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def reference_start_vector(cfg, sm):
    """Same start vector as the GPU arm, computed with the plain-C oracle's H.v (no GPU needed)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oraclelib
    om = oraclelib.OracleMol(sm)
    hf = np.array([sm.hf], np.uint64)
    hf_en = om.diag(hf)[0]
    # H shifted by hf_en, as the DistVec diagonal shortcut does: (1 - 0.5 (H - E_HF)) applied twice
    k2, v2 = hf, np.ones(1)
    for npar in start_parents(cfg):
        order = np.argsort(-np.abs(v2), kind="stable")[:npar]
        k2, v2 = om.h_apply(k2[order], v2[order], 1.0 + 0.5 * hf_en, -0.5)
    order = np.argsort(-np.abs(v2), kind="stable")[: cfg["vec_nonz"]]
    keys, vals = k2[order], v2[order]
    vals = vals * (cfg["target"] / np.abs(vals).sum())
    perm = np.random.default_rng(7).permutation(keys.size)
    return np.ascontiguousarray(keys[perm]), np.ascontiguousarray(vals[perm])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # default: BASELINE.json configs[2], the configuration the metric "at 1/2/4/8 B200" is quoted on (H2O cc-pVDZ-sized,
    # 1e6-element vector, 1e6 matrix samples; total size FIXED as N grows: strong scaling).  --config ne: configs[1].
    ap.add_argument("--config", default="h2o", choices=sorted(CONFIGS))
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()

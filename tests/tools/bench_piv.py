"""Measure the pivotal sampler kernel (csrc/piv.cu: piv_samp_kernel) alone on one B200.

    python tests/tools/bench_piv.py [--n 20000000] [--n_samp 4000000] [--reps 6]

Inputs resident in HBM and larger than L2 (160 MB of values at the default size); CUDA events around every launch
(fries_ctx_set_profile).  Prints one JSON line: ms per launch, algorithmic GB/s (52 B per vector element
+ 20 B per sample incl. its two draws) against MEASURED_PEAKS.json:hbm_gbs.  A diagnostic, not the bench.py contract."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20_000_000)
    ap.add_argument("--n_samp", type=int, default=4_000_000)
    ap.add_argument("--reps", type=int, default=6)
    a = ap.parse_args()
    import fries_b200
    from fries_b200._capi import check, lib
    ctx = fries_b200.Context(0)
    g = torch.Generator(device="cuda").manual_seed(1)
    master = torch.rand(a.n, device="cuda", dtype=torch.float64, generator=g) * \
        torch.where(torch.rand(a.n, device="cuda", generator=g) < 0.5, -1.0, 1.0).double()
    keep0 = (torch.rand(a.n, device="cuda", generator=g) < 0.1).to(torch.uint8)
    norm = float(master.abs()[keep0 == 0].sum().item())
    draws = torch.randint(-2**31, 2**31 - 1, (2 * a.n_samp,), device="cuda", dtype=torch.int32, generator=g)
    work = torch.empty(8 * a.n + 12 * a.n_samp + 4096, dtype=torch.uint8, device="cuda")
    res = torch.zeros(4, dtype=torch.float64, device="cuda")
    vals, keep = master.clone(), keep0.clone()
    ctx.set_profile(1)
    times = []
    for rep in range(a.reps + 2):
        vals.copy_(master)
        keep.copy_(keep0)
        torch.cuda.synchronize()
        before = ctx.kernel_ms("piv_samp")[0]
        check(lib.fries_piv_samp_dev(ctx.h, vals.data_ptr(), a.n, norm, a.n_samp, keep.data_ptr(), draws.data_ptr(),
                                     work.data_ptr(), work.numel(), res.data_ptr()))
        ctx.sync()
        if rep >= 2:
            times.append(ctx.kernel_ms("piv_samp")[0] - before)
    r = res.cpu().numpy()
    assert int(r[1]) == a.n_samp and int(r[3]) == 0, r
    ms = float(np.mean(times))
    alg = 52.0 * a.n + 20.0 * a.n_samp
    peak = 6543.7
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    print(json.dumps({"kernel": "piv_samp", "n": a.n, "n_samp": a.n_samp, "ms_per_launch": round(ms, 4),
                      "ms_min": round(float(np.min(times)), 4), "algorithmic_bytes": alg,
                      "achieved_GBps": round(alg / ms / 1e6, 1), "peak_GBps": peak,
                      "frac": round(alg / ms / 1e6 / peak, 4), "elements_per_sec": round(a.n / ms * 1e3, 1),
                      "samples_drawn": int(r[1]), "anomalies": int(r[3])}))
    ctx.close()


if __name__ == "__main__":
    main()

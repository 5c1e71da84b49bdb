"""GPU tier, >= 2 GPUs: tests/multi_gpu_check.py under torchrun (one rank per GPU): in-kernel exchange ping-pong,
distributed find_preserve / sys_comp against the single-rank oracle, distributed piv_comp_parallel against the oracle's chain, 30 frisys_mol iterations on both spawn routes with
the ownership invariant, routed H.v against the single-GPU H.v, multi-rank frifull_mol.  Skipped on a one-GPU box (the
driver's GPU tier); run by hand with gpurun --gpus 2 / 8 during the round."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_multi_gpu_check_under_torchrun():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs, found {n}")
    world = 2 if n < 8 else 8
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "multi_gpu_check.py")],
                       cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("[multi]")]
    assert r.returncode == 0, r.stdout[-3000:]
    assert any("== single-rank oracle" in ln for ln in lines), lines
    assert any("distributed piv_comp_parallel" in ln for ln in lines), lines
    assert any("route p2p" in ln for ln in lines) and any("route nccl" in ln for ln in lines), lines
    assert any("routed H.v == single-GPU H.v" in ln for ln in lines), lines

"""GPU tier, after the verified files: measurement variants that are compiled into the library but not used by default.

FRIES_STAGE_CTAS=1 selects the one-CTA-per-SM build of the five HB-PP stage kernels (116 registers, no spills, half the
resident warps; csrc/hbpp.cu).  Same source, same arithmetic: the systematic-pipeline parity, golden and bracket tests must
pass unchanged with it.  Child process (the variant is chosen when the library first launches a stage)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def test_stage_kernels_one_cta_per_sm():
    env = dict(os.environ, FRIES_STAGE_CTAS="1")
    code = ("import ctypes, fries_b200._capi as c; v = ctypes.c_int(0); "
            "assert c.lib.fries_debug_stage_ctas(ctypes.byref(v)) == 0; print(v.value)")
    sel = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, stdout=subprocess.PIPE, text=True)
    assert sel.returncode == 0 and sel.stdout.strip() == "1"  # the children below launch the one-CTA build
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-k", "hbpp and not piv",
                        "tests/test_gpu_parity.py", "tests/test_gpu_golden.py", "tests/test_gpu_bracket.py",
                        "tests/test_hbpp_exact_limit.py"], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=500)
    print(r.stdout[-1500:])
    assert r.returncode == 0, r.stdout[-3000:]
    assert " passed" in r.stdout

"""CPU tier: the C-ABI library loads and exports every symbol include/fries_b200.h declares; without a GPU every
compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fries_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fries_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from fries_b200 import _capi
    names = declared_symbols()
    assert len(names) >= 45
    for n in names:
        assert hasattr(_capi.lib, n), f"{n} declared in include/fries_b200.h but not exported"
    assert set(_capi.EXPORTS) == set(names), set(_capi.EXPORTS) ^ set(names)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import fries_b200
    from fries_b200._capi import ERR_CUDA, FriesError
    with pytest.raises(FriesError) as e:
        fries_b200.Context(0)
    assert e.value.code == ERR_CUDA and "no CPU fallback" in str(e.value)


def test_product_does_not_touch_the_oracle():
    """nothing under fries_b200/ may import, link or execute oracle/ (or tests/hostcheck)"""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "fries_b200")):
        if "build" in dp or "__pycache__" in dp:
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"oracle[/_]|oraclelib|reflib|hostcheck|libfries_ref|fries_oracle", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_stage_kernel_variant_selection():
    """FRIES_STAGE_CTAS=1 selects the one-CTA-per-SM build of the stage kernels, anything else the product's two (read once
    per process: child processes)"""
    import subprocess
    import sys
    code = ("import ctypes, fries_b200._capi as c; v = ctypes.c_int(0); "
            "assert c.lib.fries_debug_stage_ctas(ctypes.byref(v)) == 0; print(v.value)")
    for env_val, want in ((None, "2"), ("1", "1"), ("0", "2"), ("12", "2")):
        env = dict(os.environ)
        env.pop("FRIES_STAGE_CTAS", None)
        if env_val is not None:
            env["FRIES_STAGE_CTAS"] = env_val
        r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, stdout=subprocess.PIPE, text=True)
        assert r.returncode == 0 and r.stdout.strip() == want, (env_val, r.stdout)

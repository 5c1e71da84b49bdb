"""CPU tier: the drivers' command-line failure modes, which come before any GPU work and follow the reference's argparse
use (Ext_Libs/argparse.hpp:347-393; frisys_mol.cpp:17-75, frifull_mol.cpp:15-50, frisys_hh.cpp:15-45): a missing required flag
lists every missing flag on stderr and exits with -1; an unknown --distribution is refused with the reference's message and
exit code 1; an unknown flag is only a warning."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "fries_b200", "host", "bin")

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(BIN, "frisys_mol")), reason="host programs not built")


def run(exe, *args):
    return subprocess.run([os.path.join(BIN, exe)] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                          text=True, timeout=60)


def test_missing_required_flags():
    r = run("frisys_mol", "--vec_nonz", 10)
    assert r.returncode == 255
    for flag in ("fcidump_path", "distribution", "mat_nonz", "max_dets", "epsilon"):
        assert f"Argument missing: --{flag}" in r.stderr
    assert "vec_nonz" not in r.stderr
    r = run("frifull_mol", "--vec_nonz", 10)
    assert r.returncode == 255 and "Argument missing: --hf_path" in r.stderr and "Argument missing: --max_dets" in r.stderr
    r = run("frisys_hh", "--max_dets", 5)
    assert r.returncode == 255 and "Argument missing: --params_path" in r.stderr and "Argument missing: --vec_nonz" in r.stderr


def test_legacy_input_needs_no_distribution_or_epsilon():
    """examples/run_neon.sh: --hf_path in place of --fcidump_path; --distribution and --epsilon then have defaults"""
    r = run("frisys_mol", "--hf_path", "/nonexistent/", "--vec_nonz", 10, "--mat_nonz", 10)
    assert r.returncode == 255 and "Argument missing: --max_dets" in r.stderr
    assert "distribution" not in r.stderr and "epsilon" not in r.stderr and "fcidump_path" not in r.stderr


def test_unknown_distribution_and_unknown_flag():
    r = run("frisys_mol", "--fcidump_path", "x", "--distribution", "NU2", "--vec_nonz", 10, "--mat_nonz", 10, "--max_dets", 100,
            "--epsilon", 0.01, "--bogus_flag", 1)
    assert r.returncode == 1
    assert "Error parsing command line" in r.stderr and "unrecognised commandline argument: bogus_flag" in r.stderr

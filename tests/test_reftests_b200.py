"""The reference's OWN test files and link example, unmodified, against the product's reference-signature header tree
(include/FRIES/*): /root/reference/tests/{unit_tests,test_compression,test_bitstrings,test_vector}.cpp and
examples/fries_test.cpp are compiled by oracle/Makefile (`make -C oracle ref`, only where /root/reference is mounted) with
-Iinclude and linked with libfries_b200.so; the binaries live in oracle/_ref (git-ignored, they travel to the GPU box).

CPU tier: the ten test cases that are host arithmetic in the reference as well (bit strings, sorted lists, excitation
bookkeeping, Neel / Hartree-Fock strings, phonon fields, alias method) and the link example.  GPU tier: all twelve, i.e.
also "[comp_preserve]" (find_preserve + sys_comp + sys_comp_serial + piv_samp_serial on the device) and "[vector_add]"
(DistVec<int>::add / perform_add / operator[] / indices through the device store)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "unit_tests_b200")
LINK = os.path.join(ROOT, "oracle", "_ref", "fries_test_b200")
HOST_CASES = "[binary],[hf_bits],[neel_bits],[flip_spins],[id_excite],[flip_connect],[ins_sorted],[ex_occ],[phonon_bits],[alias]"
need = pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/unit_tests_b200 not built (make -C oracle ref needs /root/reference)")


def run(args, tmp_path):
    out = str(tmp_path) + "/"
    return subprocess.run([EXE, "--out_path", out] + args, cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          text=True, timeout=300)


@need
def test_reference_host_side_cases_pass_unmodified(tmp_path):
    r = run([HOST_CASES], tmp_path)
    assert r.returncode == 0 and "All tests passed" in r.stdout and "10 test cases" in r.stdout, r.stdout[-2000:]


@need
def test_reference_link_example(tmp_path):
    r = subprocess.run([LINK], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.startswith("Binomial random number:"), r.stdout


@need
@pytest.mark.gpu
def test_reference_cases_pass_unmodified_on_gpu(tmp_path):
    r = run([], tmp_path)
    assert r.returncode == 0 and "All tests passed" in r.stdout and "12 test cases" in r.stdout, r.stdout[-3000:]

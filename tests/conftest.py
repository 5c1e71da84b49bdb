import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The built libraries are git-ignored: in a fresh checkout build them once before collection (the product library, the
    host programs, the oracle and -- where the reference's sources are mounted -- oracle/_ref).  A no-op when they exist."""
    need = [os.path.join(ROOT, "fries_b200", "libfries_b200.so"), os.path.join(ROOT, "oracle", "_build", "libfries_oracle.so")]
    if all(os.path.exists(p) for p in need):
        return
    import shutil
    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        return  # nothing to build with: the tests report the missing library themselves
    import __graft_entry__
    __graft_entry__.build()

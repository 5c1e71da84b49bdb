"""GPU tier: projected energy of the product's frisys_mol against the reference's own driver at a production-like size
(north_star: "within combined error bars").  tests/tools/energy_parity.py runs both on the same FCIDUMP / start vector / command
line with different seeds; this test runs a shortened form (Ne-sized molecule, a fifth of the vector and matrix budgets, 2500
iterations: ~1 min of reference on the host cores); the full-length run (2e4 iterations at the full Ne sizes) is recorded
under profiles/r02_energy_parity_ne.json."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "frisys_mol")), reason="oracle/_ref/frisys_mol not built")
def test_energy_within_error_bars_of_the_reference_ne_sized():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "energy_parity.py"), "--config", "ne", "--scale", "0.2",
                        "--iters", "2500", "--burn", "500"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=800)
    assert r.returncode == 0, r.stderr[-1500:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    print(out)
    assert "energy" in out["ours"] and "energy" in out["reference"], out
    # 3 sigma here (a shortened run in a test tier that must not flake); the recorded full-length run is held to 2 sigma
    assert abs(out["delta"]) < 3 * out["combined_sigma"] + 1e-6, out

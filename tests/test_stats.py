"""CPU tier: fries_b200/stats.py (the reference's Benchmarks/calc_stats.py on numpy) on processes with known answers."""
import numpy as np
import pytest

from fries_b200.stats import autocorr_function, integrated_time, trajectory_stats


def ar1(rng, n, rho):
    x = np.empty(n)
    x[0] = rng.standard_normal() / np.sqrt(1 - rho**2)
    e = rng.standard_normal(n)
    for i in range(1, n):
        x[i] = rho * x[i - 1] + e[i]
    return x


def test_autocorrelation_of_ar1():
    rng = np.random.default_rng(0)
    x = ar1(rng, 400000, 0.8)
    acf = autocorr_function(x)
    assert acf[0] == 1.0
    assert np.allclose(acf[1:6], 0.8 ** np.arange(1, 6), atol=0.01)
    # direct O(n t) evaluation of the same estimator
    y = x - x.mean()
    direct = np.array([np.dot(y[:len(y) - t], y[t:]) for t in range(4)]) / np.dot(y, y)
    assert np.allclose(acf[:4], direct, rtol=1e-10)


@pytest.mark.parametrize("rho", [0.0, 0.5, 0.9])
def test_integrated_time_of_ar1(rho):
    # exact: (1 + rho) / (1 - rho); Sokal's window with c = 2 truncates early and underestimates strongly correlated
    # chains by a known, bounded amount -- c = 6 recovers the exact value
    x = ar1(np.random.default_rng(1), 400000, rho)
    exact = (1 + rho) / (1 - rho)
    assert integrated_time(x, c=6) == pytest.approx(exact, rel=0.08)
    assert 0.5 * exact <= integrated_time(x, c=2) <= 1.1 * exact


def test_ratio_estimator_and_error_bar():
    # num = E * den + noise with correlated den: energy recovered within the reported error bar, in 40 of 40 seeds at 4 sigma
    E = -0.3217
    for seed in range(40):
        rng = np.random.default_rng(seed)
        den = 5.0 + ar1(rng, 20000, 0.7)
        num = E * den + 0.4 * ar1(rng, 20000, 0.5)
        st = trajectory_stats(num, den, burn_in=2000, exact=E)
        assert st["n"] == 18000 and st["iat"] >= 1.0
        assert abs(st["error"]) <= 4 * st["std_err"] * np.sqrt((1 + 0.5) / (1 - 0.5) / st["iat"]) + 1e-12
        assert st["efficiency"] == pytest.approx(1 / (st["variance"] * st["iat"]))


def test_cli_prints_the_reference_lines(tmp_path, capsys):
    from fries_b200 import stats
    rng = np.random.default_rng(3)
    den = 2.0 + 0.01 * rng.standard_normal(5000)
    np.savetxt(tmp_path / "projnum.txt", -0.5 * den + 0.001 * rng.standard_normal(5000))
    np.savetxt(tmp_path / "projden.txt", den)
    stats.main([str(tmp_path) + "/", "--burn_in", "1000", "--exact", "-0.5"])
    out = capsys.readouterr().out
    assert out.startswith("iat: ") and "Mean error ± 2 sigma (millihartrees) = " in out and "Efficiency: " in out

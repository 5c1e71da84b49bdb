"""Post-processing of a FRI trajectory: the reference's Benchmarks/calc_stats.py:1-36 on numpy.

The reference script reads projnum.txt / projden.txt, drops a burn-in, linearises the ratio estimator
E = <num> / <den> (calc_stats.py:24), and gets the standard error from the integrated autocorrelation time of the
linearised trajectory, computed by emcee.autocorr.integrated_time(x, c=2) (calc_stats.py:26).  emcee is a third-party
package that is neither vendored in the reference nor installed here and the reference pins no version; this module
restates the published algorithm it implements -- Sokal's automatic windowing: tau(M) = 1 + 2 sum_{t=1..M} rho(t) with
the smallest window M >= c * tau(M), rho from the FFT autocorrelation -- in a few lines of numpy.  Host-side analysis,
not part of the GPU hot path."""
from __future__ import annotations

import numpy as np


def autocorr_function(x: np.ndarray) -> np.ndarray:
    """normalised autocorrelation function rho(t), t = 0 .. len(x) - 1, by FFT (zero-padded to avoid wrap-around)"""
    x = np.asarray(x, np.float64)
    n = x.size
    m = 1 << int(np.ceil(np.log2(max(2 * n, 2))))
    f = np.fft.rfft(x - x.mean(), m)
    acf = np.fft.irfft(f * np.conj(f), m)[:n]
    if acf[0] <= 0:
        return np.concatenate([[1.0], np.zeros(n - 1)])
    return acf / acf[0]


def integrated_time(x: np.ndarray, c: float = 2.0) -> float:
    """integrated autocorrelation time with Sokal's window: smallest M with M >= c * tau(M)"""
    rho = autocorr_function(x)
    taus = 2.0 * np.cumsum(rho) - 1.0
    ok = np.arange(taus.size) >= c * taus
    window = int(np.argmax(ok)) if ok.any() else taus.size - 1
    return float(max(taus[window], 1.0))


def trajectory_stats(proj_num, proj_den, burn_in: int = 0, exact: float | None = None, c: float = 2.0) -> dict:
    """calc_stats.py:16-36: mean energy of the ratio estimator, its standard error and the statistical efficiency.

    energy = mean(num) / mean(den); the linearised trajectory num / <den> - <num> den / <den>^2 carries the variance
    (delta method); std_err = sqrt(var * iat / n); efficiency = 1 / (var * iat)."""
    num = np.asarray(proj_num, np.float64)
    den = np.asarray(proj_den, np.float64)
    n = min(num.size, den.size)
    num, den = num[burn_in:n], den[burn_in:n]
    if num.size < 4:
        raise ValueError("trajectory shorter than the burn-in")
    nm, dm = num.mean(), den.mean()
    traj = num / dm - nm * den / dm**2
    iat = integrated_time(traj, c)
    var = float(traj.var())
    out = {"energy": float(nm / dm), "iat": iat, "variance": var, "std_err": float(np.sqrt(var * iat / num.size)),
           "efficiency": float(1.0 / (var * iat)) if var > 0 else float("inf"), "n": int(num.size)}
    if exact is not None:
        out["error"] = out["energy"] - exact
    return out


def main(argv=None):
    """python -m fries_b200.stats <result_dir prefix> [--burn_in N] [--exact E]: the reference script's printout"""
    import argparse
    ap = argparse.ArgumentParser(description=main.__doc__)
    ap.add_argument("result_dir", help="prefix of projnum.txt / projden.txt (the drivers' --result_dir)")
    ap.add_argument("--burn_in", type=int, default=40000)
    ap.add_argument("--exact", type=float, default=None, help="exact correlation energy (hartree)")
    a = ap.parse_args(argv)
    st = trajectory_stats(np.genfromtxt(a.result_dir + "projnum.txt"), np.genfromtxt(a.result_dir + "projden.txt"),
                          a.burn_in, a.exact)
    print("iat: " + str(st["iat"]))
    if a.exact is not None:
        print("Mean error ± 2 sigma (millihartrees) = {0:.2f} ± {1:.2f}".format(st["error"] * 1e3, 2e3 * st["std_err"]))
    else:
        print("Energy ± 2 sigma (hartree) = {0:.8f} ± {1:.8f}".format(st["energy"], 2 * st["std_err"]))
    print("Efficiency: " + str(st["efficiency"]))


if __name__ == "__main__":
    main()

"""Generate tests/golden/sokal_golden.npz from the REFERENCE's own sokal()
(/root/reference/src/sokal.cc compiled by oracle/Makefile into
oracle/_ref/libsokal_ref.so).  Run in the build container (the reference tree
does not exist on the GPU box):  python tools/make_golden_sokal.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402


def ar1(rng, n, phi):
    x = np.zeros(n)
    e = rng.standard_normal(n)
    x[0] = e[0] / np.sqrt(1 - phi * phi)
    for i in range(1, n):
        x[i] = phi * x[i - 1] + e[i]
    return x


def main():
    assert orc.ref_sokal_lib() is not None, "build oracle/_ref first: make -C oracle"
    rng = np.random.default_rng(20260101)
    series = []
    for n in (4, 8, 64, 1024, 2048):
        series.append(rng.standard_normal(n))
    for phi in (0.3, 0.5, 0.9, 0.99, -0.5):
        series.append(ar1(rng, 1024, phi) * 0.3 - 2.0)
    series.append(np.log(rng.gamma(3.1, 0.5, 1024)))           # a log-Gamma trace, like log(mu)
    series.append(np.linspace(-1.0, 1.0, 1024))                  # trend: window never closes early
    series.append(np.sin(np.arange(1024) * 0.05) + 0.01 * rng.standard_normal(1024))
    series.append(np.full(1024, 1.25))                            # constant: var 0, tau nan
    out = {}
    meta = []
    for i, x in enumerate(series):
        rc, var, tau, m = orc.sokal_reference(x)
        out[f"x{i}"] = x
        meta.append([rc, var, tau, m])
    for n in (3, 6, 1000):                                        # failure codes
        rc = orc.sokal_reference(rng.standard_normal(n))[0]
        meta.append([rc, np.nan, np.nan, n])
    out["meta"] = np.array(meta, np.float64)
    out["nseries"] = np.array(len(series))
    path = os.path.join(ROOT, "tests", "golden", "sokal_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "\n", out["meta"])


if __name__ == "__main__":
    main()

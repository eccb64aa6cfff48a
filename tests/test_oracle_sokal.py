"""orc_sokal (restatement) pinned against golden vectors produced by the
reference's own src/sokal.cc (tools/make_golden_sokal.py), and — when
oracle/_ref exists — against that library live.  CPU only."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden", "sokal_golden.npz")


def _cases():
    g = np.load(GOLD)
    n = int(g["nseries"])
    return g, n


def test_restated_sokal_matches_reference_golden():
    g, n = _cases()
    meta = g["meta"]
    for i in range(n):
        rc, var, tau, m = orc.sokal(g[f"x{i}"])
        erc, evar, etau, em = meta[i]
        assert rc == int(erc)
        assert m == int(em), (i, m, em)
        assert np.isclose(var, evar, rtol=1e-10, atol=1e-300)
        if np.isnan(etau):
            assert np.isnan(tau)
        else:
            assert np.isclose(tau, etau, rtol=1e-9, atol=1e-12)
    # failure codes: src/sokal.cc:108-126
    assert orc.sokal(np.zeros(3))[0] == 200
    assert orc.sokal(np.zeros(6))[0] == 201
    assert orc.sokal(np.zeros(1000))[0] == 201


def test_restated_sokal_matches_reference_live():
    if orc.ref_sokal_lib() is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(5)
    for n in (16, 256, 1024):
        for _ in range(20):
            x = np.cumsum(rng.standard_normal(n)) * 0.05 + rng.standard_normal(n)
            a = orc.sokal(x); b = orc.sokal_reference(x)
            assert a[0] == b[0] == 0 and a[3] == b[3]
            assert np.isclose(a[1], b[1], rtol=1e-10) and np.isclose(a[2], b[2], rtol=1e-9)


def test_sokal_known_answers():
    """white noise tau ~ 1; AR(1) tau -> (1+phi)/(1-phi)."""
    rng = np.random.default_rng(3)
    taus = [orc.sokal(rng.standard_normal(1024))[2] for _ in range(200)]
    assert abs(np.mean(taus) - 1.0) < 0.1
    phi = 0.5
    taus = []
    for _ in range(200):
        e = rng.standard_normal(2048)
        x = np.zeros(2048)
        for i in range(1, 2048):
            x[i] = phi * x[i - 1] + e[i]
        taus.append(orc.sokal(x)[2])
    assert abs(np.mean(taus) - 3.0) < 0.35

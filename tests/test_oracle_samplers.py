"""The shared samplers (include/mmq_sampler.h, compiled into the oracle by gcc)
against known answers and against scipy; and against the independent GSL-style
samplers of oracle/mmseq_oracle.cpp.  CPU only."""
import numpy as np
import pytest
from scipy import special, stats

from oracle import oracle as orc


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert [hex(v) for v in orc.philox([0, 0, 0, 0], [0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(v) for v in orc.philox([0xffffffff] * 4, [0xffffffff] * 2)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(v) for v in orc.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_uniform_open_interval_and_flat():
    u = orc.draw_uniform(7, 200000)
    assert u.min() > 0.0 and u.max() < 1.0
    assert stats.kstest(u, "uniform").pvalue > 1e-4


def test_log_exp_ndtri_accuracy():
    rng = np.random.default_rng(1)
    x = np.exp(rng.uniform(-700, 700, 200000))
    assert np.max(np.abs(orc.math_fn("log", x) - np.log(x)) / np.maximum(np.abs(np.log(x)), 1e-300)) < 4e-16
    sub = np.array([5e-324, 1e-310, 2.2250738585072014e-308])
    assert np.allclose(orc.math_fn("log", sub), np.log(sub), rtol=1e-15)
    y = rng.uniform(-708, 709, 200000)
    assert np.max(np.abs(orc.math_fn("exp", y) / np.exp(y) - 1)) < 4e-16
    ysub = rng.uniform(-745, -708, 20000)   # subnormal results: absolute accuracy of one subnormal ulp
    assert np.max(np.abs(orc.math_fn("exp", ysub) - np.exp(ysub))) <= 2 * 5e-324
    assert orc.math_fn("exp", np.array([-800.0]))[0] == 0.0 and np.isinf(orc.math_fn("exp", np.array([800.0]))[0])
    p = np.concatenate([rng.uniform(0, 1, 100000), 10.0 ** rng.uniform(-300, -1, 2000), [1e-9, 1 - 1e-9, 0.5]])
    ref = special.ndtri(p)
    got = orc.math_fn("ndtri", p)
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3)) < 1e-13
    z = -10.0 ** rng.uniform(-18, -0.31, 10000)
    assert np.allclose(orc.math_fn("log1p", z), np.log1p(z), rtol=2e-15, atol=0)


@pytest.mark.parametrize("a,rate", [(0.1, 0.1), (0.1, 3.0), (0.9, 1.0), (1.0, 2.0), (1.1, 0.5), (7.1, 1e-3), (500.1, 2.0), (1e6 + 0.1, 30.0)])
def test_gamma_distribution(a, rate):
    x = orc.draw_gamma(11, 100000, a, rate)
    assert np.all(x >= 0)
    assert stats.kstest(x, "gamma", args=(a, 0, 1.0 / rate)).pvalue > 1e-4
    # analytic anchor of src/mmseq.cpp:1372-1373: E log mu = psi(a) - log(rate), sd = sqrt(psi1(a))
    lg = np.log(x[x > 0])
    assert abs(lg.mean() - (special.digamma(a) - np.log(rate))) < 5 * np.sqrt(special.polygamma(1, a) / len(lg)) + 1e-12


@pytest.mark.parametrize("n,p", [(1, 0.3), (2, 0.5), (7, 0.01), (50, 0.93), (1000, 0.3), (1000, 0.0005), (100000, 0.5),
                                   (2000000, 1e-5), (2000000000, 0.25), (30, 0.4), (25, 0.39), (40, 0.26)])
def test_binomial_distribution(n, p):
    cnt = 200000
    x = orc.draw_binomial(3, cnt, n, p)
    assert x.min() >= 0 and x.max() <= n
    lo, hi = int(stats.binom.ppf(1e-5, n, p)), int(stats.binom.ppf(1 - 1e-5, n, p))
    edges = np.arange(lo, hi + 2)
    if len(edges) > 60:  # coarse bins for wide supports
        edges = np.unique(np.linspace(lo, hi + 1, 50).astype(np.int64))
    obs = np.histogram(x, bins=np.concatenate([[-0.5], edges[1:-1] - 0.5, [n + 0.5]]))[0]
    cdf = stats.binom.cdf(np.concatenate([edges[1:-1] - 1, [n]]), n, p)
    exp = np.diff(np.concatenate([[0.0], cdf])) * cnt
    keep = exp > 5
    obs = np.concatenate([obs[keep], [obs[~keep].sum()]]); exp = np.concatenate([exp[keep], [exp[~keep].sum()]])
    if exp[-1] == 0:
        obs, exp = obs[:-1], exp[:-1]
    chi2 = ((obs - exp) ** 2 / exp).sum()
    assert stats.chi2.sf(chi2, len(exp) - 1) > 1e-5, (n, p, chi2)


def test_alloc_known_answers():
    """The random-stream contract of mmq_alloc_row is pinned by committed vectors
    (tools/make_golden_alloc.py): k = 1 (CAT stream, quad-shared block), categorical draws across
    block boundaries up to MMQ_CAT_K = 64, binomial chains above."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "alloc_kat.npz"))
    assert int(g["n_cases"]) >= 9
    for i in range(int(g["n_cases"])):
        x = orc.draw_alloc(int(g[f"seed_{i}"]), 8, g[f"p_{i}"], int(g[f"k_{i}"]))
        assert np.array_equal(x, g[f"x_{i}"]), f"case {i}: the allocation contract moved"
        assert (x.sum(axis=1) == int(g[f"k_{i}"])).all()


def test_alloc_row_is_multinomial_and_conserves():
    p = np.array([3.0, 0.0, 1.0, 1e-30, 6.0])
    for k in (1, 2, 5, 40, 100000):
        x = orc.draw_alloc(5, 50000, p, k)
        assert (x.sum(axis=1) == k).all()
        assert (x[:, 1] == 0).all()          # zero-probability members never receive fragments
        mean = x.mean(axis=0) / k
        assert np.allclose(mean, p / p.sum(), atol=5 * np.sqrt(0.25 / (50000 * min(k, 50))) + 1e-6)
    # singleton classes are deterministic
    assert (orc.draw_alloc(5, 10, np.array([2.0]), 17) == 17).all()
    # all-zero row: everything to the last member (documented convention)
    assert (orc.draw_alloc(5, 10, np.zeros(3), 4)[:, 2] == 4).all()


@pytest.mark.parametrize("k,reps", [(30, 100000), (64, 50000), (65, 50000), (700, 20000), (8193, 4000)])
def test_alloc_matches_gsl_style_multinomial_in_distribution(k, reps):
    """Every regime of mmq_alloc_row against the GSL-style chain of binomials (the reference's
    gsl_ran_multinomial, src/mmseq.cpp:880): categorical draws with 32-bit uniforms four to a Philox
    block (k <= MMQ_CAT_K = 64), binomial chain above."""
    p = np.array([0.5, 2.5, 1.0, 4.0])
    a = orc.draw_alloc(9, reps, p, k)
    b = orc.gsl_multinomial(9, reps, p, k)
    assert (a.sum(axis=1) == k).all() and (b.sum(axis=1) == k).all()
    for j in range(len(p)):
        if k <= 100:
            ha = np.bincount(a[:, j], minlength=k + 1); hb = np.bincount(b[:, j], minlength=k + 1)
        else:   # coarse bins around the mean for the large counts
            q = p[j] / p.sum()
            edges = k * q + np.sqrt(k * q * (1 - q)) * np.linspace(-3, 3, 13)
            ha = np.histogram(a[:, j], bins=np.concatenate([[-1], edges, [k + 1]]))[0]; hb = np.histogram(b[:, j], bins=np.concatenate([[-1], edges, [k + 1]]))[0]
        keep = (ha + hb) > 20
        assert stats.chi2_contingency(np.vstack([ha[keep], hb[keep]]))[1] > 1e-5
    # covariance structure: cov(x_i, x_j) = -k p_i p_j
    q = p / p.sum()
    cov = np.cov(a.T) / k
    assert np.allclose(cov, np.diag(q) - np.outer(q, q), atol=0.15 / 30 if k == 30 else 0.02)


@pytest.mark.parametrize("n,p", [(20, 0.3), (500, 0.2), (100000, 0.7)])
def test_gsl_style_binomial_distribution(n, p):
    x = orc.gsl_binomial(17, 100000, n, p)
    assert abs(x.mean() - n * p) < 5 * np.sqrt(n * p * (1 - p) / 100000)
    assert abs(x.var() / (n * p * (1 - p)) - 1) < 0.03
    y = orc.draw_binomial(17, 100000, n, p)
    assert stats.ks_2samp(x, y).pvalue > 1e-4


def test_gsl_style_gamma_distribution():
    for a, scale in [(0.1, 2.0), (3.1, 0.5), (200.1, 0.01)]:
        x = orc.gsl_gamma(5, 100000, a, scale)
        assert stats.kstest(x, "gamma", args=(a, 0, scale)).pvalue > 1e-4

"""Oracle EM / init against the analytic anchors of SURVEY.md section 4.  CPU only."""
import numpy as np

from oracle import oracle as orc


def test_tiny_hand_example_init_mu():
    # 3 transcripts, eff. len 1000; classes {A}:10, {A,B}:5, {B,C}:3  => N = 18, l = 1.8e-5
    row_ptr = [0, 1, 3, 5]; col = [0, 0, 1, 1, 2]; k = [10, 5, 3]
    l = np.full(3, 1000 * 18 / 1e9)
    P = orc.Problem(row_ptr, col, k, l)
    mu, uh, cs = P.init_mu()
    assert np.allclose(mu, np.array([12.5, 4.0, 1.5]) / 1.8e-5, rtol=1e-15)
    assert list(uh) == [10, 0, 0]
    assert cs[0, 0] == 10 and cs[0, 1] == 5 and cs[1, 1] == 8 and cs[2, 1] == 3


def test_em_mass_conservation_and_monotone(small_problem):
    h = small_problem
    P = orc.Problem(h.row_ptr, h.col, h.k, h.len)
    mu0, uh, _ = P.init_mu()
    assert np.isclose((mu0 * h.len).sum(), h.N, rtol=1e-12)   # init also conserves mass
    ll_prev = P.loglik(mu0)
    mu = mu0
    for it in range(5):
        mu, iters, ll, llr = P.em(mu, max_iter=1, eps=-1.0)
        assert iters == 1
        assert np.isclose((mu * h.len).sum(), h.N, rtol=1e-12)   # sum_t mu_t l_t = N after any EM step
        assert ll >= ll_prev - 1e-9 * abs(ll_prev)                # log-likelihood non-decreasing
        assert np.isclose(llr, ll - ll_prev, rtol=0, atol=1e-6 * abs(ll))
        ll_prev = ll


def test_em_stop_rule(small_problem):
    h = small_problem
    P = orc.Problem(h.row_ptr, h.col, h.k, h.len)
    mu0, _, _ = P.init_mu()
    mu, iters, ll, llr = P.em(mu0, max_iter=1000, eps=0.1)
    assert 1 <= iters < 1000 and llr <= 0.1
    mu2, iters2, _, llr2 = P.em(mu0, max_iter=iters - 1, eps=0.1)
    assert iters2 == iters - 1 and llr2 > 0.1                   # stops on max_iter first
    _, iters3, _, _ = P.em(mu0, max_iter=0, eps=0.1)
    assert iters3 == 0


def test_unique_hit_transcript_has_closed_form_posterior():
    """A transcript whose classes are all singletons: posterior exactly Gamma(alpha+k, beta+l)
    (src/mmseq.cpp:1372-1373 closed form) — both chains must reproduce E log mu."""
    from scipy import special
    row_ptr = [0, 1, 2]; col = [0, 1]; k = [40, 7]
    l = np.array([2e-3, 5e-4])
    P = orc.Problem(row_ptr, col, k, l, alpha=0.1, beta=0.1)
    mu0, _, _ = P.init_mu()
    _, tr, = P.gibbs_replay(mu0, seed=1234, first_sweep=0, n_sweeps=4096, stride=1, trace_len=4096)
    _, tr2, _ = P.gibbs_gsl(mu0, seed=1234, n_sweeps=4096, stride=1, trace_len=4096, threads=1)
    for t, kk in enumerate(k):
        want = special.digamma(0.1 + kk) - np.log(0.1 + l[t])
        sd = np.sqrt(special.polygamma(1, 0.1 + kk))
        for trace in (tr, tr2):
            lg = np.log(trace[t])
            assert abs(lg.mean() - want) < 5 * sd / np.sqrt(4096)
            assert abs(lg.std() / sd - 1) < 0.08

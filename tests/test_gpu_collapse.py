"""mmcollapse's covariance step on the GPU (src/mmcollapse.cpp:483-561) against the oracle,
through the C ABI.  Tolerances, in correlation units (|cov_gpu - cov_oracle| / (sd_i sd_j)): the tensor-core products are
split bf16 (nsplit terms per value, fp32 accumulation): 1e-2 for nsplit = 1, 3e-5 for 2, 1e-5 for 3 (fp32 accumulation over 6 L terms is the floor); everything else
(means, variances = the diagonal, the scaling back to covariance, mean_corrs) is fp64: 1e-12."""
import numpy as np
import pytest

from mmseq_b200 import capi
from oracle import oracle as orc
from tests.collapse_case import make_traces

pytestmark = pytest.mark.gpu
TOL = {1: 1e-2, 2: 3e-5, 3: 1e-5}


def corr_err(R, ref):
    d = np.sqrt(np.diag(ref))
    d = np.where(d > 0, d, 1.0)
    return np.max(np.abs(R - ref) / (d[:, None] * d[None, :]))


@pytest.mark.parametrize("L,C", [(64, 1), (64, 5), (1024, 37), (1024, 128), (1024, 129), (1024, 300), (512, 1000), (2048, 257)])
def test_cov_matches_oracle(L, C):
    M = make_traces(L, C, seed=L + C)
    ref = orc.trace_cov(M)
    for nsplit in (1, 2, 3):
        R = capi.trace_cov(M, nsplit=nsplit)
        assert np.array_equal(R, R.T), "not exactly symmetric"
        assert np.allclose(np.diag(R), np.diag(ref), rtol=1e-12, atol=0.0), "diagonal is the fp64 variance"
        e = corr_err(R, ref)
        assert e < TOL[nsplit], (nsplit, e)
    if C > 8:
        d = np.sqrt(np.diag(R))
        assert abs(R[0, C - 1] / (d[0] * d[C - 1]) - 1.0) < TOL[3]     # duplicated feature


def test_single_cta_kernel_gives_the_same_matrix(monkeypatch):
    """k_cov_gemm2 (CTA pairs, the default) and k_cov_gemm (MMQ_COV_PAIR=0) accumulate the same products in the same order."""
    M = make_traces(1024, 700, seed=31)
    R2 = capi.trace_cov(M, nsplit=2)
    monkeypatch.setenv("MMQ_COV_PAIR", "0")
    R1 = capi.trace_cov(M, nsplit=2)
    assert np.array_equal(R1, R1.T) and corr_err(R1, R2) < 1e-6


def test_cov_full_size_tile_grid():
    """Several tiles per side incl. a ragged last tile; every tile of the upper triangle and its mirror written."""
    L, C = 1024, 5 * 128 + 77
    M = make_traces(L, C, seed=99)
    ref = orc.trace_cov(M)
    R = capi.trace_cov(M, nsplit=2)
    assert np.array_equal(R, R.T)
    assert corr_err(R, ref) < TOL[2]
    r_gpu = R / np.sqrt(np.outer(np.diag(R), np.diag(R)))
    r_ref = ref / np.sqrt(np.outer(np.diag(ref), np.diag(ref)))
    assert np.sqrt(np.mean((r_gpu - r_ref) ** 2)) < 2e-6                 # typical error far below the bound


def test_cov_non_finite_and_constant_features():
    M = make_traces(256, 140, seed=4)
    M[17, 3] = np.nan
    M[5, 130] = np.inf
    M[:, 77] = 2.5
    ref = orc.trace_cov(M)
    R = capi.trace_cov(M, nsplit=2)
    for c in (3, 130, 77):
        assert np.all(R[c, :] == 0) and np.all(R[:, c] == 0)
    assert np.all(np.isfinite(R)) and corr_err(R, ref) < TOL[2]


def test_cov_rejects_bad_arguments():
    with pytest.raises(capi.MmqError):
        capi.trace_cov(np.ones((100, 4)))          # L not a multiple of 64
    with pytest.raises(capi.MmqError):
        capi.trace_cov(np.ones((128, 4)), nsplit=4)


def test_cov_device_pointers_and_stream():
    import torch
    L, C = 1024, 384
    M = make_traces(L, C, seed=8)
    dev = torch.device("cuda:0")
    Md = torch.from_numpy(np.ascontiguousarray(M.T)).to(dev)            # [C][L]: column c of M contiguous
    Rd = torch.empty((C, C), dtype=torch.float64, device=dev)
    ws = torch.empty(capi.trace_cov_workspace_bytes(L, C, 2), dtype=torch.uint8, device=dev)
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        capi.trace_cov_dev(Md.data_ptr(), L, C, 2, Rd.data_ptr(), ws.data_ptr(), st.cuda_stream)
    st.synchronize()
    R = Rd.cpu().numpy()
    assert np.array_equal(R, capi.trace_cov(M, nsplit=2))


def test_cov_from_a_handles_trace(small_problem):
    p = small_problem
    H = capi.Handle(p.row_ptr, p.col, p.k, p.len)
    H.init_mu()
    H.em(50, 0.1)
    L = 256
    H.gibbs(7, 0, L * 4, stride=4, trace_len=L)
    tr = H.get_trace()                                                   # [n][L]
    feats = np.arange(0, p.n, 2, dtype=np.int32)[::-1].copy()            # any order, any subset
    R = H.trace_cov(feats, nsplit=3)
    ref = orc.trace_cov(tr[feats].T)
    assert corr_err(R, ref) < TOL[3]
    assert np.array_equal(R, capi.trace_cov(tr[feats].T, nsplit=3))
    H.close()


def test_mean_corrs_matches_oracle():
    rng = np.random.default_rng(7)
    ns, C = 6, 150
    R = np.stack([capi.trace_cov(make_traces(256, C, seed=20 + s), nsplit=2) for s in range(ns)])
    S = (rng.random((C, ns)) < 0.8).astype(np.uint8)
    S[1, :] = 0
    S[2, :] = 0
    S[2, 1] = 1
    ts = np.arange(C)
    for pen in (0.0, 0.5):
        V, W = capi.mean_corrs(R, S, ts, pen)
        Vo, Wo = orc.mean_corrs(R, S, ts, pen)
        # the sd is a difference of two nearly equal terms when the correlations agree over the samples: sqrt(1e-16) = 1e-8
        assert np.allclose(V, Vo, rtol=1e-12, atol=1e-14 + 1e-7 * pen, equal_nan=True)
        assert np.allclose(W, Wo, rtol=1e-9, atol=1e-7, equal_nan=True)
    V0, W0 = capi.mean_corrs(R, S, ts)
    V1, W1 = capi.mean_corrs(R, S, [4, 9], 0.0, V0.copy() + 1.0, W0.copy())
    Vo, Wo = orc.mean_corrs(R, S, [4, 9], 0.0, V0.copy() + 1.0, W0.copy())
    assert np.allclose(V1, Vo, rtol=1e-12, atol=1e-14, equal_nan=True)

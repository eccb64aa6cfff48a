"""BASELINE.json configs 3 and 5 at parity-test size.
  config 3: haplotype-specific transcriptome (every transcript as _A/_B copies, deep multi-mapping)
            with gene and isoform-proportion aggregation;
  config 5: a batch of samples, one independent handle / stream / chain per sample."""
import numpy as np
import pytest

from mmseq_b200 import capi, hostlib, synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def test_config3_haplotype_transcriptome_with_gene_aggregation():
    s = synth.Synth(20260101 + 3, 400, 30000, haplo=True)
    h = hostlib.from_records(s.T, s.efflen, s.frag_ptr, s.frag_tid)
    assert np.diff(h.row_ptr).mean() > 5           # deep multi-mapping: both haplotype copies are usually hit
    P = orc.Problem(h.row_ptr, h.col, h.k, h.len)
    mu0, uh_o, _ = P.init_mu()
    mu_em, it_o, _, _ = P.em(mu0, 1000, 0.1)
    L, stride = 64, 2
    _, tr_o = P.gibbs_replay(mu_em, 1234, 0, L * stride, stride, L)
    gene_of_col = s.gene_of[h.col2hdr]
    G = s.G
    gptr = np.concatenate([[0], np.cumsum(np.bincount(gene_of_col, minlength=G))])
    members = np.argsort(gene_of_col, kind="stable").astype(np.int32)
    rp, col, k, cid = hostlib.sort_classes_by_cost(h)        # the host program's device order
    with capi.Handle(rp, col, k, h.len, class_id=cid) as H:
        assert np.array_equal(H.init_mu(), uh_o)
        it, _, _ = H.em(1000, 0.1)
        assert it == it_o and np.max(np.abs(H.get_mu() / mu_em - 1)) <= 1e-6
        H.set_mu(mu_em)
        H.gibbs(1234, 0, L * stride, stride=stride, trace_len=L)
        tr = H.get_trace()
        assert np.array_equal(tr, tr_o)                        # bit-exact chain
        H.set_groups(capi.MMQ_GROUP_GENE, gptr, members)
        gt = H.get_group_trace(2)
        multi = (np.diff(gptr)[gene_of_col] > 1).astype(np.uint8)
        PR = H.prop_summaries(gene_of_col, multi, [3, 32, 60], want_trace=True)
        S = H.summarize(2, [3, 32, 60])
        uh_gene = H.unique_hits_sets(gene_of_col.astype(np.int32), G)
    gt_o = np.zeros((G, L)); np.add.at(gt_o, gene_of_col, tr_o)
    assert np.allclose(gt, gt_o, rtol=1e-13)
    obs = np.diff(gptr) > 0
    prop = tr_o / gt_o[gene_of_col]
    assert np.allclose(PR["prop_trace"], prop, rtol=1e-15) and np.allclose(PR["mean_prop"], prop.mean(axis=1), rtol=1e-12)
    # the two haplotype copies of a transcript and its siblings share the gene: proportions sum to one
    sums = np.zeros((G, L)); np.add.at(sums, gene_of_col, PR["prop_trace"])
    assert np.allclose(sums[obs], 1.0, rtol=1e-12)
    O = orc.summaries_transcripts(gt_o[obs])
    assert np.allclose(S["log_mean"][obs], O["log_mu"], rtol=1e-12) and np.array_equal(S["win"][obs], O["win"])
    assert np.array_equal(uh_gene, orc.uh_literal(h.row_ptr, h.col, h.k, gptr, members))


def test_config5_batch_of_samples_independent_chains():
    import torch
    samples = []
    for i in range(4):
        s = synth.Synth(20260101 + 5, 250, 8000, frag_seed=i)       # same transcriptome, different samples
        h = hostlib.from_records(s.T, s.efflen, s.frag_ptr, s.frag_tid, layout=[hostlib.LAYOUT_COLLAPSED, hostlib.LAYOUT_PER_FRAGMENT_BY_LENGTH][i % 2])
        samples.append(h)
    handles = [capi.Handle(h.row_ptr, h.col, h.k, h.len) for h in samples]
    streams = [torch.cuda.Stream() for _ in samples]
    for H, st in zip(handles, streams):
        H.set_stream(st.cuda_stream)
        H.init_mu()
    mus = [H.get_mu() for H in handles]
    L, stride = 16, 4
    for chunk in range(4):                                            # interleave the samples' sweeps
        for i, H in enumerate(handles):
            H.gibbs(100 + i, chunk * 16, 16, stride=stride, trace_len=L)
    for i, (H, h) in enumerate(zip(handles, samples)):
        P = orc.Problem(h.row_ptr, h.col, h.k, h.len)
        mu_end, tr_o = P.gibbs_replay(mus[i], 100 + i, 0, 64, stride, L)
        assert np.array_equal(H.get_trace(), tr_o) and np.array_equal(H.get_mu(), mu_end)
        S = H.summarize(0, [1, 8, 14])
        O = orc.summaries_transcripts(tr_o, (5, 50, 95))
        assert np.allclose(S["log_mean"], O["log_mu"], rtol=1e-12, atol=1e-12)
        assert np.allclose(S["tau"], O["iact"], rtol=1e-8, atol=1e-10, equal_nan=True)   # Sokal IACT per sample
        H.close()


def test_config5_eight_handles_per_gpu_on_every_gpu_from_threads():
    """The batch driver's shape (bench.py --batch, `mmseq -batch`): 8 samples in flight per GPU on every visible GPU,
    each driven by its own host thread on its own stream; every chain equals its replay, Sokal summaries per sample."""
    import threading
    import torch
    ngpu = torch.cuda.device_count()
    per_gpu = 8
    jobs = [(g, i) for g in range(ngpu) for i in range(per_gpu)]
    results, errors = {}, []

    def work(g, i):
        try:
            s = synth.Synth(20260101 + 5, 200, 6000, frag_seed=100 * g + i)
            h = hostlib.from_records(s.T, s.efflen, s.frag_ptr, s.frag_tid)
            with capi.Handle(h.row_ptr, h.col, h.k, h.len, device=g) as H:
                H.init_mu()
                it, _, _ = H.em(1000, 0.1)
                mu_em = H.get_mu()
                H.gibbs(7 + i, 0, 64, stride=4, trace_len=16)
                S = H.summarize(0, [1, 8, 14])
                results[(g, i)] = (h, it, mu_em, H.get_trace(), H.get_mu(), S)
        except Exception as e:   # pragma: no cover
            errors.append(repr(e))

    th = [threading.Thread(target=work, args=j) for j in jobs]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors
    assert len(results) == ngpu * per_gpu
    for (g, i), (h, it, mu_em, tr, mu_end, S) in results.items():
        P = orc.Problem(h.row_ptr, h.col, h.k, h.len)
        mu_o, it_o, _, _ = P.em(P.init_mu()[0], 1000, 0.1)
        assert it == it_o and np.max(np.abs(mu_em / mu_o - 1)) <= 1e-6
        mu_r, tr_o = P.gibbs_replay(mu_em, 7 + i, 0, 64, 4, 16)
        assert np.array_equal(tr, tr_o) and np.array_equal(mu_end, mu_r), (g, i)
        O = orc.summaries_transcripts(tr_o, (5, 50, 95))
        assert np.allclose(S["log_mean"], O["log_mu"], rtol=1e-12, atol=1e-12)
        assert np.allclose(S["tau"], O["iact"], rtol=1e-8, atol=1e-10, equal_nan=True)


def test_block_cache_recycles_memory_without_changing_results(small_problem):
    """mmq_destroy gives its device blocks to the per-device cache and the next mmq_create takes them back zeroed: the chain of a
    handle does not depend on what ran before it in the process; mmq_release_cache returns the memory to the driver."""
    import torch
    p = small_problem

    def run(seed):
        with capi.Handle(p.row_ptr, p.col, p.k, p.len) as H:
            H.init_mu()
            H.em(100, 0.1)
            H.gibbs(seed, 0, 48, stride=4, trace_len=12)
            return H.get_trace().copy(), H.get_mu().copy()

    first = run(11)
    # something else in between leaves other data in the recycled blocks
    rng = np.random.default_rng(0)
    capi.trace_cov(np.exp(rng.normal(0, 1, (256, 300))), nsplit=2)
    run(12)
    again = run(11)
    assert np.array_equal(first[0], again[0]) and np.array_equal(first[1], again[1])
    free0, _ = torch.cuda.mem_get_info()
    assert capi.release_cache() == 0
    free1, _ = torch.cuda.mem_get_info()
    assert free1 >= free0
    third = run(11)
    assert np.array_equal(first[0], third[0])

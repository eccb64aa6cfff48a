"""Host side of the class plan (mmseq_b200/csrc/mmq_cls_plan.h: how mmq_create re-orders a
collapsed shard into slots / chunks for k_alloc_cls) built on the CPU and replayed with the shared
sampler (oracle.cls_plan_replay): the counts must be those of the canonical replay of
mmq_alloc_row, class by class — whatever the slot order, the number of host threads that built the
plan, the class-id base or an explicit class-id permutation.  The GPU side of the same contract is
tests/test_gpu_parity.py::test_class_plan_kernel_bit_exact."""
import numpy as np
import pytest

from oracle import oracle as orc


def _problem(seed=5, n=4000):
    rng = np.random.default_rng(seed)
    sizes = np.concatenate([rng.integers(1, 17, 5000), rng.integers(17, 65, 300), rng.integers(65, 200, 20), [2] * 37, [16] * 33])
    rows = [np.sort(rng.choice(n, size=int(d), replace=False)) for d in sizes]
    rows[10] = np.array([0, 1, 2]); rows[11] = np.array([0, 1]); rows[12] = np.array([1, 2, 3, 5])
    m = len(rows)
    k = np.where(rng.random(m) < 0.5, 1, rng.integers(0, 70, m))
    big = rng.random(m) < 0.05
    k[big] = rng.integers(65, 100000, big.sum())
    k[:64] = np.arange(64) + 1
    k[100:120] = [4, 5, 8, 61, 62, 63, 64, 65, 66, 127, 128, 129, 512, 1000, 1024, 1025, 8191, 8192, 8193, 8194]
    k = k.astype(np.int32)
    row_ptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    col = np.concatenate(rows).astype(np.int32)
    l = rng.uniform(1e-6, 1e-2, n)
    mu = rng.gamma(0.3, 100.0, n)
    mu[0:3] = 0.0            # rows 10, 11: all-zero; row 12: zeros in front
    mu[3990:] = 1e-300
    return row_ptr, col, k, l, mu


@pytest.mark.parametrize("threads", ["1", "3", "8"])
@pytest.mark.parametrize("cid_base", [0, 7, (1 << 32) * 5 + 3])
def test_plan_replay_equals_canonical_replay(monkeypatch, threads, cid_base):
    monkeypatch.setenv("MMQ_PLAN_THREADS", threads)
    row_ptr, col, k, l, mu = _problem()
    P = orc.Problem(row_ptr, col, k, l)
    for sweep in (0, 3):
        _, c_o, _ = P.sweep_replay(mu, 4321, sweep, class_id_base=cid_base, do_gamma=False)
        c_p, st = P.cls_plan_replay(mu, 4321, sweep, class_id_base=cid_base)
        # three sets: k <= 64 (one slot each), k > 64 with <= 64 members (chain set), more than 64 members (rest)
        assert st["in_use"] == 1 and st["small_classes"] > 4000 and st["rest_classes"] >= 20 and st["chain_classes"] > 150
        assert st["class_slots"] % 32 == 0 and st["class_slots"] >= st["small_classes"] and st["chain_slots"] % 32 == 0
        assert c_p.sum() == k.sum()
        assert np.array_equal(c_p, c_o)


def test_plan_with_explicit_class_ids_and_permuted_rows():
    """The host program hands classes over in any order with mmq_problem.class_id: the plan keeps the ids."""
    row_ptr, col, k, l, mu = _problem(seed=9)
    m = len(k)
    P = orc.Problem(row_ptr, col, k, l)
    _, c_o, _ = P.sweep_replay(mu, 77, 2, do_gamma=False)
    perm = np.random.default_rng(1).permutation(m)
    d = np.diff(row_ptr)
    rp2 = np.concatenate([[0], np.cumsum(d[perm])]).astype(np.int64)
    col2 = np.concatenate([col[row_ptr[i]:row_ptr[i + 1]] for i in perm]).astype(np.int32)
    P2 = orc.Problem(rp2, col2, k[perm], l)
    c_p, st = P2.cls_plan_replay(mu, 77, 2, class_id=perm.astype(np.int64))
    assert st["in_use"] == 1 and np.array_equal(c_p, c_o)


def test_plan_declines_class_ids_spread_over_several_2_32_blocks():
    row_ptr, col, k, l, mu = _problem(seed=3)
    cid = np.arange(len(k), dtype=np.int64)
    cid[5] += 1 << 33
    c_p, st = orc.Problem(row_ptr, col, k, l).cls_plan_replay(mu, 1, 0, class_id=cid)
    assert c_p is None and st is None


def test_plan_on_synthetic_sample(small_problem):
    h = small_problem
    P = orc.Problem(h.row_ptr, h.col, h.k, h.len)
    mu, _, _ = P.init_mu()
    _, c_o, _ = P.sweep_replay(mu, 1234, 5, do_gamma=False)
    c_p, st = P.cls_plan_replay(mu, 1234, 5)
    assert st["in_use"] == 1 and np.array_equal(c_p, c_o) and c_p.sum() == h.N

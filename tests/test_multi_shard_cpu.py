"""N > 1 host logic on CPU (gloo, world_size 2): hit classes are partitioned into contiguous row
blocks with class_id_base, each rank allocates its block, ONE all-reduce of the int32 count vector
per sweep, then every rank draws the same Gamma stream — the result must equal the unsharded
chain bit for bit (integer sums are order-independent; src/mmseq.cpp:895-899 is the only coupling).
The device path makes the same calls (tests/test_gpu_multi.py); here the oracle stands in for the
kernels so that the sharding, the collective and the replicated Gamma step are covered without a GPU."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmseq_b200 import hostlib, synth
from oracle import oracle as orc


def shard_rows(row_ptr, world):
    """Contiguous class blocks balanced by CSR entries (the rule of mmseq_main.cpp)."""
    m = len(row_ptr) - 1
    nnz = int(row_ptr[-1])
    cuts = [0]
    r = 0
    for g in range(world):
        target = nnz * (g + 1) // world
        while r < m and row_ptr[r + 1] <= target:
            r += 1
        if g == world - 1:
            r = m
        cuts.append(r)
    return cuts


def _worker(rank, world, port, layout, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = synth.Synth(20260101 + 1, 300, 20000)
    h = hostlib.from_records(s.T, s.efflen, s.frag_ptr, s.frag_tid, layout=layout)
    P = orc.Problem(h.row_ptr, h.col, h.k, h.len)
    cuts = shard_rows(h.row_ptr, world)
    a, b = cuts[rank], cuts[rank + 1]
    mu, _, _ = P.init_mu()
    trace = []
    for sweep in range(6):
        _, c, _ = P.sweep_replay(mu, 1234, sweep, class_id_base=a, do_gamma=False, rows=(a, b))
        t = torch.from_numpy(c.astype(np.int32))
        dist.all_reduce(t)                                   # the per-sweep collective
        mu = P.gamma_replay(t.numpy(), 1234, sweep)          # replicated: same stream on every rank
        trace.append((t.numpy().copy(), mu.copy()))
    out_q.put((rank, (a, b), trace))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("layout", [hostlib.LAYOUT_COLLAPSED, hostlib.LAYOUT_PER_FRAGMENT_BY_LENGTH])
def test_sharded_chain_equals_unsharded(layout):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + layout
    procs = [ctx.Process(target=_worker, args=(r, 2, port, layout, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    s = synth.Synth(20260101 + 1, 300, 20000)
    h = hostlib.from_records(s.T, s.efflen, s.frag_ptr, s.frag_tid, layout=layout)
    P = orc.Problem(h.row_ptr, h.col, h.k, h.len)
    (a0, b0), (a1, b1) = res[0][1], res[1][1]
    assert a0 == 0 and b0 == a1 and b1 == h.m and 0 < b0 < h.m
    nnz0 = h.row_ptr[b0]
    assert abs(nnz0 - h.nnz / 2) <= np.diff(h.row_ptr).max()      # balanced by CSR entries
    mu, _, _ = P.init_mu()
    for sweep in range(6):
        _, c, mu = P.sweep_replay(mu, 1234, sweep)
        for r in res:
            assert np.array_equal(r[2][sweep][0], c) and np.array_equal(r[2][sweep][1], mu)

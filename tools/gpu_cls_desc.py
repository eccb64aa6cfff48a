"""A/B of the chunk-descriptor variant of k_alloc_cls in one process (MMQ_CLS_DESC is read per launch)."""
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from mmseq_b200 import capi, hostlib, synth
s = synth.Synth(20260103, 180000, 30000000)
h = hostlib.from_records(s.T, s.efflen, s.frag_ptr, s.frag_tid, layout=hostlib.LAYOUT_COLLAPSED | hostlib.LAYOUT_HEADER_ORDER_COLUMNS)
length = s.efflen[h.col2hdr] * 30000000 / 1e9
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
H = capi.Handle(h.row_ptr, h.col, h.k, length)
H.init_mu()
mu0 = H.get_mu()
res = {}
for rep in range(3):
    for desc in ("0", "1"):
        os.environ["MMQ_CLS_DESC"] = desc
        H.set_mu(mu0)
        H.gibbs(1234, 0, 16, stride=16, trace_len=4)
        H.synchronize(); H.kernel_times()
        for st in range(4):
            flush.zero_(); torch.cuda.synchronize()
            H.gibbs(1234, 16 * (st + 1), 16, stride=16, trace_len=8, flags=capi.MMQ_GIBBS_TIME_KERNELS)
        a, an, g, gn = H.kernel_times()
        res.setdefault(desc, []).append(a / an)
        mu = H.get_mu()
        res.setdefault("mu" + desc, []).append(mu)
print("alloc ms per sweep: run table", [round(x, 4) for x in res["0"]], " descriptors", [round(x, 4) for x in res["1"]])
print("same chain:", all(np.array_equal(a, b) for a, b in zip(res["mu0"], res["mu1"])))

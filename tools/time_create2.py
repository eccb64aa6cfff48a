"""Phase timing of mmq_create / e2e pieces for the two layouts (MMQ_CREATE_TIMING=1)."""
import time, numpy as np, torch, sys, os
sys.path.insert(0, '.')
from mmseq_b200 import capi, hostlib, synth
s = synth.Synth(20260103, 180000, 30000000)
pin = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
for name, layout in (("perfragment", hostlib.LAYOUT_PER_FRAGMENT_BY_LENGTH), ("collapsed", hostlib.LAYOUT_COLLAPSED)):
    h = hostlib.from_records(s.T, s.efflen, s.frag_ptr, s.frag_tid, layout=layout | hostlib.LAYOUT_HEADER_ORDER_COLUMNS)
    rp, col, kk, ln = pin(h.row_ptr), pin(h.col), pin(h.k), pin(h.len)
    mu0 = pin(np.full(h.n, 1.0))
    torch.cuda.synchronize()
    for it in range(3):
        print(f"--- {name} iteration {it}", file=sys.stderr, flush=True)
        t0 = time.perf_counter(); H = capi.Handle(rp, col, kk, ln); t1 = time.perf_counter()
        H.set_mu(mu0); t2 = time.perf_counter()
        H.gibbs(1, 0, 320, stride=16, trace_len=20); H.synchronize(); t3 = time.perf_counter()
        mu = H.get_mu(); tr = H.get_trace(); t4 = time.perf_counter(); H.close(); t5 = time.perf_counter()
        print(name, "create %.1f ms  set_mu %.1f ms  320 sweeps %.1f ms  get_mu+trace %.1f ms close %.1f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t5-t4)*1e3), flush=True)

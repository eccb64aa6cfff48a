"""Row plan (k_alloc_rows) against the round-1 segment kernel (k_alloc_seg4) on the per-fragment C2 sample, with and
without per-hit weights: device time of the allocation launch per sweep."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmseq_b200 import capi, hostlib, synth

for weights in (False, True):
    s = synth.Synth(20260101 + 2, 180000, 30000000, weights=weights)
    h = hostlib.from_records(s.T, s.efflen, s.frag_ptr, s.frag_tid, frag_w=s.frag_w if weights else None,
                             layout=hostlib.LAYOUT_PER_FRAGMENT_BY_LENGTH | hostlib.LAYOUT_HEADER_ORDER_COLUMNS)
    length = s.efflen[h.col2hdr] * 30000000 / 1e9
    H = capi.Handle(h.row_ptr, h.col, None, length, weight=h.w, device=0)
    H.init_mu()
    mu0 = H.get_mu()
    st = H.rows_stats()
    for name, flags in (("rows", capi.MMQ_GIBBS_ROWS_KERNEL), ("seg4", capi.MMQ_GIBBS_DEFAULT)):
        H.set_mu(mu0)
        H.gibbs(1234, 0, 16, stride=16, trace_len=8, flags=flags | capi.MMQ_GIBBS_NO_GRAPH)
        H.kernel_times()
        H.gibbs(1234, 16, 64, stride=16, trace_len=8, flags=flags | capi.MMQ_GIBBS_TIME_KERNELS)
        a, an, g, gn = H.kernel_times()
        print(f"weights={weights} {name}: alloc {1000 * a / an:.1f} us per sweep, gamma {1000 * g / gn:.1f} us; plan bytes {st['bytes_per_sweep'] / 1e6:.0f} MB", flush=True)
    H.close()

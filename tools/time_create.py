import time, numpy as np, torch, sys
sys.path.insert(0,'.')
from mmseq_b200 import capi, hostlib, synth
s=synth.Synth(20260103,180000,30000000)
h=hostlib.from_records(s.T,s.efflen,s.frag_ptr,s.frag_tid,layout=hostlib.LAYOUT_PER_FRAGMENT_BY_LENGTH|hostlib.LAYOUT_HEADER_ORDER_COLUMNS)
pin=lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
rp,col,ln=pin(h.row_ptr),pin(h.col),pin(h.len)
torch.cuda.synchronize()
for it in range(3):
    t0=time.perf_counter(); H=capi.Handle(rp,col,None,ln); t1=time.perf_counter()
    H.init_mu(); t2=time.perf_counter()
    H.gibbs(1,0,16,stride=16,trace_len=2); H.synchronize(); t3=time.perf_counter()
    tr=H.get_trace(); t4=time.perf_counter(); H.close(); t5=time.perf_counter()
    print("create %.1f ms  init_mu(+transpose) %.1f ms  16 sweeps %.1f ms  get_trace %.1f ms close %.1f ms"%((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3,(t4-t3)*1e3,(t5-t4)*1e3))

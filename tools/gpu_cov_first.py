"""First contact / timing of the mmcollapse covariance kernels (run on the GPU box)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from mmseq_b200 import capi
from oracle import oracle as orc
from tests.collapse_case import make_traces

def corr_err(R, ref):
    d = np.sqrt(np.diag(ref)); d = np.where(d > 0, d, 1.0)
    return np.max(np.abs(R - ref) / (d[:, None] * d[None, :]))

for L, C in ((64, 5), (1024, 128), (1024, 300)):
    M = make_traces(L, C, seed=L + C)
    ref = orc.trace_cov(M)
    for ns in (1, 2, 3):
        R = capi.trace_cov(M, nsplit=ns)
        print(f"L={L} C={C} nsplit={ns} corr_err={corr_err(R, ref):.3e} sym={np.array_equal(R, R.T)} "
              f"diag={np.max(np.abs(np.diag(R)/np.diag(ref)-1)):.2e}", flush=True)

dev = torch.device("cuda:0")
L = 1024
for C in (4096, 8192, 16384, 24576):
    g = torch.Generator(device=dev); g.manual_seed(1)
    Md = torch.exp(torch.randn((C, L), dtype=torch.float64, device=dev, generator=g))
    Rd = torch.empty((C, C), dtype=torch.float64, device=dev)
    for nsplit in (1, 2, 3):
        ws = torch.empty(capi.trace_cov_workspace_bytes(L, C, nsplit), dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream()
        for _ in range(2):
            capi.trace_cov_dev(Md.data_ptr(), L, C, nsplit, Rd.data_ptr(), ws.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 5
        e0.record()
        for _ in range(K):
            capi.trace_cov_dev(Md.data_ptr(), L, C, nsplit, Rd.data_ptr(), ws.data_ptr(), st.cuda_stream)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        nterms = nsplit * (nsplit + 1) // 2
        fl = 1.0 * C * (C + 1) * L * nterms      # algorithmic: the upper triangle, one multiply-add per split product
        print(f"C={C} nsplit={nsplit} {ms:.3f} ms  tensor {fl/ms/1e9:.1f} TFLOP/s  out {C*C*8/ms/1e6:.0f} GB/s", flush=True)
    if C <= 8192:
        ref = torch.cov(Md)  # rows = variables
        d = torch.sqrt(torch.diag(ref))
        err = ((Rd - ref).abs() / (d[:, None] * d[None, :])).max().item()
        print(f"C={C} nsplit=3 corr_err vs torch.cov fp64 = {err:.3e}", flush=True)
    del Md, Rd, ws

#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
python - <<'PY'
from mmseq_b200 import synth
synth.Synth(20260101 + 2, 180000, 30000000).write_hits_fast("/tmp/c2.bin.hits", True)
PY
nproc
( MMQ_TIMING=1 MMQ_LOADER_TIMING=1 mmseq_b200/bin/mmseq /tmp/c2.bin.hits /tmp/c2_tr > /dev/null ) 2> $O/r02_cli_c2_timing.txt; cat $O/r02_cli_c2_timing.txt
ls -la /tmp/c2_tr*
python - <<'PY'
import gzip
with gzip.open("/tmp/c2_tr.trace_gibbs.gz", "rt") as f:
    head = f.readline(); l1 = f.readline()
print("trace header ids", len(head.split()), "first line values", len(l1.split()), l1[:80])
PY
timeout 900 python -m pytest tests/test_gpu_cli.py -x -q 2>&1 | tail -3

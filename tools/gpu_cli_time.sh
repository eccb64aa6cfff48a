#!/bin/bash
mkdir -p gpurun_out
python - <<'PY'
from mmseq_b200 import synth
synth.Synth(20260101 + 1, 1000, 100000).write_hits_fast("/tmp/c1.bin.hits", True)
synth.Synth(20260101 + 3, 20000, 3000000).write_hits_fast("/tmp/mid.bin.hits", True)
PY
MMQ_TIMING=1 mmseq_b200/bin/mmseq /tmp/c1.bin.hits /tmp/c1_ours 2>&1 >/dev/null | grep -E "timing|Gibbs:|EM:"
MMQ_TIMING=1 mmseq_b200/bin/mmseq /tmp/mid.bin.hits /tmp/mid_ours 2>&1 >/dev/null | grep -E "timing|Gibbs:|EM:"
ls -la /tmp/mid_ours*
python - <<'PY'
import gzip, numpy as np
ids=None
with gzip.open("/tmp/mid_ours.trace_gibbs.gz","rt") as f:
    lines=f.read().split("\n")
print("trace lines", len(lines), "ids", len(lines[0].split(" "))-1, "cols in line 1", len(lines[1].split(" "))-1, "last nonempty", len(lines[-2].split(" "))-1)
PY
timeout 600 python -m pytest tests/test_gpu_cli.py -x -q -m gpu 2>&1 | tail -2

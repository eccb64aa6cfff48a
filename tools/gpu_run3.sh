#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
# config 1 (C1): the reference's own CPU-runnable case, default flags, through the host program
python - <<'PY'
from mmseq_b200 import synth
s = synth.Synth(20260101 + 1, 1000, 100000)
synth.write_hits_text(s, "/tmp/c1.hits")
synth.write_hits_binary(s, "/tmp/c1.bin.hits")
PY
( time ./mmseq_b200/bin/mmseq /tmp/c1.bin.hits /tmp/c1_out ) > gpurun_out/cli_c1.log 2>&1
tail -12 gpurun_out/cli_c1.log
head -3 /tmp/c1_out.mmseq; head -3 /tmp/c1_out.gene.mmseq

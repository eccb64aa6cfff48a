#!/bin/bash
# the host program end to end on config 1 and on a config-2 sized hits file (phase clock on stderr)
mkdir -p gpurun_out
python - <<'PY'
import time
from mmseq_b200 import synth
t=time.time(); synth.Synth(20260101 + 1, 1000, 100000).write_hits_fast("/tmp/c1.bin.hits", True); print("c1 written", round(time.time()-t,1))
t=time.time(); synth.Synth(20260101 + 2, 180000, 30000000).write_hits_fast("/tmp/c2.bin.hits", True); print("c2 written", round(time.time()-t,1))
PY
ls -la /tmp/c1.bin.hits /tmp/c2.bin.hits
( s=$(date +%s.%N); MMQ_TIMING=1 mmseq_b200/bin/mmseq /tmp/c1.bin.hits /tmp/c1_ours > /dev/null 2>/tmp/t1.txt; e=$(date +%s.%N); cat /tmp/t1.txt | tail -25; echo "C1 wall $(python -c "print(round($e - $s, 2))") s" ) 2>&1 | tee gpurun_out/cli_c1_timing.txt
( s=$(date +%s.%N); MMQ_TIMING=1 timeout 900 mmseq_b200/bin/mmseq -notraces /tmp/c2.bin.hits /tmp/c2_ours > /tmp/o2.txt 2>/tmp/t2.txt; e=$(date +%s.%N); cat /tmp/t2.txt | tail -30; echo "C2 -notraces wall $(python -c "print(round($e - $s, 2))") s" ) 2>&1 | tee gpurun_out/cli_c2_timing.txt
head -3 /tmp/c2_ours.mmseq | cut -c1-200
wc -l /tmp/c2_ours.mmseq /tmp/c2_ours.gene.mmseq

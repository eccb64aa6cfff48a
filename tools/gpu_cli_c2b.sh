#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cli.py -x -q -m gpu --timeout 300 > gpurun_out/cli_tests.log 2>&1; echo "cli tests rc=$?"; tail -2 gpurun_out/cli_tests.log
python - <<'PY'
from mmseq_b200 import synth
synth.Synth(20260101 + 1, 1000, 100000).write_hits_fast("/tmp/c1.bin.hits", True)
synth.Synth(20260101 + 2, 180000, 30000000).write_hits_fast("/tmp/c2.bin.hits", True)
PY
( s=$(date +%s.%N); MMQ_TIMING=1 mmseq_b200/bin/mmseq /tmp/c1.bin.hits /tmp/c1_ours > /dev/null 2>/tmp/t1.txt; e=$(date +%s.%N); grep -v "^Counting" /tmp/t1.txt | tail -25; echo "C1 wall $(python -c "print(round($e - $s, 2))") s" ) 2>&1 | tee gpurun_out/cli_c1_timing.txt
( s=$(date +%s.%N); MMQ_TIMING=1 timeout 900 mmseq_b200/bin/mmseq -notraces /tmp/c2.bin.hits /tmp/c2_ours > /tmp/o2.txt 2>/tmp/t2.txt; e=$(date +%s.%N); grep -v "^Counting" /tmp/t2.txt | tail -30; echo "C2 -notraces wall $(python -c "print(round($e - $s, 2))") s" ) 2>&1 | tee gpurun_out/cli_c2_timing.txt
( s=$(date +%s.%N); MMQ_TIMING=1 timeout 900 mmseq_b200/bin/mmseq /tmp/c2.bin.hits /tmp/c2_full > /tmp/o3.txt 2>/tmp/t3.txt; e=$(date +%s.%N); grep -v "^Counting" /tmp/t3.txt | tail -30; echo "C2 with traces wall $(python -c "print(round($e - $s, 2))") s" ) 2>&1 | tee gpurun_out/cli_c2_full_timing.txt
ls -la /tmp/c2_full.* | head -12

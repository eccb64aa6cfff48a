#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
MMQ_CREATE_TIMING=1 python tools/time_create2.py 2>&1 | grep -A16 "collapsed iteration 2"
python bench.py --layout collapsed --no-cpu-baseline > gpurun_out/bench_collapsed.json 2>gpurun_out/q.err || tail -3 gpurun_out/q.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_collapsed.json")); r=d["roofline"]
print("collapsed | sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", round(r["avg_launch_ms"],4), "e2e", round(d["e2e"]["sweeps_per_s"],1), d["e2e"]["wall_s"])
PY
python - <<'PY'
from mmseq_b200 import synth
synth.Synth(20260101 + 2, 180000, 30000000).write_hits_fast("/tmp/c2.bin.hits", True)
PY
( s=$(date +%s.%N); MMQ_TIMING=1 timeout 900 mmseq_b200/bin/mmseq -notraces /tmp/c2.bin.hits /tmp/c2_ours > /tmp/o2.txt 2>/tmp/t2.txt; e=$(date +%s.%N); grep -v "^Counting" /tmp/t2.txt | tail -30; echo "C2 -notraces wall $(python -c "print(round($e - $s, 2))") s" ) 2>&1 | tee gpurun_out/cli_c2_timing.txt

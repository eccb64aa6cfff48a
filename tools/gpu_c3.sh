#!/bin/bash
# config 3 (haplotype-specific transcriptome) at full size, both layouts; host program on the C2 file with traces
mkdir -p gpurun_out
for L in perfragment collapsed; do
  timeout 900 python bench.py --haplo --layout $L --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_$L.json 2>gpurun_out/bench_c3_$L.err || tail -3 gpurun_out/bench_c3_$L.err
  python - $L <<'PY'
import json,sys
L=sys.argv[1]
d=json.loads(open(f"gpurun_out/bench_c3_{L}.json").read().strip().split("\n")[-1]); r=d["roofline"]; c=d["config"]
print("C3",L,"n",c["n_columns"],"m",c["classes_per_gpu"],"nnz",c["nnz_per_gpu"],"| sweeps/s",round(d["sweeps_per_s"],1),"alloc_ms",round(r["avg_launch_ms"],4),"gamma_ms",round(r["gamma_avg_launch_ms"],4),"frac",round(r["frac"],4),"e2e",round(d["e2e"]["sweeps_per_s"],1), r.get("class_plan"))
PY
done
python - <<'PY'
from mmseq_b200 import synth
synth.Synth(20260101 + 2, 180000, 30000000).write_hits_fast("/tmp/c2.bin.hits", True)
PY
( s=$(date +%s.%N); MMQ_TIMING=1 timeout 900 mmseq_b200/bin/mmseq /tmp/c2.bin.hits /tmp/c2_full > /tmp/o3.txt 2>/tmp/t3.txt; e=$(date +%s.%N); grep -v "^Counting" /tmp/t3.txt | tail -30; echo "C2 with traces wall $(python -c "print(round($e - $s, 2))") s" ) 2>&1 | tee gpurun_out/cli_c2_full_timing.txt
( s=$(date +%s.%N); MMQ_TIMING=1 timeout 900 mmseq_b200/bin/mmseq -notraces /tmp/c2.bin.hits /tmp/c2_ours > /tmp/o2.txt 2>/tmp/t2.txt; e=$(date +%s.%N); grep -v "^Counting" /tmp/t2.txt | tail -30; echo "C2 -notraces wall $(python -c "print(round($e - $s, 2))") s" ) 2>&1 | tee gpurun_out/cli_c2_timing.txt

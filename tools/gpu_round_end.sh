#!/bin/bash
# what the driver does at round end, plus the profiles we commit
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
timeout 900 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 900 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
kill $SMI
timeout 600 python bench.py --weights --no-cpu-baseline > gpurun_out/bench_weighted.json 2>> gpurun_out/bench_ours.err
timeout 600 python bench.py --layout collapsed --no-cpu-baseline > gpurun_out/bench_collapsed.json 2>> gpurun_out/bench_ours.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 200 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-collapsed > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_alloc_seg4 -s 5 -c 1 -o gpurun_out/prof_r01_k_alloc_seg4 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-collapsed > gpurun_out/ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_alloc_cls -s 10 -c 2 -o gpurun_out/prof_r01_k_alloc_cls -f python bench.py --layout collapsed --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_cls.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 200 --csv --log-file gpurun_out/launches_r01_collapsed.csv python bench.py --layout collapsed --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
# config 1 end to end: the reference's own main() (unmodified sources + oracle/shim) on the host cores vs the host program on the GPU
python - <<'PY'
from mmseq_b200 import synth
synth.Synth(20260101 + 1, 1000, 100000).write_hits_fast("/tmp/c1.bin.hits", True)
PY
( s=$(date +%s.%N); OMP_NUM_THREADS=$(nproc) oracle/_ref/mmseq_ref /tmp/c1.bin.hits /tmp/c1_ref > /dev/null 2>&1; e=$(date +%s.%N); echo "reference mmseq_ref (unmodified sources + oracle/shim) wall $(python -c "print(round($e - $s, 2))") s on $(nproc) threads" ) > gpurun_out/c1_compare.txt 2>&1
( s=$(date +%s.%N); mmseq_b200/bin/mmseq /tmp/c1.bin.hits /tmp/c1_ours > /dev/null 2>/tmp/t_ours.txt; e=$(date +%s.%N); grep -E "Gibbs:|EM:" /tmp/t_ours.txt; echo "mmseq_b200 mmseq wall $(python -c "print(round($e - $s, 2))") s on 1 GPU" ) >> gpurun_out/c1_compare.txt 2>&1
cat gpurun_out/c1_compare.txt
python - <<'PY'
import json
for f in ("reference","ours","weighted","collapsed"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().split("\n")[-1])
        r=d.get("roofline") or {}
        print(f, "value %.4g"%d["value"], "sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", r.get("avg_launch_ms"), "frac", r.get("frac"), "e2e", d["e2e"] and round(d["e2e"].get("sweeps_per_s",0),1), "cpu", d.get("cpu_baseline",{}).get("sweeps_per_s"), "clocks", d.get("clocks"))
    except Exception as e: print(f,"failed",e)
PY

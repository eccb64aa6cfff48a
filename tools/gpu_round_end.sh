#!/bin/bash
# What the driver does at round end on one GPU, plus the evidence committed under profiles/ (r02_*).
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/r02_clocks_during_bench.csv &
SMI=$!
timeout 900 python bench.py --impl reference > $O/r02_bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"
timeout 900 python bench.py > $O/r02_bench_ours.json 2> $O/bench_ours.err; echo "ours rc=$?"; tail -2 $O/bench_ours.err
kill $SMI
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-gates --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_alloc|k_gamma|k_trace|k_add2|k_set2" -c 600 --csv --log-file $O/r02_launches.csv $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_alloc_chain|k_alloc_cls|k_gamma" -s 10 -c 5 -o $O/prof_r02_sweep -f $B > $O/ncu.log 2>&1; tail -1 $O/ncu.log
ncu --set full --clock-control none --import-source on -k regex:"k_alloc_seg4" -s 4 -c 1 -o $O/prof_r02_seg4w -f python bench.py --weights --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-gates --no-extras > $O/ncu_seg4w.log 2>&1; tail -1 $O/ncu_seg4w.log
ncu --set full --clock-control none --import-source on -k regex:k_cov_gemm -s 1 -c 1 -o $O/prof_r02_cov -f python tools/gpu_cov_prof.py > $O/ncu_cov.log 2>&1; tail -1 $O/ncu_cov.log
timeout 300 python tools/gpu_cov_first.py > $O/r02_cov_sizes.txt 2>&1; tail -4 $O/r02_cov_sizes.txt
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_collapse.py -q -x -k "not full_size" > $O/r02_memcheck_cov.log 2>&1; echo "memcheck cov rc=$?"; tail -3 $O/r02_memcheck_cov.log
# config 5 on one GPU: 4 samples, the reference's default chain length
timeout 900 python bench.py --batch 4 --batch-per-gpu 4 > $O/r02_bench_batch_1gpu.json 2> $O/bench_batch.err; echo "batch rc=$?"
# config 1 end to end: the reference's own main() (unmodified sources + oracle/shim) on the host cores vs the host program on the GPU
python - <<'PY'
from mmseq_b200 import synth
synth.Synth(20260101 + 1, 1000, 100000).write_hits_fast("/tmp/c1.bin.hits", True)
synth.Synth(20260101 + 2, 180000, 30000000).write_hits_fast("/tmp/c2.bin.hits", True)
PY
( s=$(date +%s.%N); OMP_NUM_THREADS=$(nproc) oracle/_ref/mmseq_ref /tmp/c1.bin.hits /tmp/c1_ref > /dev/null 2>&1; e=$(date +%s.%N); echo "reference mmseq_ref (unmodified sources + oracle/shim) wall $(python -c "print(round($e - $s, 2))") s on $(nproc) threads" ) > $O/r02_c1_compare.txt 2>&1
( s=$(date +%s.%N); mmseq_b200/bin/mmseq /tmp/c1.bin.hits /tmp/c1_ours > /dev/null 2>/tmp/t_ours.txt; e=$(date +%s.%N); echo "mmseq_b200 mmseq wall $(python -c "print(round($e - $s, 2))") s on 1 GPU" ) >> $O/r02_c1_compare.txt 2>&1
python - >> $O/r02_c1_compare.txt 2>&1 <<'PY'
# the two programs' outputs on config 1: .k / .M byte for byte, deterministic columns equal, log_mu within Monte-Carlo error
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import tables
from tests.test_reference_run import compare_with_reference
for ext in (".k", ".M"):
    print(ext, "byte-identical:", open("/tmp/c1_ref" + ext).read() == open("/tmp/c1_ours" + ext).read())
print(".identical.mmseq (header only: the file declares no identical sets) byte-identical:", open("/tmp/c1_ref.identical.mmseq").read() == open("/tmp/c1_ours.identical.mmseq").read())
for ext, kind in ((".mmseq", "mmseq"), (".gene.mmseq", "gene")):
    z = compare_with_reference(tables.read_table("/tmp/c1_ref" + ext), tables.read_table("/tmp/c1_ours" + ext), kind, z_max=6.0, frac=0.95)
    print(ext, "deterministic columns equal;", len(z), "observed features, |z| of log_mu: median %.2f, 99th percentile %.2f, max %.2f" % (np.median(z), np.percentile(z, 99), z.max()))
PY
cat $O/r02_c1_compare.txt
( MMQ_TIMING=1 MMQ_LOADER_TIMING=1 mmseq_b200/bin/mmseq -notraces /tmp/c2.bin.hits /tmp/c2_ours > /dev/null ) 2> $O/r02_cli_c2_notraces_timing.txt; tail -12 $O/r02_cli_c2_notraces_timing.txt
( MMQ_TIMING=1 mmseq_b200/bin/mmseq /tmp/c2.bin.hits /tmp/c2_tr > /dev/null ) 2> $O/r02_cli_c2_timing.txt; grep "timing" $O/r02_cli_c2_timing.txt
python - <<'PY'
import json
for f in ("reference","ours"):
    try:
        d=json.loads(open(f"gpurun_out/r02_bench_{f}.json").read().strip().split("\n")[-1])
        r=d.get("roofline") or {}
        print(f, "value %.4g"%d["value"], "sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", r.get("avg_launch_ms"), "gamma_ms", r.get("gamma_avg_launch_ms"), "frac", r.get("frac"), "e2e", d["e2e"] and round(d["e2e"].get("sweeps_per_s",0),1), "cpu", d.get("cpu_baseline",{}).get("sweeps_per_s"), "clocks", d.get("clocks"))
        print("  gates", json.dumps(d.get("gates")))
        print("  weighted", json.dumps(d.get("perfragment_weighted")))
        print("  trace_cov", json.dumps(d.get("trace_cov"))[:900])
    except Exception as e: print(f,"failed",e)
try:
    b=json.loads(open("gpurun_out/r02_bench_batch_1gpu.json").read().strip().split("\n")[-1]); print("batch", {k:b[k] for k in ("value","wall_s","sweeps_per_s_per_gpu","gpu_pipeline_s_median","prep_s_median")})
except Exception as e: print("batch failed", e)
PY

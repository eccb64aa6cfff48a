#!/bin/bash
# launch list of one short bench run + full captures of the sweep's kernels (collapsed C2)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-gates --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_alloc|k_gamma|k_trace|k_add2|k_set2" -c 600 --csv --log-file gpurun_out/launches_r02.csv $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_alloc_chain|k_alloc_cls|k_gamma" -s 8 -c 4 -o gpurun_out/prof_r02_sweep -f $B > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches_r02.csv")) if len(r)>5]
hdr=rows[0]; iN=hdr.index("Kernel Name"); iV=hdr.index("Metric Value")
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[iN][:60]].append(float(r[iV].replace(",","")))
    except: pass
for k,v in sorted(agg.items(), key=lambda x:-sum(x[1])):
    print(f"{k:60s} n={len(v):4d} avg_us={sum(v)/len(v)/1000:9.2f} tot_ms={sum(v)/1e6:8.3f}")
PY

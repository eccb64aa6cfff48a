"""Summarise an exported ncu report (all captured launches):
   ncu -i X.ncu-rep --page raw --csv > raw.csv ; ncu -i X.ncu-rep --page source --csv > sass.csv
   python tools/ncu_summary.py raw.csv [sass.csv [top_n]]"""
import csv, sys
from collections import Counter
WANT = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__warps_eligible.avg.per_cycle_active','smsp__warps_active.avg.per_cycle_active','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','lts__t_sectors.sum','lts__t_sectors_srcunit_tex_op_red.sum','l1tex__throughput.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__cycles_active.avg','sm__cycles_elapsed.max']
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr, units = rows[0], rows[1]
for d in rows[2:]:
    print(d[hdr.index('Kernel Name')][:70])
    for w in WANT:
        if w in hdr:
            print(' ', w, d[hdr.index(w)], units[hdr.index(w)])
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            v = float(d[i])
            if v > 0.3:
                print('  stall', h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), round(v, 2))
if len(sys.argv) > 2:
    rows = list(csv.reader(open(sys.argv[2])))
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
    seen = set()
    for s in starts:
        name = rows[s - 1][1][:70] if s > 0 and len(rows[s - 1]) > 1 else ''
        if name in seen:
            continue
        seen.add(name)
        hdr = rows[s]
        iS = hdr.index("Source"); iE = hdr.index("Instructions Executed"); iSamp = hdr.index("# Samples")
        data = []
        for r in rows[s + 1:]:
            if len(r) < len(hdr) or r[0] in ("Kernel Name", "Address"):
                break
            data.append(r)
        tot = sum(int(r[iE]) for r in data)
        print("sass of", name, "lines", len(data), "warp instructions", tot)
        out = [(i, int(r[iE]), int(r[iSamp]), r[iS][:64]) for i, r in enumerate(data)]
        for x in sorted(out, key=lambda x: -x[2])[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
            print('  ', x)
        c = Counter()
        for i, e, s_, src in out:
            op = src.split()[0] if not src.startswith('@') else src.split()[1]
            c[op.split('.')[0]] += e
        print([(k, round(v / tot * 100, 1)) for k, v in c.most_common(22)])

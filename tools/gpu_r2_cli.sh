#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
python - <<'PY'
from mmseq_b200 import synth
synth.Synth(20260101 + 2, 180000, 30000000).write_hits_fast("/tmp/c2.bin.hits", True)
PY
nproc
( MMQ_TIMING=1 MMQ_LOADER_TIMING=1 mmseq_b200/bin/mmseq -notraces /tmp/c2.bin.hits /tmp/c2_ours > /dev/null ) 2> $O/r02_cli_c2_notraces_timing.txt; cat $O/r02_cli_c2_notraces_timing.txt
( MMQ_LOADER_SERIAL_INFLATE=1 MMQ_TIMING=1 MMQ_LOADER_TIMING=1 mmseq_b200/bin/mmseq -notraces /tmp/c2.bin.hits /tmp/c2_serial > /dev/null ) 2> $O/r02_cli_c2_serial_loader_timing.txt; grep -E "loader|hits file loaded|tables" $O/r02_cli_c2_serial_loader_timing.txt
for e in .mmseq .k .gene.mmseq; do cmp /tmp/c2_ours$e /tmp/c2_serial$e && echo "$e identical between the parallel and the serial loader"; done
timeout 900 python -m pytest tests/test_gpu_cli.py -x -q 2>&1 | tail -3

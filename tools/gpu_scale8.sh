#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L > gpurun_out/smi_multi.txt
MMQ_P2P_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_$N.json 2>gpurun_out/scale_$N.err
grep "p2p trace" gpurun_out/scale_$N.err | head -3
python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/scale_{N}.json").read().strip().split("\n")[-1]); r=d["roofline"]
    print("N",N,"value",d["value"],"sweeps/s",round(d["sweeps_per_s"],1),"step_ms",round(d["ms_per_step"],3),"e2e",d["e2e"] and round(d["e2e"]["sweeps_per_s"],1), r.get("per_rank"))
except Exception as e: print("N",N,"failed",e, open(f"gpurun_out/scale_{N}.err").read()[-600:])
PY
timeout 300 mmseq_b200/bin/mmseq -gpus $N -gibbs_iter 2048 /tmp/nonexistent.hits /tmp/x 2>&1 | tail -1
python - <<'PY'
from mmseq_b200 import synth
synth.Synth(20260101 + 1, 1000, 100000).write_hits_fast("/tmp/c1.bin.hits", True)
PY
( s=$(date +%s.%N); timeout 300 mmseq_b200/bin/mmseq -gpus $N /tmp/c1.bin.hits /tmp/c1_n > /dev/null 2>/tmp/t_n.txt; e=$(date +%s.%N); grep -E "Gibbs:|EM:" /tmp/t_n.txt; echo "mmseq -gpus $N config 1 wall $(python -c "print(round($e - $s, 2))") s" )
( timeout 300 mmseq_b200/bin/mmseq -gpus 1 /tmp/c1.bin.hits /tmp/c1_1 > /dev/null 2>/dev/null; cmp /tmp/c1_1.mmseq /tmp/c1_n.mmseq && echo "tables identical for 1 and $N GPUs" )

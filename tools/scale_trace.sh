#!/bin/bash
# the strong-scaled headline workload on every GPU of the box: plain run, then per-stage times and the exchange timeline (LKM_STEP_TRACE=1)
N=$(python -c "import torch; print(torch.cuda.device_count())")
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "$N or device_group" 2>&1 | tail -2
for rep in 1 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$rep bench.py --gpus $N --steps 20 --warmup 3 --no-others --no-e2e --no-cpu-baseline --no-init 2>/dev/null | grep "^{" | python -c "
import json,sys
b=json.loads(sys.stdin.readline()); print('plain', b['n_gpus'], b['value'], b['ms_per_step'], b['clocks']['sm_mhz'])"
done
LKM_STEP_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus $N --steps 40 --warmup 3 --no-others --no-e2e --no-cpu-baseline --no-init 2>&1 | grep "step_trace\|^{" > gpurun_out/scale_trace_${N}gpu.txt
grep "step_trace" gpurun_out/scale_trace_${N}gpu.txt | grep -v finalize_cl | awk '{k=$2; s[k]+=$6; c[k]++; if ($6>m[k]) m[k]=$6} END {for (k in s) print k, "mean", s[k]/c[k], "max", m[k], "n", c[k]}'
grep "finalize_cl" gpurun_out/scale_trace_${N}gpu.txt | sed 's/step_trace finalize_cl (last iteration, CTA 0, ns from its start): //' | awk '{print $2, $4, $8, $10, $15, $17}' | sort -n | awk 'NR%60==1'

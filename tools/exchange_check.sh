#!/bin/bash
# self-validating exchange of finalize_cl_kernel on N GPUs: multi-GPU tests, 1.25M-row shards per GPU with and without it, exchange timeline
N=$(python -c "import torch; print(torch.cuda.device_count())")
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -3
F="--gpus $N --steps 40 --warmup 3 --no-others --no-e2e --no-cpu-baseline --no-init --rows $((1250000 * N))"
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py $F 2>&1; }
show() { grep "^{" | python -c "
import json,sys
b=json.loads(sys.stdin.readline()); print('$1', b['n_gpus'], b['value'], b['ms_per_step'], b['clocks']['sm_mhz'])"; }
LKM_P2P_LL=1 run 29611 | show ll
LKM_P2P_LL=0 run 29612 | show flags
LKM_P2P_LL=1 run 29613 | show ll
LKM_STEP_TRACE=1 LKM_P2P_LL=1 run 29614 | grep "step_trace" > gpurun_out/r2c_trace_ll_${N}gpu.txt
grep finalize_cl gpurun_out/r2c_trace_ll_${N}gpu.txt | tail -4
grep "step_trace" gpurun_out/r2c_trace_ll_${N}gpu.txt | grep -v finalize_cl | awk '{k=$2; s[k]+=$6; c[k]++} END {for (k in s) print k, "mean", s[k]/c[k], "n", c[k]}'

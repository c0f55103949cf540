#!/bin/bash
# the headline command on every GPU of the box, as the driver launches it (other workloads and the CPU baseline skipped)
N=$(python -c "import torch; print(torch.cuda.device_count())")
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 20 --warmup 3 --no-others --no-cpu-baseline > gpurun_out/r2c_bench_cfg3_${N}gpu.json 2> gpurun_out/r2c_bench_cfg3_${N}gpu.err; echo "rc=$?"
python -c "
import json
for l in open('gpurun_out/r2c_bench_cfg3_${N}gpu.json'):
    if l.startswith('{'):
        b=json.loads(l); print(b['n_gpus'], b['value'], b['ms_per_step'], b['clocks'], b['e2e'])"

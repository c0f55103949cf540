#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "8 or device_group" > gpurun_out/r2q_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -5 gpurun_out/r2q_pytest_multi.log
N=$(python -c "import torch; print(torch.cuda.device_count())")
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 3 --no-others > gpurun_out/r2q_bench_${N}gpu.json 2> gpurun_out/r2q_bench_${N}gpu.err; echo "bench rc=$?"
python -c "
import json;b=json.load(open('gpurun_out/r2q_bench_${N}gpu.json'))
print(b['n_gpus'], b['value'], b['ms_per_step'], b['e2e'])"
tail -3 gpurun_out/r2q_bench_${N}gpu.err

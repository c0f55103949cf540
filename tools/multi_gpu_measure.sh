# Multi-GPU validation under gpurun --gpus N (N=2|4|8 via env N): tests/multi_gpu_check.py + weak-scaling bench lines of cfg2 and cfg5.
N=${N:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/r1c_multi_gpu_check_$N.log 2>&1; grep -c " ok " gpurun_out/r1c_multi_gpu_check_$N.log; grep -c "MISMATCH\|Error\|error" gpurun_out/r1c_multi_gpu_check_$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-init 2>gpurun_out/r1c_bench_cfg2_${N}gpu.err | tail -1 > gpurun_out/r1c_bench_cfg2_${N}gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-init 2>/dev/null | tail -1 > gpurun_out/r1c_bench_cfg5_${N}gpu.json
for f in cfg2_${N}gpu cfg5_${N}gpu; do python -c "
import json
j=json.load(open('gpurun_out/r1c_bench_$f.json')); print('$f', j['value'], j['ms_per_step'], j['n_gpus'], (j.get('e2e') or {}).get('value'))"; done

python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/r1c_multi_gpu_check.log 2>&1; tail -8 gpurun_out/r1c_multi_gpu_check.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-init 2>/dev/null | tail -1 > gpurun_out/r1c_bench_cfg2_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-init 2>/dev/null | tail -1 > gpurun_out/r1c_bench_cfg5_2gpu.json
for f in cfg2_2gpu cfg5_2gpu; do python -c "
import json
j=json.load(open('gpurun_out/r1c_bench_$f.json')); print('$f', j['value'], j['ms_per_step'], j['n_gpus'], (j.get('e2e') or {}).get('value'))"; done

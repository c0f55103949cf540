for i in 1 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-init --no-e2e 2>/dev/null | tail -1 > gpurun_out/r1c_bench_cfg2_2gpu.json
python -c "
import json
j=json.load(open('gpurun_out/r1c_bench_cfg2_2gpu.json')); print('cfg2 x2', j['value'], j['ms_per_step'], j['config']['l2_warm_ms_per_step'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tests/multi_gpu_check.py 2>&1 | grep -c " ok "

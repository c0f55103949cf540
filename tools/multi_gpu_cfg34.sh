for w in cfg3 cfg4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${N:-2} --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus ${N:-2} --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-init 2>/dev/null | tail -1 > gpurun_out/r1c_bench_${w}_${N:-2}gpu.json
python -c "
import json
j=json.load(open('gpurun_out/r1c_bench_${w}_${N:-2}gpu.json')); print('$w x${N:-2}', j['value'], j['ms_per_step'], j['n_gpus'])"
done

#!/bin/bash
# the driver's multi-GPU command line, with the side workloads
N=$(python -c "import torch; print(torch.cuda.device_count())")
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_full_${N}gpu.json 2> gpurun_out/r2_bench_full_${N}gpu.err ) 2>&1 | grep real
echo "rc=$?"
python - <<PY
import json
L=[l for l in open('gpurun_out/r2_bench_full_${N}gpu.json').read().splitlines() if l.startswith('{')]
b=json.loads(L[-1])
print(b['n_gpus'], b['value'], b['ms_per_step'], 'e2e', b['e2e']['value'], list(b['config'].keys()))
for k,v in b['roofline']['other_workloads'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'))
PY
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29632 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_ref_${N}gpu.json 2>/dev/null ) 2>&1 | grep real
grep -c "^{" gpurun_out/r2_bench_ref_${N}gpu.json; cut -c1-300 gpurun_out/r2_bench_ref_${N}gpu.json | grep "^{"

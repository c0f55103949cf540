#!/bin/bash
# multi-GPU check after the sorted update / pivot stage: the multi-GPU pytest (also with the sorted update forced) and the driver's bench line
N=$(python -c "import torch; print(torch.cuda.device_count())")
python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
LKM_SORT_UPD=1 LKM_ROWS_PIVOT=1 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2b_bench_full_${N}gpu.json 2> gpurun_out/r2b_bench_full_${N}gpu.err ) 2>&1 | grep real
python - <<PY
import json
L=[l for l in open('gpurun_out/r2b_bench_full_${N}gpu.json').read().splitlines() if l.startswith('{')]
b=json.loads(L[-1])
print(b['n_gpus'], b['value'], b['ms_per_step'], 'e2e', b['e2e']['value'])
for k,v in b['roofline']['other_workloads'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'))
PY

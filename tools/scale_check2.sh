#!/bin/bash
N=$(python -c "import torch; print(torch.cuda.device_count())")
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "$N or device_group" 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --workload cfg2 --steps 20 --warmup 3 --no-others --no-e2e --no-cpu-baseline --no-init 2>/dev/null | grep "^{" | python -c "
import json,sys
b=json.loads(sys.stdin.readline()); print('cfg2 weak', b['n_gpus'], b['value'], b['ms_per_step'])"
python bench.py --workload cfg2 --steps 20 --warmup 3 --no-others --no-e2e --no-cpu-baseline --no-init 2>/dev/null | grep "^{" | python -c "
import json,sys
b=json.loads(sys.stdin.readline()); print('cfg2 1 gpu', b['n_gpus'], b['value'], b['ms_per_step'])"

#!/bin/bash
N=$(python -c "import torch; print(torch.cuda.device_count())")
for cl in 1 0; do
LKM_FIN_CL=$cl python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$cl bench.py --gpus $N --rows $((1250000*N)) --steps 40 --warmup 3 --no-others --no-e2e --no-cpu-baseline --no-init 2>/dev/null | grep "^{" | python -c "
import json,sys
b=json.loads(sys.stdin.readline()); print('fin_cl=$cl', b['n_gpus'], b['value'], b['ms_per_step'], b['detail']['wall_ms_per_step'], b['clocks']['sm_mhz'])"
done

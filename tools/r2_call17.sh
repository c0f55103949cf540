#!/bin/bash
N=$(python -c "import torch; print(torch.cuda.device_count())")
LKM_STEP_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus $N --rows $((1250000*N)) --steps 40 --warmup 3 --no-others --no-e2e --no-cpu-baseline --no-init 2>&1 | grep "finalize_cl\|^{" | cut -c1-260 | tail -12

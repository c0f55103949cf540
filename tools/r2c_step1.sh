#!/bin/bash
# finalize_cl_kernel rewritten (warp-parallel source sums, self-validating exchange): tests, headline step, 1.25M-row shard step, stage trace
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -q -m gpu -x) > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c_pytest_gpu.log
F="--steps 20 --warmup 3 --no-others --no-e2e --no-cpu-baseline --no-init"
show() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        b=json.loads(l); print('$1', b['value'], b['ms_per_step'], b['clocks']['sm_mhz'], b['gpu_launches'])"; }
python bench.py $F 2>/dev/null | show full
python bench.py $F --rows 1250000 2>/dev/null | show shard1250k
python bench.py $F --rows 1250000 2>/dev/null | show shard1250k
LKM_STEP_TRACE=1 python bench.py $F --rows 1250000 2>&1 | grep step_trace | tail -12

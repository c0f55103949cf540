#!/bin/bash
# E role balanced over the fma and alu pipes: tests, cfg4 shard and headline steps
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -q -m gpu -x) > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c_pytest_gpu.log
F="--steps 20 --warmup 3 --no-others --no-e2e --no-cpu-baseline --no-init"
show() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        b=json.loads(l); print('$1', b['value'], b['ms_per_step'], b['clocks']['sm_mhz'], b['gpu_launches'], b['roofline'].get('mean_launch_us'))"; }
timeout 150 python bench.py $F --workload cfg4 2>/dev/null | show cfg4
timeout 150 python bench.py $F 2>/dev/null | show cfg3
timeout 150 python bench.py $F --shuffled 2>/dev/null | show cfg3_hashed
WL=cfg3 KPP=1 ITERS=6 timeout 150 python tools/r2_probe.py 2>&1 | tail -3

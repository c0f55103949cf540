#!/bin/bash
# A/B of library builds (LKM_LIB_PATH): cfg4 shard and headline steps
F="--steps 20 --warmup 3 --no-others --no-e2e --no-cpu-baseline --no-init"
show() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        b=json.loads(l); print('$1', b['value'], b['ms_per_step'], b['clocks']['sm_mhz'], b['gpu_launches'], b['roofline'].get('mean_launch_us'))"; }
for lib in "$@"; do
  export LKM_LIB_PATH=/root/repo/linfa_b200/$lib
  timeout 150 python bench.py $F --workload cfg4 2>/dev/null | show "$lib cfg4"
  timeout 150 python bench.py $F 2>/dev/null | show "$lib cfg3"
done

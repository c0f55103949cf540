#!/bin/bash
# cfg5 shard step, twice (the closing bench line showed 2.53 ms per step against 1.89 ms of kernel time)
F="--steps 20 --warmup 3 --no-others --no-e2e --no-cpu-baseline --no-init"
for i in 1 2; do
timeout 200 python bench.py $F --workload cfg5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        b=json.loads(l); print('cfg5', b['value'], b['ms_per_step'], b['clocks'], b['gpu_launches'], b['roofline'].get('mean_launch_us'))"
done

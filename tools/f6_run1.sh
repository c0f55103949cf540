set -x
python -m pytest tests/test_gpu_parity.py -x -q -k "f64_screened" 2>&1 | tail -15 > gpurun_out/f6_tests.log
for v in "LKM_F64S_PACKED=1" "LKM_F64S_PACKED=0" "LKM_F64S=0"; do
  env $v python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-init 2>/dev/null | tail -1 > "gpurun_out/f6_bench_${v}.json"
done
cat gpurun_out/f6_tests.log
for f in gpurun_out/f6_bench_*.json; do python -c "
import json,sys
j=json.load(open('$f')); print('$f', j['ms_per_step'], j['roofline']['frac'], j['roofline']['mean_launch_us'])"; done

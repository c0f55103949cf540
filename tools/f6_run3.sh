python -m pytest tests/test_gpu_parity.py -x -q -k "f64_screened" 2>&1 | tail -8 > gpurun_out/f6_tests.log
cat gpurun_out/f6_tests.log
python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-init 2>/dev/null | tail -1 > gpurun_out/f6_bench.json
python -c "
import json
j=json.load(open('gpurun_out/f6_bench.json')); print(j['ms_per_step'], j['roofline']['frac'], j['roofline']['mean_launch_us'], j['clocks'])"

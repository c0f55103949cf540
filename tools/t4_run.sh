python -m pytest tests/test_gpu_parity.py -x -q -k "pipeline_kernel or tc_path" 2>&1 | tail -5 > gpurun_out/t4_tests.log
cat gpurun_out/t4_tests.log
for i in 1 2; do
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-init 2>/dev/null | tail -1 > gpurun_out/t4_bench.json
python -c "
import json
j=json.load(open('gpurun_out/t4_bench.json')); print(j['ms_per_step'], j['roofline']['frac'], j['roofline']['mean_launch_us'], j['clocks'])"
done

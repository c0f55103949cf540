for v in 1 0; do
for i in 1 2; do
LKM_FLUSH_SWEEP=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-init 2>/dev/null | tail -1 > gpurun_out/sweep_$v.json
python -c "
import json
j=json.load(open('gpurun_out/sweep_$v.json')); print('SWEEP=$v', j['ms_per_step'], j['config']['l2_warm_ms_per_step'], j['roofline']['frac'], j['roofline']['mean_launch_us'])"
done
done

for r in 2 1; do
LKM_F64S_RPT=$r timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "f64_screened" 2>&1 | tail -2
for i in 1 2; do
LKM_F64S_RPT=$r python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-init 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('RPT=$r cfg5', round(d['ms_per_step'],4), round(d['roofline']['mean_launch_us'],1), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"
done
done
LKM_F64S_RPT=1 python tools/shuffled_timing_f64.py 2>&1 | grep "path 3"

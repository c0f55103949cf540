timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -k "pipeline_kernel or tc_path or golden or kmeanspp_default" 2>&1 | tail -4
for v in "LKM_FUSED_FINALIZE=1 LKM_COOP=1" "LKM_FUSED_FINALIZE=1 LKM_COOP=0" "LKM_FUSED_FINALIZE=0"; do
for i in 1 2; do
env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-init 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step']*1e3,2), round(d['config']['l2_warm_ms_per_step']*1e3,2), round(d['roofline']['mean_launch_us'],2), d['gpu_launches'])"
done
done

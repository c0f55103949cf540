for lib in liblkm.so liblkm_alt.so; do
  for i in 1 2; do
  LKM_LIB_PATH=/root/repo/linfa_b200/$lib python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-init 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', round(d['ms_per_step'],4), round(d['roofline']['mean_launch_us'],1), d['clocks']['sm_mhz'])"
  done
done

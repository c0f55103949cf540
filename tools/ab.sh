for lib in liblkm.so liblkm_alt.so; do
  export LKM_LIB_PATH=/root/repo/linfa_b200/$lib
  for w in cfg2 cfg3; do
    python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-init 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib $w', round(d['ms_per_step'],4), round(d['roofline']['mean_launch_us'],1))"
  done
done

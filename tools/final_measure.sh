# Round-1 closing measurements (run under gpurun): bench lines, launch lists, ncu captures of the dominant kernels.
set -x
python bench.py > gpurun_out/r1b_bench_cfg2.json 2> gpurun_out/r1b_bench_cfg2.err
for w in cfg3 cfg4 cfg5 cfg1; do python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-init > gpurun_out/r1b_bench_$w.json 2>/dev/null; done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1b_plain_cfg2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1b_launches_cfg2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1b_ncu_cfg2.log 2>&1
python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-init > gpurun_out/r1b_plain_cfg3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1b_launches_cfg3.csv python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-init > gpurun_out/r1b_ncu_cfg3.log 2>&1
N=2000000 python tools/cfg3_trace.py > gpurun_out/r1b_plain_cfg3s.log 2>&1 && \
N=2000000 ncu --set full --clock-control none --import-source on -k regex:'assign_tc5s|update_rows' -s 4 -c 2 -o gpurun_out/r1b_prof_cfg3 python tools/cfg3_trace.py > gpurun_out/r1b_ncu_cfg3s.log 2>&1
python tools/cfg2_run.py > gpurun_out/r1b_plain_cfg2s.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lloyd_tc4 -s 3 -c 2 -o gpurun_out/r1b_prof_cfg2 python tools/cfg2_run.py > gpurun_out/r1b_ncu_cfg2s.log 2>&1
tail -c 600 gpurun_out/r1b_bench_cfg2.json

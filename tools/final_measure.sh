# Round-1 closing measurements (run under gpurun): full GPU test suite, bench lines, launch lists, ncu captures of the dominant kernels.
set -x
P=r1c
(time python -m pytest tests -x -q -m gpu) > gpurun_out/${P}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${P}_pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/${P}_smoke.log 2>&1; tail -1 gpurun_out/${P}_smoke.log
python bench.py > gpurun_out/${P}_bench_cfg2.json 2> gpurun_out/${P}_bench_cfg2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${P}_bench_reference.json 2>/dev/null
for w in cfg3 cfg4 cfg5 cfg1; do python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-init > gpurun_out/${P}_bench_$w.json 2>/dev/null; done
# launch lists (each after its plain run has exited 0)
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${P}_plain_cfg2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${P}_launches_cfg2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${P}_ncu_cfg2.log 2>&1
python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-init > gpurun_out/${P}_plain_cfg5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${P}_launches_cfg5.csv python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-init > gpurun_out/${P}_ncu_cfg5.log 2>&1
# full captures of the dominant kernels
ncu --set full --clock-control none --import-source on -k regex:lloyd_f64s -s 2 -c 1 -o gpurun_out/${P}_prof_cfg5 python bench.py --workload cfg5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-init > gpurun_out/${P}_ncu_cfg5s.log 2>&1
python tools/cfg2_run.py > gpurun_out/${P}_plain_cfg2s.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lloyd_tc4 -s 3 -c 2 -o gpurun_out/${P}_prof_cfg2 python tools/cfg2_run.py > gpurun_out/${P}_ncu_cfg2s.log 2>&1
tail -c 700 gpurun_out/${P}_bench_cfg2.json

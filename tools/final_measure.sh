# Round-2 closing measurements (run under gpurun): full GPU test suite, smoke, bench lines, launch list, ncu captures.
P=${P:-r2}
mkdir -p gpurun_out
(time python -m pytest tests -q -m gpu) > gpurun_out/${P}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${P}_pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${P}_smoke.log
python bench.py > gpurun_out/${P}_bench_default.json 2> gpurun_out/${P}_bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${P}_bench_reference.json 2>/dev/null; echo "reference rc=$?"
# launch list of the headline command (after the plain run exited 0); shares, not absolutes
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-others --no-init > gpurun_out/${P}_plain_cfg3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${P}_launches_cfg3.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-others --no-init > gpurun_out/${P}_ncu_cfg3.log 2>&1
python tools/launch_summary.py gpurun_out/${P}_launches_cfg3.csv > gpurun_out/${P}_launch_summary_cfg3.txt; cat gpurun_out/${P}_launch_summary_cfg3.txt
# full capture of the dominant kernel (DRAM traffic per launch, pipe utilisation)
WL=cfg3 ITERS=3 python tools/r2_probe.py > gpurun_out/${P}_plain_probe.log 2>&1 && \
WL=cfg3 ITERS=3 ncu --set full --clock-control none --import-source on -k regex:assign_tc5s -s 3 -c 1 -f -o gpurun_out/${P}_prof_cfg3 python tools/r2_probe.py > gpurun_out/${P}_ncu_cfg3s.log 2>&1
ls -la gpurun_out/${P}_prof_cfg3.ncu-rep

# Round-2 closing measurements, second pass (after the sorted update / pivot pruning): tests, smoke, both bench arms, launch list and
# one full capture for the new kernels (unordered rows from a k-means++ start)
P=r2b
mkdir -p gpurun_out
(time python -m pytest tests -q -m gpu) > gpurun_out/${P}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${P}_pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${P}_smoke.log
python bench.py > gpurun_out/${P}_bench_default.json 2> gpurun_out/${P}_bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${P}_bench_reference.json 2>/dev/null; echo "reference rc=$?"
WL=cfg3 HASHED=1 KPP=1 ITERS=4 python tools/r2_probe.py > gpurun_out/${P}_plain_probe.log 2>&1 && \
WL=cfg3 HASHED=1 KPP=1 ITERS=4 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${P}_launches_cfg3_hashed_kpp.csv python tools/r2_probe.py > gpurun_out/${P}_ncu_l.log 2>&1
python tools/launch_summary.py gpurun_out/${P}_launches_cfg3_hashed_kpp.csv > gpurun_out/${P}_launch_summary_cfg3_hashed_kpp.txt; cat gpurun_out/${P}_launch_summary_cfg3_hashed_kpp.txt
# (the full capture of update_sorted_kernel — profiles/r2b_update_sorted_ncu.txt — was taken by an earlier run of this script:
#  WL=cfg3 HASHED=1 ITERS=3 ncu --set full --clock-control none --import-source on -k regex:update_sorted -s 1 -c 1 -f -o gpurun_out/r2b_prof_update_sorted python tools/r2_probe.py)

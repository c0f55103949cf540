#!/bin/bash
# round 2, first GPU call: tests, smoke, default bench, fixed-overhead probe of the cfg3 shape
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest.log
tail -5 gpurun_out/r2_pytest.log
python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"
for rows in 1250000 2500000 5000000; do
  python bench.py --rows $rows --no-others --no-e2e --no-cpu-baseline --no-init > gpurun_out/r2_bench_cfg3_rows$rows.json 2>> gpurun_out/r2_bench_default.err
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2>&1

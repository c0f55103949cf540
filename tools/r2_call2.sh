#!/bin/bash
# round 2: tests, smoke, default bench after the per-call adaptive choice
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2b_pytest.log
tail -8 gpurun_out/r2b_pytest.log
python __graft_entry__.py --smoke > gpurun_out/r2b_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2b_smoke.log
python bench.py > gpurun_out/r2b_bench_default.json 2> gpurun_out/r2b_bench_default.err; echo "bench rc=$?"

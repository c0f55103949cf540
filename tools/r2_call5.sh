#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2f_pytest.log
python bench.py > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err; echo "bench rc=$?"

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_harness.py tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2i_pytest.log

#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_rowwise.py tests/test_metrics_io.py -m gpu -q -x -k "DMMA or silhouette or example" > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2g_pytest.log
python tools/dmma_timing.py 2>&1 | tee gpurun_out/r2g_dmma_timing.txt

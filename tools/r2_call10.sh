#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rowwise.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py tests/test_harness.py -m gpu -q -x > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2l_pytest.log
for rows in 1250000 10000000; do
  echo "== cfg3 rows=$rows"; LKM_STEP_TRACE=1 WL=cfg3 ROWS=$rows ITERS=20 python tools/r2_probe.py 2>&1 | grep -v "^$" | tail -6
  WL=cfg3 ROWS=$rows ITERS=20 python tools/r2_probe.py 2>&1 | tail -1
done
echo "== cfg3 hashed 2M"; LKM_STEP_TRACE=1 WL=cfg3 ROWS=2000000 HASHED=1 ITERS=10 python tools/r2_probe.py 2>&1 | tail -6
echo "== cfg4"; LKM_STEP_TRACE=1 WL=cfg4 ROWS=1562500 ITERS=10 python tools/r2_probe.py 2>&1 | tail -6
WL=cfg4 ROWS=1562500 ITERS=10 python tools/r2_probe.py 2>&1 | tail -1

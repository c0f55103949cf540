#!/bin/bash
python -m pytest tests/test_gpu_rowwise.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-150
echo "== cfg3 hashed 2M"; LKM_STEP_TRACE=1 WL=cfg3 ROWS=2000000 HASHED=1 ITERS=10 python tools/r2_probe.py 2>&1 | tail -6 | grep "update\|step "
echo "== cfg3 blob 2M nofuse"; LKM_TC5_FUSE=0 LKM_STEP_TRACE=1 WL=cfg3 ROWS=2000000 ITERS=10 python tools/r2_probe.py 2>&1 | tail -6 | grep "update\|step "
echo "== cfg4 hashed"; LKM_STEP_TRACE=1 WL=cfg4 ROWS=1562500 HASHED=1 ITERS=10 python tools/r2_probe.py 2>&1 | tail -6 | grep "update\|step "

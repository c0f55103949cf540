#!/bin/bash
mkdir -p gpurun_out
echo "== cfg3 hashed 2M"; LKM_STEP_TRACE=1 WL=cfg3 ROWS=2000000 HASHED=1 ITERS=10 python tools/r2_probe.py 2>&1 | tail -6 | grep "update\|step "
echo "== cfg3 blob 2M nofuse"; LKM_TC5_FUSE=0 LKM_STEP_TRACE=1 WL=cfg3 ROWS=2000000 ITERS=10 python tools/r2_probe.py 2>&1 | tail -6 | grep "update\|step "
echo "== cfg4 hashed"; LKM_STEP_TRACE=1 WL=cfg4 ROWS=1562500 HASHED=1 ITERS=10 python tools/r2_probe.py 2>&1 | tail -6 | grep "update\|step "
WL=cfg3 ROWS=2000000 HASHED=1 ITERS=3 ncu --set full --clock-control none --import-source on -k regex:update_rows -s 3 -c 1 -o gpurun_out/r2n_upd_hashed -f python tools/r2_probe.py > /dev/null 2>&1
ls -la gpurun_out/r2n_upd_hashed.ncu-rep

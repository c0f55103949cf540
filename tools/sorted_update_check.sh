#!/bin/bash
# sorted update (lkm_sortupd.cuh): parity tests, then hashed-row timings with and without it (GPU box)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sorted_update.py tests/test_harness.py -x -q 2>&1 | tail -15
for wl in cfg3 cfg4; do
  for s in 0 -1; do
    echo "LKM_SORT_UPD=$s"
    LKM_SORT_UPD=$s WL=$wl HASHED=1 ITERS=10 python tools/r2_probe.py 2>&1 | tail -1
    LKM_SORT_UPD=$s WL=$wl HASHED=1 KPP=1 ITERS=10 python tools/r2_probe.py 2>&1 | tail -1
  done
  LKM_SORT_UPD=-1 WL=$wl HASHED=0 ITERS=10 python tools/r2_probe.py 2>&1 | tail -1
done
LKM_STEP_TRACE=1 LKM_SORT_UPD=-1 WL=cfg3 HASHED=1 ITERS=4 python tools/r2_probe.py 2>&1 | tail -30

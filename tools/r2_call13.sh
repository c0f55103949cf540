#!/bin/bash
for tc in 0 1; do
  LKM_PARA_TC=$tc WL=cfg3 ROWS=4000000 python tools/init_timing.py 2>&1 | tail -4
  LKM_PARA_TC=$tc WL=cfg4 ROWS=4000000 python tools/init_timing.py 2>&1 | tail -2
done
LKM_STEP_TRACE=1 WL=cfg3 ROWS=2000000 HASHED=1 ITERS=10 python tools/r2_probe.py 2>&1 | tail -6 | grep "update\|step "

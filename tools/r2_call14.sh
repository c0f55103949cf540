#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2p_pytest.log
python __graft_entry__.py --smoke > gpurun_out/r2p_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2p_smoke.log
WL=cfg3 ROWS=4000000 python tools/init_timing.py 2>&1 | tail -4
WL=cfg4 ROWS=4000000 python tools/init_timing.py 2>&1 | tail -2
WL=cfg2 python tools/init_timing.py 2>&1 | tail -4

#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_rowwise.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2j_pytest.log
NCU="ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv -k regex:assign_tc5|update_rows|finalize|rows_top2"
run() { # name, env...
  name=$1; shift
  env "$@" python tools/r2_probe.py > gpurun_out/r2j_$name.txt 2>&1
  env "$@" $NCU --log-file gpurun_out/r2j_$name.csv python tools/r2_probe.py > /dev/null 2>&1
  cat gpurun_out/r2j_$name.txt | tail -1
  python tools/launch_summary.py gpurun_out/r2j_$name.csv
}
run cfg3_1p25m WL=cfg3 ROWS=1250000 ITERS=8
run cfg3_1p25m_dense WL=cfg3 ROWS=1250000 ITERS=8 LKM_SPARSE_PARTIALS=0
run cfg3_10m WL=cfg3 ITERS=8
run cfg4_2m WL=cfg4 ROWS=2000000 ITERS=6

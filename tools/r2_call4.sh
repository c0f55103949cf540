#!/bin/bash
# round 2: Hamerly, deferred re-scans (rows_top2_kernel), round-based update of unordered rows
mkdir -p gpurun_out
python -m pytest tests/test_hamerly.py -m gpu -x -q > gpurun_out/r2d_hamerly.log 2>&1; echo "hamerly rc=$?"; tail -15 gpurun_out/r2d_hamerly.log
python -m pytest tests -m gpu -q -x --deselect tests/test_hamerly.py > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2d_pytest.log
NCU="ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv -k regex:assign_tc5|update_rows|finalize|rows_top2"
run() { # name, env...
  name=$1; shift
  env "$@" python tools/r2_probe.py > gpurun_out/r2d_$name.txt 2>&1
  env "$@" SKIP_INIT_CAPTURE=1 $NCU --log-file gpurun_out/r2d_$name.csv python tools/r2_probe.py > /dev/null 2>&1
  cat gpurun_out/r2d_$name.txt | tail -1
  python tools/launch_summary.py gpurun_out/r2d_$name.csv | grep -v "gen_blobs\|row_norm\|block_sums\|fast_pick\|lloyd_kernel<float, 0, 0>"
}
run cfg3_hashed WL=cfg3 ROWS=2000000 HASHED=1
run cfg3_kpp WL=cfg3 ROWS=2000000 KPP=1 ITERS=6
run cfg3_kpp_p3 WL=cfg3 ROWS=2000000 KPP=1 PASSES=3 ITERS=6
run cfg3_kpp_p1 WL=cfg3 ROWS=2000000 KPP=1 PASSES=1 ITERS=6
run cfg4_hashed WL=cfg4 ROWS=2000000 HASHED=1
run cfg4_kpp WL=cfg4 ROWS=2000000 KPP=1 ITERS=6

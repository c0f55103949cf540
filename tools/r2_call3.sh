#!/bin/bash
# per-kernel launch lists of the unordered / k-means++-started shapes
mkdir -p gpurun_out
NCU="ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv"
run() { # name, env...
  name=$1; shift
  env "$@" python tools/r2_probe.py > gpurun_out/r2c_$name.txt 2>&1
  env "$@" $NCU --log-file gpurun_out/r2c_$name.csv python tools/r2_probe.py > /dev/null 2>&1
  cat gpurun_out/r2c_$name.txt | tail -1
  python tools/launch_summary.py gpurun_out/r2c_$name.csv | grep -v "gen_blobs\|row_norm"
}
run cfg3_blob WL=cfg3 ROWS=2000000
run cfg3_hashed WL=cfg3 ROWS=2000000 HASHED=1
run cfg3_hashed_nofuse WL=cfg3 ROWS=2000000 HASHED=1 LKM_TC5_FUSE=0
run cfg3_kpp WL=cfg3 ROWS=2000000 KPP=1
run cfg3_kpp_p3 WL=cfg3 ROWS=2000000 KPP=1 PASSES=3
run cfg2_hashed WL=cfg2 HASHED=1
run cfg2_kpp WL=cfg2 KPP=1
run cfg4_hashed WL=cfg4 ROWS=2000000 HASHED=1
python -m pytest tests/test_gpu_rowwise.py -m gpu -q -k "kmeans_para" 2>&1 | tail -3

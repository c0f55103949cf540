#!/bin/bash
# Event trace of the pair kernel (tiles 32-47 of CTAs 0 and 1; columns in lkm_api.cu, "t5trace").  Build the trace library first, here:
#   python -c "import os; from linfa_b200 import build; build._compile(os.path.join(build.HERE, 'liblkm_trace.so'), ['-DLKM_T5_TRACE'], False)"
# then on the GPU box: bash tools/t5_trace.sh   (cfg4 shape, no fused update; N / D / K / LKM_TC5_FUSE override)
export LKM_LIB_PATH=/root/repo/linfa_b200/liblkm_trace.so
LKM_TC5_FUSE=${LKM_TC5_FUSE:-0} N=${N:-2000000} D=${D:-64} K=${K:-1024} timeout 120 python tools/cfg3_trace.py 2>&1 | tail -34

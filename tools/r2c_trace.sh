#!/bin/bash
# -DLKM_T5_TRACE build (linfa_b200/liblkm_trace.so): event trace of the pair kernel (tiles 32-47 of CTAs 0 and 1), cfg4 shape, no fused update
export LKM_LIB_PATH=/root/repo/linfa_b200/liblkm_trace.so
LKM_TC5_FUSE=0 N=2000000 D=64 K=1024 timeout 120 python tools/cfg3_trace.py 2>&1 | tail -34

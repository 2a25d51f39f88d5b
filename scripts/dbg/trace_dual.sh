#!/bin/bash
# per-warp phase trace of the fused kernel (needs weaklysuperviseddl_b200/libwsdl_b200_trace.so, see scripts/gpu_trace.sh)
cp weaklysuperviseddl_b200/libwsdl_b200.so /tmp/keep.so
cp weaklysuperviseddl_b200/libwsdl_b200_trace.so weaklysuperviseddl_b200/libwsdl_b200.so
PYTHONPATH=. timeout -s KILL 200 python scripts/trace_ctas.py dual 2>&1 | grep -v Warn
cp /tmp/keep.so weaklysuperviseddl_b200/libwsdl_b200.so

#!/bin/bash
# per-warp / per-group phase traces (globaltimer) of the pairwise kernels: needs weaklysuperviseddl_b200/libwsdl_b200_trace.so,
# built here with: WSDL_NVCC_EXTRA=-DWSDL_PS_TRACE python -m weaklysuperviseddl_b200.build --force &&
#   cp weaklysuperviseddl_b200/libwsdl_b200.so weaklysuperviseddl_b200/libwsdl_b200_trace.so && python -m weaklysuperviseddl_b200.build --force
[ -f weaklysuperviseddl_b200/libwsdl_b200_trace.so ] || { echo "build the trace library first (see the header of this script)"; exit 1; }
cp weaklysuperviseddl_b200/libwsdl_b200.so /tmp/keep.so
cp weaklysuperviseddl_b200/libwsdl_b200_trace.so weaklysuperviseddl_b200/libwsdl_b200.so
PYTHONPATH=. timeout -s KILL 200 python scripts/trace_stream.py 2>&1 | grep -v Warn
cp /tmp/keep.so weaklysuperviseddl_b200/libwsdl_b200.so

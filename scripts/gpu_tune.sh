#!/bin/bash
# swaps in the tuning build (weaklysuperviseddl_b200/libwsdl_b200_tune.so, built with WSDL_NVCC_EXTRA=-DWSDL_TUNING) and runs "$@"
[ -f weaklysuperviseddl_b200/libwsdl_b200_tune.so ] || { echo "no tuning library"; exit 1; }
cp weaklysuperviseddl_b200/libwsdl_b200.so /tmp/keep.so
cp weaklysuperviseddl_b200/libwsdl_b200_tune.so weaklysuperviseddl_b200/libwsdl_b200.so
"$@"
cp /tmp/keep.so weaklysuperviseddl_b200/libwsdl_b200.so

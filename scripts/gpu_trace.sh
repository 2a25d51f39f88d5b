[ -f weaklysuperviseddl_b200/libwsdl_b200_trace.so ] || { echo "build it first: WSDL_NVCC_EXTRA=-DWSDL_PS_TRACE python -m weaklysuperviseddl_b200.build --force && cp weaklysuperviseddl_b200/libwsdl_b200.so weaklysuperviseddl_b200/libwsdl_b200_trace.so && python -m weaklysuperviseddl_b200.build --force"; exit 1; }
# per-warp phase trace of the default pairwise kernel (needs weaklysuperviseddl_b200/libwsdl_b200_trace.so, built
# with WSDL_NVCC_EXTRA=-DWSDL_PS_TRACE)
cp weaklysuperviseddl_b200/libwsdl_b200.so /tmp/keep.so
cp weaklysuperviseddl_b200/libwsdl_b200_trace.so weaklysuperviseddl_b200/libwsdl_b200.so
PYTHONPATH=. python scripts/trace_ctas.py $TRACE_CASES 2>&1 | grep -v Warn
cp /tmp/keep.so weaklysuperviseddl_b200/libwsdl_b200.so

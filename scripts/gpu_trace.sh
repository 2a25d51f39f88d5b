cp weaklysuperviseddl_b200/libwsdl_b200.so /tmp/keep.so
cp weaklysuperviseddl_b200/libwsdl_b200_trace.so weaklysuperviseddl_b200/libwsdl_b200.so
PYTHONPATH=. python scripts/trace_ctas.py 2>&1 | grep -v Warning
cp /tmp/keep.so weaklysuperviseddl_b200/libwsdl_b200.so

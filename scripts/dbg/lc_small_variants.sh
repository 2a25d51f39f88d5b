cp weaklysuperviseddl_b200/libwsdl_b200.so /tmp/keep.so
cp weaklysuperviseddl_b200/libwsdl_b200_tune.so weaklysuperviseddl_b200/libwsdl_b200.so
echo "== cluster path, unroll 8"; PYTHONPATH=. python scripts/bench_layercam_small.py 2>&1 | tail -5
echo "== two-kernel path (round 1)"; WSDL_LAYERCAM_NO_CLUSTER=1 PYTHONPATH=. python scripts/bench_layercam_small.py 2>&1 | tail -5
cp weaklysuperviseddl_b200/libwsdl_b200_tune4.so weaklysuperviseddl_b200/libwsdl_b200.so
echo "== cluster path, unroll 4"; PYTHONPATH=. python scripts/bench_layercam_small.py 2>&1 | tail -5
cp /tmp/keep.so weaklysuperviseddl_b200/libwsdl_b200.so

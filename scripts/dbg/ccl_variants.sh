cp weaklysuperviseddl_b200/libwsdl_b200.so /tmp/keep.so
for f in weaklysuperviseddl_b200/libccl_*.so; do cp $f weaklysuperviseddl_b200/libwsdl_b200.so; echo "== $f"; PYTHONPATH=. python scripts/bench_ccl.py 2>&1 | tail -3; done
cp /tmp/keep.so weaklysuperviseddl_b200/libwsdl_b200.so

# times bench.py with alternative builds of the library (weaklysuperviseddl_b200/libwsdl_b200_<tag>.so)
cp weaklysuperviseddl_b200/libwsdl_b200.so /tmp/keep.so
for f in weaklysuperviseddl_b200/libwsdl_b200.so weaklysuperviseddl_b200/libwsdl_b200_S*.so; do
  cp $f /tmp/v.so; cp /tmp/v.so weaklysuperviseddl_b200/libwsdl_b200.so
  echo -n "$(basename $f): "; python bench.py --steps 400 --warmup 50 --no-cpu-baseline --no-also 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('us/step', round(1000*d['ms_per_step'],2), 'Gpix/s', round(d['value'],2), d['roofline']['per_kernel_ms_direct_launch'])"
  cp /tmp/keep.so weaklysuperviseddl_b200/libwsdl_b200.so
done

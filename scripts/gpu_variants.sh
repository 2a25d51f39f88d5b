cp weaklysuperviseddl_b200/libwsdl_b200.so /tmp/keep.so
for v in "" _NOTRANSFORM _NOMARCH _NOTAIL _MARCHONLY _LOADONLY; do
  cp weaklysuperviseddl_b200/libwsdl_b200$v.so /tmp/v.so; cp /tmp/v.so weaklysuperviseddl_b200/libwsdl_b200.so
  echo -n "variant '$v': "; python bench.py --steps 400 --warmup 50 --no-cpu-baseline --no-also 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('us/step', 1000*d['ms_per_step'])"
  cp /tmp/keep.so weaklysuperviseddl_b200/libwsdl_b200.so
done

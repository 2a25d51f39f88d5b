# sweeps bench protocol knobs / build variants (weaklysuperviseddl_b200/libwsdl_b200_S_<tag>.so) for the pairwise workload
run() { echo -n "$1 | $2: "; env ${1//,/ } python bench.py --steps 1600 --warmup 100 --no-cpu-baseline --no-also $2 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('us/step', round(1000*d['ms_per_step'],2), 'Gpix/s', round(d['value'],2), 'alone', round(1000*d['roofline']['per_kernel_ms_direct_launch']['fused cut+boundary'],2))"; }
cp weaklysuperviseddl_b200/libwsdl_b200.so /tmp/keep.so
for f in ${LIBS:-weaklysuperviseddl_b200/libwsdl_b200.so}; do
  cp $f /tmp/v.so; cp /tmp/v.so weaklysuperviseddl_b200/libwsdl_b200.so; echo "== $f"
  for e in ${ENVS:-X=0}; do for a in "${ARGS[@]:---graph-steps 32}"; do run "$e" "$a"; done; done
  cp /tmp/keep.so weaklysuperviseddl_b200/libwsdl_b200.so
done

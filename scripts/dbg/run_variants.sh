#!/bin/bash
# on the GPU box: the configs[1] bench line (device-timed, overlapped and alone) for every experiment library
cp weaklysuperviseddl_b200/libwsdl_b200.so /tmp/keep.so
for f in /tmp/keep.so weaklysuperviseddl_b200/libx_*.so; do
  [ -f "$f" ] || continue
  [ "$f" != /tmp/keep.so ] && cp "$f" weaklysuperviseddl_b200/libwsdl_b200.so
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$f', 'Gpix/s %.1f' % d['value'], 'step_us %.2f' % (d['ms_per_step']*1e3), 'alone_us %.2f' % (r['kernel_ms_alone']*1e3), 'frac %.3f' % r['frac'])
" || echo "$f FAILED"
done
cp /tmp/keep.so weaklysuperviseddl_b200/libwsdl_b200.so

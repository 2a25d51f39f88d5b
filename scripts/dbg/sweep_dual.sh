#!/bin/bash
# tuning build: stagger / row-block sweeps of the fused launch (scripts/gpu_tune.sh swaps the library in)
run() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1', 'step_us %.2f' % (d['ms_per_step']*1e3), 'alone_us %.2f' % (r['kernel_ms_alone']*1e3))"; }
run default
for s in 0 200 800 1600; do WSDL_PS_STAGGER_NS=$s run "stagger=$s"; done
for nb in 6 7 8; do WSDL_PS_NB=$nb run "nb=$nb"; done

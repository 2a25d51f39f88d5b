for ns in 0 1000 2000 3000 4000 6000; do
  WSDL_SG_STAGGER_NS=$ns python bench.py --no-cpu-baseline --no-also --steps 2000 --warmup 100 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('stagger $ns', round(d['value'],1), round(d['ms_per_step']*1e3,2), round(d['roofline']['per_kernel_ms_direct_launch']['fused cut+boundary']*1e3,2))"
done

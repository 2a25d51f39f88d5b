cp weaklysuperviseddl_b200/libwsdl_b200.so /tmp/keep.so
cp weaklysuperviseddl_b200/libwsdl_b200_trace.so weaklysuperviseddl_b200/libwsdl_b200.so
for pad in 0 20 55 120; do echo "=== smem pad $pad KB"; WSDL_PS_SMEM_PAD_KB=$pad PYTHONPATH=. python scripts/trace_ctas.py 2>&1 | grep -A30 "== cut" | grep -E "== cut|first wave|step[0-3]|march end|CTA lifetime|per-SM"; done
cp /tmp/keep.so weaklysuperviseddl_b200/libwsdl_b200.so

[ -f weaklysuperviseddl_b200/libwsdl_b200_trace.so ] || { echo "build it first: WSDL_NVCC_EXTRA=-DWSDL_PS_TRACE python -m weaklysuperviseddl_b200.build --force && cp weaklysuperviseddl_b200/libwsdl_b200.so weaklysuperviseddl_b200/libwsdl_b200_trace.so && python -m weaklysuperviseddl_b200.build --force"; exit 1; }
export WSDL_PAIRWISE_PIPE=1
timeout 900 python -m pytest tests/test_gpu_pairwise.py tests/test_gpu_dropin.py -x -q > gpurun_out/pytest_pipe.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_pipe.log | grep -E "passed|failed|Error|assert"
timeout 300 python bench.py --steps 800 --warmup 50 --no-cpu-baseline --no-also 2> gpurun_out/bench_pipe.err | python -c "import json,sys; d=json.load(sys.stdin); print('PIPE Gpix/s', d['value'], 'ms/step', d['ms_per_step'], d['roofline']['per_kernel_ms_direct_launch'])"; tail -3 gpurun_out/bench_pipe.err
cp weaklysuperviseddl_b200/libwsdl_b200.so /tmp/keep.so
cp weaklysuperviseddl_b200/libwsdl_b200_trace.so weaklysuperviseddl_b200/libwsdl_b200.so
PYTHONPATH=. python scripts/trace_pipe.py 2>&1 | grep -v Warn
cp /tmp/keep.so weaklysuperviseddl_b200/libwsdl_b200.so

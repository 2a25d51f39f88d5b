timeout 900 python -m pytest tests/test_gpu_pairwise.py -x -q > gpurun_out/pytest_pw.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_pw.log
timeout 600 python bench.py --steps 800 --warmup 50 --no-cpu-baseline --no-also > gpurun_out/bench_pw.json 2> gpurun_out/bench_pw.err; echo "bench rc=$?"; cat gpurun_out/bench_pw.json | python -c "import json,sys; d=json.load(sys.stdin); print('Gpix/s', d['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])"; tail -3 gpurun_out/bench_pw.err
CMD="python bench.py --steps 16 --warmup 8 --no-graph --no-cpu-baseline --no-also"
$CMD > gpurun_out/plain_pair2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pairwise_fast -s 20 -c 2 -f -o gpurun_out/prof_pairwise_v2 $CMD > gpurun_out/ncu_pair_full.log 2>&1
echo "pair full rc=$?"

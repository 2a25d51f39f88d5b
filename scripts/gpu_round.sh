# full round-end style run: smoke, GPU tests, bench (both arms), launch list + full ncu capture of the dominant kernels
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --workload layercam --steps 60 --warmup 5 > gpurun_out/bench_layercam.json 2> gpurun_out/bench_layercam.err; echo "bench layercam rc=$?"; cat gpurun_out/bench_layercam.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cat gpurun_out/bench_ref.json
timeout 600 python bench.py --impl reference --workload layercam --steps 5 --warmup 1 > gpurun_out/bench_ref_layercam.json 2>> gpurun_out/bench_ref.err; echo "bench ref layercam rc=$?"; cat gpurun_out/bench_ref_layercam.json
CMD="python bench.py --steps 16 --warmup 8 --no-graph --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:pairwise_ -s 20 -c 3 -f -o gpurun_out/prof_pairwise_sym $CMD > gpurun_out/ncu_pair_full.log 2>&1; echo "ncu pair rc=$?"
ncu --set full --clock-control none --import-source on -k regex:layercam -s 4 -c 2 -f -o gpurun_out/prof_layercam python bench.py --workload layercam --steps 4 --warmup 3 > gpurun_out/ncu_lc_full.log 2>&1; echo "ncu layercam rc=$?"

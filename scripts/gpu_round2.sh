#!/bin/bash
# round-2 evidence run on one B200: tests, bench lines, ncu launch lists and one --set full capture (after plain runs exited 0)
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_pytest_gpu.log; cat gpurun_out/r2_pytest_gpu.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err || tail -5 gpurun_out/r2_bench_n1.err
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_n1.json 2>/dev/null
python bench.py --workload trainstep > gpurun_out/r2_bench_trainstep_n1.json 2> gpurun_out/r2_bench_trainstep_n1.err || tail -5 gpurun_out/r2_bench_trainstep_n1.err
python bench.py --workload layercam --steps 5 > gpurun_out/r2_bench_layercam_n1.json 2> gpurun_out/r2_bench_layercam_n1.err || tail -5 gpurun_out/r2_bench_layercam_n1.err
CMD="python bench.py --steps 16 --warmup 8 --no-graph --no-cpu-baseline --no-also"
$CMD > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv $CMD > gpurun_out/ncu_bench.log 2>&1
python scripts/ncu_target.py > gpurun_out/plain_target.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"pairwise_dual|weak_loss|ccl_" -c 40 --csv --log-file gpurun_out/r2_launches_target.csv python scripts/ncu_target.py > gpurun_out/ncu_target.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pairwise_dual|weak_loss" -s 5 -c 5 -f -o gpurun_out/r2_prof_pairwise python scripts/ncu_target.py > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
for f in gpurun_out/r2_bench_n1.json gpurun_out/r2_bench_trainstep_n1.json gpurun_out/r2_bench_layercam_n1.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], d["metric"][:40], round(d["value"],2), d["unit"], "ms/step", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],2), d.get("stage_share",""))
PY
done
PYTHONPATH=. python scripts/bench_ccl.py > gpurun_out/r2_ccl_bench.txt 2>&1
PYTHONPATH=. python scripts/bench_refine.py > gpurun_out/r2_refine.txt 2>&1
PYTHONPATH=. python scripts/sweep_layercam.py > gpurun_out/r2_sweep_layercam.txt 2>gpurun_out/sweep_layercam.err
PYTHONPATH=. python scripts/bench_layercam_small.py > gpurun_out/r2_layercam_small.txt 2>&1

set -x
CMD="python bench.py --steps 16 --warmup 8 --no-graph --no-cpu-baseline --no-also"
$CMD > gpurun_out/plain_pair.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pairwise.csv $CMD > gpurun_out/ncu_pair.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain_pair2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pairwise_fwd_bwd -s 20 -c 2 -f -o gpurun_out/prof_pairwise $CMD > gpurun_out/ncu_pair_full.log 2>&1
echo "pair full rc=$?"
CMD2="python bench.py --workload layercam --steps 3 --warmup 3"
$CMD2 > gpurun_out/plain_lc.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:layercam -s 6 -c 2 -f -o gpurun_out/prof_layercam $CMD2 > gpurun_out/ncu_lc_full.log 2>&1
echo "lc full rc=$?"
ls -la gpurun_out

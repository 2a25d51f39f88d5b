# launch list + one full capture of the pairwise kernels (run after the plain command exited 0)
CMD="python bench.py --steps 16 --warmup 8 --no-graph --no-cpu-baseline --no-also"
$CMD > gpurun_out/plain_pair.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_pair.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pairwise -s 20 -c 8 --csv --log-file gpurun_out/launches_pairwise.csv $CMD > gpurun_out/ncu_pair.log 2>&1
grep '^"[0-9]' gpurun_out/launches_pairwise.csv | awk -F'","' '{print $5, $NF}'
ncu --set full --clock-control none --import-source on -k regex:pairwise_ -s 20 -c 3 -f -o gpurun_out/prof_pairwise_sym $CMD > gpurun_out/ncu_pair_full.log 2>&1
echo "full rc=$?"

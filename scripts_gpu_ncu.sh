set -x
CMD="python bench.py --steps 16 --warmup 8 --no-graph --no-cpu-baseline --no-also"
$CMD > gpurun_out/plain_pair.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pairwise_sym -s 20 -c 2 -f -o gpurun_out/prof_pairwise_sym $CMD > gpurun_out/ncu_pair_full.log 2>&1
echo "pair full rc=$?"

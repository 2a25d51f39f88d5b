CMD="python bench.py --steps 16 --warmup 8 --no-graph --no-cpu-baseline --no-also"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pairwise -s 20 -c 8 --csv --log-file gpurun_out/launches_pairwise.csv $CMD > gpurun_out/ncu_pair.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/launches_pairwise.csv') if l.startswith('"')))
h=rows[0]
for r in rows[1:]:
    print(r[h.index('Kernel Name')][:60], r[h.index('Metric Value')])
PY

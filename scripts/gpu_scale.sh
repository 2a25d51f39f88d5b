#!/bin/bash
# bench.py under torchrun on N GPUs of one box, the way the driver launches it: scripts/gpu_scale.sh N
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err || tail -20 gpurun_out/r2_bench_n$N.err
python - gpurun_out/r2_bench_n$N.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
a=d.get("also",{})
print("N", d["n_gpus"], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],2), {k:round(v["value"],2) for k,v in d["e2e"]["variants"].items()})
for k,v in a.items():
    print(" also", k, v.get("value"), v.get("unit"), v.get("error"), v.get("ms_per_pass", v.get("ms_per_step")), (v.get("roofline") or {}).get("frac"), v.get("stage_share"))
PY

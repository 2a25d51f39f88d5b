"""Turns the artefacts of scripts/gpu_round.sh (gpurun_out/) into the tracked summaries under profiles/."""
import collections, csv, json, os, shutil, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
go, pr = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")
rows = [r for r in csv.reader(l for l in open(os.path.join(go, "launches_bench.csv")) if l.startswith('"'))]
h = rows[0]; ik, iv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    a = agg.setdefault(r[ik], [0, 0.0]); a[0] += 1; a[1] += float(r[iv].replace(",", ""))
tot = sum(v[1] for v in agg.values())
with open(os.path.join(pr, f"{tag}_b_launches_bench.txt"), "w") as f:
    f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none), final state of the round:\n"
            "#   python bench.py --steps 16 --warmup 8 --no-graph --no-cpu-baseline   (first 400 launches: input synthesis, then\n"
            "#   warm-up/timed steps of configs[1] and the per-kernel / e2e / layercam legs; per-launch times are cold-cache and\n"
            "#   serialised -- shares, not absolutes)\n# kernel | launches | total ns | mean ns | share\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k[:110]} | {n} | {t:.0f} | {t/n:.0f} | {100*t/tot:.1f}%\n")
def run(*a):
    return subprocess.run([sys.executable, *a], capture_output=True, text=True, cwd=R).stdout
sym = os.path.join(go, "prof_pairwise_sym.ncu-rep")
with open(os.path.join(pr, f"{tag}_c_ncu_full_sym.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none: python bench.py --steps 16 --warmup 8 --no-graph --no-cpu-baseline, -k regex:pairwise_ -s 20 -c 3\n")
    f.write(run("scripts/ncu_summary.py", sym))
    f.write("\n# SASS opcode mix + stall samples of launch 0\n" + run("scripts/ncu_sass.py", sym, "0"))
    f.write("\n# march loop only\n" + run("scripts/ncu_hot.py", sym, "12"))
with open(os.path.join(pr, f"{tag}_c_ncu_full_layercam.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none: python bench.py --workload layercam --steps 4 --warmup 3, -k regex:layercam -s 4 -c 2\n")
    f.write(run("scripts/ncu_summary.py", os.path.join(go, "prof_layercam.ncu-rep")))
for src, dst in (("bench.json", "bench_n1.json"), ("bench_layercam.json", "bench_layercam_n1.json"), ("bench_ref.json", "bench_reference_n1.json")):
    shutil.copy(os.path.join(go, src), os.path.join(pr, f"{tag}_{dst}"))
d = json.load(open(os.path.join(go, "bench.json")))
print("pairwise", round(d["value"], 2), "Gpix/s", round(d["ms_per_step"] * 1e3, 2), "us/step frac", round(d["roofline"]["frac"], 4),
      d["roofline"]["per_kernel_ms_direct_launch"], "e2e", round(d["e2e"]["value"], 2), "cpu", d["cpu_baseline"]["value"],
      "layercam", round(d["also"]["layercam_512"]["value"]), d["also"]["layercam_512"]["roofline"]["frac"])

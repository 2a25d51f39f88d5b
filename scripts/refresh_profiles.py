"""Turns the artefacts of scripts/gpu_round2.sh / gpu_scale.sh (gpurun_out/r2_*) into the tracked summaries under profiles/
(r02_*).  Runs here (no GPU): reads .ncu-rep files with `ncu -i`."""
import collections, csv, json, os, shutil, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = "r02"
go, pr = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")


def ncu_rows(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    return H, [r for r in rows[hdr + 1:] if len(r) >= len(H)]


def run(*a):
    return subprocess.run([sys.executable, *a], capture_output=True, text=True, cwd=R).stdout


# ---- launch list of the default bench line
p = os.path.join(go, "r2_launches_bench.csv")
if os.path.exists(p):
    H, rows = ncu_rows(p)
    ik, iv = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(r[ik], [0, 0.0]); a[0] += 1; a[1] += float(r[iv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(pr, f"{tag}_b_launches_bench.txt"), "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none), final state of round 2:\n"
                "#   python bench.py --steps 16 --warmup 8 --no-graph --no-cpu-baseline --no-also   (first 600 launches: input\n"
                "#   synthesis, warm-up / timed steps of configs[1], the per-kernel legs and the three e2e legs; per-launch times are\n"
                "#   cold-cache and serialised -- shares, not absolutes)\n# kernel | launches | total ns | mean ns | share\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:110]} | {n} | {t:.0f} | {t/n:.0f} | {100*t/tot:.1f}%\n")

# ---- per-launch metrics of the pairwise / keep_largest kernels at full size
p = os.path.join(go, "r2_launches_target.csv")
if os.path.exists(p):
    H, rows = ncu_rows(p)
    ix = {h: i for i, h in enumerate(H)}
    d = collections.OrderedDict()
    for r in rows:
        key = (int(r[ix["ID"]]), r[ix["Kernel Name"]].split("(")[0].replace("void ", ""), r[ix["Grid Size"]])
        d.setdefault(key, {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    with open(os.path.join(pr, f"{tag}_d_launches_target.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,\n"
                "#   smsp__issue_active --clock-control none: python scripts/ncu_target.py (32x2x224x224 pairwise launches with a 256 MB\n"
                "#   L2 flush before each; keep_largest on 128 512x512 masks of two kinds).  Cold and serialised: compare shares and counts.\n"
                "# id | kernel | grid | us | dram read MB | dram write MB | warp instructions (M) | issue active %\n")
        ccl = collections.defaultdict(list)
        for (i, k, g), m in d.items():
            f.write(f"{i:3d} | {k:34s} | {g:14s} | {m['gpu__time_duration.sum']/1e3:7.1f} | {m['dram__bytes_read.sum']/1e6:6.1f} | "
                    f"{m['dram__bytes_write.sum']/1e6:6.1f} | {m['smsp__inst_executed.sum']/1e6:7.2f} | "
                    f"{m['smsp__issue_active.avg.pct_of_peak_sustained_active']:5.1f}\n")
            if k.startswith("ccl_"):
                ccl[k].append(m)
        if ccl:  # scripts/ncu_target.py alternates the two mask kinds: noise-outline masks, then CAM-like masks
            for which, name in ((-2, "box-filtered-noise masks (ragged outlines)"), (-1, "CAM-like masks (bilinear 16x16 -> 512x512)")):
                rd = sum(v[which]["dram__bytes_read.sum"] for v in ccl.values()) / 1e6
                wr = sum(v[which]["dram__bytes_write.sum"] for v in ccl.values()) / 1e6
                us = sum(v[which]["gpu__time_duration.sum"] for v in ccl.values()) / 1e3
                f.write(f"# keep_largest, one call (128 x 512^2), {name}: {us:.0f} us under ncu, DRAM {rd:.0f} MB read + {wr:.0f} MB written = "
                        f"{rd + wr:.0f} MB for 67 MB algorithmic (mask in + mask out); round 1: 473 MB, round 2 first version: 122 MB\n")

# ---- one --set full capture of the pairwise kernels
rep = os.path.join(go, "r2_prof_pairwise.ncu-rep")
if os.path.exists(rep):
    with open(os.path.join(pr, f"{tag}_c_ncu_full_pairwise.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:'pairwise_dual|weak_loss' -s 5 -c 5: python scripts/ncu_target.py\n"
                "# (32x2x224x224, 256 MB L2 flush before every launch)\n")
        f.write(run("scripts/ncu_summary.py", rep))
        for i in range(5):
            s = run("scripts/ncu_sass.py", rep, str(i))
            if s.strip():
                f.write(f"\n# SASS opcode mix + stall samples of launch {i}\n" + s)
        f.write("\n# march loop of launch 0\n" + run("scripts/ncu_hot.py", rep, "12"))

# ---- bench lines
for n in (1, 2, 4, 8):
    src = os.path.join(go, f"r2_bench_n{n}.json")
    if os.path.exists(src) and os.path.getsize(src) > 10:
        shutil.copy(src, os.path.join(pr, f"{tag}_bench_n{n}.json"))
for name in ("reference_n1", "trainstep_n1", "layercam_n1"):
    src = os.path.join(go, f"r2_bench_{name}.json")
    if os.path.exists(src) and os.path.getsize(src) > 10:
        shutil.copy(src, os.path.join(pr, f"{tag}_bench_{name}.json"))
with open(os.path.join(pr, f"{tag}_a_scale.txt"), "w") as f:
    f.write("# bench.py under torchrun on N B200 of one box (scripts/gpu_scale.sh N = the driver's launch line, --steps 20 --warmup 5)\n"
            "# N | configs[1] Gpix/s (weak) | e2e u8 Gpix/s | e2e f32 | e2e bf16 | config 3: masks/s over the 3680-image set (strong) | pass ms | "
            "frac of HBM per GPU | config 4: images/s (weak, DDP) | step ms | loss stage ms\n")
    for n in (1, 2, 4, 8):
        src = os.path.join(pr, f"{tag}_bench_n{n}.json")
        if not os.path.exists(src):
            continue
        d = json.load(open(src)); e = d["e2e"]; a = d.get("also", {})
        lc, ts = a.get("layercam_3680", {}), a.get("trainstep", {})
        f.write(f"{n} | {d['value']:.1f} | {e['value']:.2f} | {e['variants']['f32_images_modules']['value']:.2f} | "
                f"{e['variants']['bf16_logits_u8_images']['value']:.2f} | {lc.get('value', float('nan')):.0f} | {lc.get('ms_per_pass', float('nan')):.2f} | "
                f"{(lc.get('roofline') or {}).get('frac', float('nan')):.3f} | {ts.get('value', float('nan')):.0f} | {ts.get('ms_per_step', float('nan')):.2f} | "
                f"{(ts.get('stage_share') or {}).get('loss_fwd_bwd_ms', float('nan')):.3f}\n")
for src, dst in (("r2_trace_stream_c.txt", "f_trace_stream.txt"), ("r2_ccl_bench.txt", "i_keep_largest_bench.txt"),
                 ("r2_refine.txt", "e_refine_native.txt"), ("r2_sweep_layercam.txt", "j_sweep_layercam_config5.txt"),
                 ("r2_layercam_small.txt", "h_layercam_small_batch.txt")):
    if os.path.exists(os.path.join(go, src)):
        shutil.copy(os.path.join(go, src), os.path.join(pr, f"{tag}_{dst}"))
print(open(os.path.join(pr, f"{tag}_a_scale.txt")).read())

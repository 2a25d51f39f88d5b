"""BASELINE config 5: multi-layer LayerCAM fusion (4 backbone stages) HBM-roofline sweep over batch and resolution.
Prints one line per (S, B): masks/s, achieved GB/s (algorithmic bytes / device time) and the fraction of the measured
HBM peak.  Synthetic ResNet-50 stage shapes: layer1 256x(S/4)^2, layer2 512x(S/8)^2, layer3 1024x(S/16)^2,
layer4 2048x(S/16)^2 (dilated), fp32.  Usage: python scripts/sweep_layercam.py > gpurun_out/sweep_layercam.txt"""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from weaklysuperviseddl_b200 import _native

lib = _native.lib()
dev = torch.device("cuda", 0)
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6650.0
print(f"# LayerCAM -> normalise -> upsample -> fuse -> threshold, 4 stages, fp32; HBM peak {peak:.0f} GB/s")
print("# S  B  layers  ms/call  masks/s  GB/s  frac")
for S in (224, 512, 1024):
    full = [(256, S // 4, S // 4), (512, S // 8, S // 8), (1024, S // 16, S // 16), (2048, S // 16, S // 16)]
    for layers, tag in ((full[2:], "l3+l4"), (full, "l1-l4")):
        per_image = sum(2 * C * h * w * 4 for (C, h, w) in layers) + S * S
        for B in (1, 2, 4, 8, 16, 32, 64):
            if per_image * B * 2 > 60e9:
                continue
            gen = torch.Generator(device=dev).manual_seed(S + B)
            sets = []
            nsets = 2 if per_image * B < 2e9 else 1
            for _ in range(max(nsets, 2 if per_image * B * 2 < 100e6 * 4 else nsets)):
                acts = [torch.randn(B, C, h, w, device=dev, generator=gen).relu_() for (C, h, w) in layers]
                grads = [torch.randn(B, C, h, w, device=dev, generator=gen).mul_(1e-3) for (C, h, w) in layers]
                sets.append((acts, grads))
            n = len(layers)
            IntArr, PtrArr = ctypes.c_int * n, ctypes.c_void_p * n
            Cs, hs, ws_ = IntArr(*[l[0] for l in layers]), IntArr(*[l[1] for l in layers]), IntArr(*[l[2] for l in layers])
            nws = lib.wsdl_layercam_workspace_bytes(Cs, hs, ws_, n, B, 0)
            wsb = torch.empty(nws, dtype=torch.uint8, device=dev)
            mask = torch.empty(B, S, S, dtype=torch.uint8, device=dev)
            flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
            sp = torch.cuda.current_stream(dev).cuda_stream

            def one(i):
                acts, grads = sets[i % len(sets)]
                rc = lib.wsdl_layercam_fused(PtrArr(*[t.data_ptr() for t in acts]), PtrArr(*[t.data_ptr() for t in grads]),
                                             Cs, hs, ws_, n, B, 0, S, S, 1.0, 0, 0.3, 1e-6, None, mask.data_ptr(), None,
                                             wsb.data_ptr(), nws, sp)
                assert rc == 0, rc
            for i in range(3):
                one(i)
            torch.cuda.synchronize()
            small = per_image * B * len(sets) < 200e6  # would live in L2: flush between calls, time each call alone
            reps = 10
            tot = 0.0
            for i in range(reps):
                if small:
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); one(i); e1.record(); torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            ms = tot / reps
            gbs = per_image * B / (ms * 1e-3) / 1e9
            print(f"{S:5d} {B:3d} {tag:6s} {ms:9.4f} {B / (ms * 1e-3):12.0f} {gbs:8.0f} {gbs / peak:6.3f}")
            del sets, acts, grads
            torch.cuda.empty_cache()

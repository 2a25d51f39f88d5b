"""refine_pseudo_mask per image (the reference's loop on the fused kernel, 2 host syncs per step)
vs refine_pseudo_masks_batched (no host sync in the loop), 16 images of 256x256, 10 steps (CutLoss.py:803-806)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import smooth_images
from oracle.make_golden import FixedLogitsNet
from weaklysuperviseddl_b200.AlternatingDirectionCutLoss import refine_pseudo_mask, refine_pseudo_masks_batched

torch.manual_seed(0)
gen = torch.Generator().manual_seed(0)
N, H, W, steps = 16, 256, 256, 10
seg = FixedLogitsNet().eval().cuda()
images = smooth_images(gen, N, H, W).cuda()
masks = ((torch.rand(N, H, W, generator=gen) > 0.5).long() * 255).cuda()
for fn_name in ("per-image loop", "batched"):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if fn_name == "batched":
            out = refine_pseudo_masks_batched(seg, images, masks, 0.1, 0.3, 1e-4, steps)
        else:
            out = torch.stack([refine_pseudo_mask(seg, images[i], masks[i], 0.1, 0.3, 1e-4, steps) for i in range(N)])
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{fn_name:15s}: {dt*1e3:8.2f} ms for {N} images x {steps} steps  ({N/dt:8.1f} images/s)  fg fraction {out.mean().item():.4f}")

# the step loop alone (one CUDA graph: 2 x steps + 1 launches), CUDA events around 20 replays
from weaklysuperviseddl_b200 import AlternatingDirectionCutLoss as ADC
ref = next(r for k, r in ADC._REFINERS.items() if k[1] == N)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20):
    ref.graph.replay()
e1.record(); torch.cuda.synchronize()
print(f"captured step loop  : {e0.elapsed_time(e1) / 20:8.3f} ms per {steps}-step refinement of {N} images ({2 * steps + 1} launches + 2 memsets)")

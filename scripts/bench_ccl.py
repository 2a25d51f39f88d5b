"""keep_largest (GPU CCL) throughput on two kinds of masks: `cam` = a thresholded bilinear up-sampling of a 16x16 random
field (what LayerCAM -> threshold produces: smooth outlines, one or two runs per row), `noise` = a thresholded 33x33 box
filter of white noise (blobs with ragged outlines: many short runs per row along every boundary; the harder case, and the
one profiles/r01_g / r02_d quote)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from weaklysuperviseddl_b200 import functional as WF
dev = "cuda"
for kind in ("noise", "cam"):
    for (B, S) in ((32, 224), (128, 512), (16, 1024)):
        g = torch.Generator(device=dev).manual_seed(0)
        if kind == "noise":
            f = torch.nn.functional.avg_pool2d(torch.rand(B, 1, S + 32, S + 32, device=dev, generator=g), 33, stride=1)[:, 0]
            m = (f > f.mean()).to(torch.uint8).contiguous()
        else:
            f = torch.nn.functional.interpolate(torch.rand(B, 1, 16, 16, device=dev, generator=g), size=(S, S), mode="bilinear",
                                                align_corners=False)[:, 0]
            m = (f > 0.5).to(torch.uint8).contiguous()
        for _ in range(3):
            out = WF.keep_largest(m)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            out = WF.keep_largest(m)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{kind:5s} B={B} S={S}: {ms:.3f} ms per call, {B/ms*1e3:.0f} masks/s, {B*S*S/ms/1e6:.2f} Gpix/s, "
              f"{2*B*S*S/ms/1e6:.0f} GB/s of mask in + result out, fg {m.float().mean().item():.2f}, kept {out.float().mean().item():.2f}")

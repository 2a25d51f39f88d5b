import sys, torch
sys.path.insert(0, ".")
from weaklysuperviseddl_b200 import functional as WF
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (B, S, layers) in ((1, 224, [(1024, 14, 14), (2048, 14, 14)]), (8, 224, [(1024, 14, 14), (2048, 14, 14)]), (16, 224, [(1024, 14, 14), (2048, 14, 14)]),
                       (1, 512, [(1024, 32, 32), (2048, 32, 32)]), (8, 512, [(1024, 32, 32), (2048, 32, 32)])):
    g = torch.Generator(device=dev).manual_seed(0)
    acts = [torch.randn(B, C, h, w, device=dev, generator=g).relu_() for (C, h, w) in layers]
    grads = [torch.randn(B, C, h, w, device=dev, generator=g).mul_(1e-3) for (C, h, w) in layers]
    ws = torch.empty(WF.layercam_workspace_bytes(layers, B), dtype=torch.uint8, device=dev)
    mask = torch.empty(B, S, S, dtype=torch.uint8, device=dev)
    for _ in range(3):
        WF.layercam_fused(acts, grads, (S, S), thresh=0.3, want_cam=False, mask_out=mask, workspace=ws)
    tot = 0.0
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); WF.layercam_fused(acts, grads, (S, S), thresh=0.3, want_cam=False, mask_out=mask, workspace=ws); e1.record()
        torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    nbytes = sum(2 * C * h * w * 4 for (C, h, w) in layers) * B + B * S * S
    print(f"B={B} S={S}: {tot / 10 * 1e3:7.1f} us per call (cold L2), {nbytes / (tot / 10 * 1e-3) / 1e9:6.0f} GB/s")

"""Low-batch LayerCAM -> mask calls (the reference's own call pattern: B = 1 per call, LayerCAM.py:38; BASELINE config 1:
B = 8): device time per call, cold L2 (rotating input sets > 126 MB), calls captured in a CUDA graph so that the host's
launch overhead is not what is measured."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from weaklysuperviseddl_b200 import functional as WF
dev = "cuda"
for (B, S, layers) in ((1, 224, [(1024, 14, 14), (2048, 14, 14)]), (8, 224, [(1024, 14, 14), (2048, 14, 14)]),
                       (16, 224, [(1024, 14, 14), (2048, 14, 14)]), (1, 512, [(1024, 32, 32), (2048, 32, 32)]),
                       (8, 512, [(1024, 32, 32), (2048, 32, 32)])):
    per_set = sum(2 * C * h * w * 4 for (C, h, w) in layers) * B
    n_sets = max(2, int(160e6 // per_set) + 1)
    g = torch.Generator(device=dev).manual_seed(0)
    sets = [([torch.randn(B, C, h, w, device=dev, generator=g).relu_() for (C, h, w) in layers],
             [torch.randn(B, C, h, w, device=dev, generator=g).mul_(1e-3) for (C, h, w) in layers]) for _ in range(n_sets)]
    ws = torch.empty(WF.layercam_workspace_bytes(layers, B), dtype=torch.uint8, device=dev)
    mask = torch.empty(B, S, S, dtype=torch.uint8, device=dev)
    def calls():
        for a, gr in sets:
            WF.layercam_fused(a, gr, (S, S), thresh=0.3, want_cam=False, mask_out=mask, workspace=ws)
    calls(); torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        calls()
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(3, 400 // n_sets)
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * n_sets)
    nbytes = per_set + B * S * S
    print(f"B={B:2d} S={S}: {us:7.1f} us per call, {B / us * 1e6:9.0f} masks/s, {nbytes / us / 1e3:6.0f} GB/s "
          f"({n_sets} rotating input sets, {reps * n_sets} calls)")

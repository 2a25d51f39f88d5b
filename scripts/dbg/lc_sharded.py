import sys, time, torch
sys.path.insert(0, ".")
import bench
from weaklysuperviseddl_b200.PsuedoMasks import generate_pseudo_masks_sharded
from weaklysuperviseddl_b200 import functional as WF
dev = torch.device("cuda", 0)
chunk, S, layers = 128, 512, bench.LCAM["layers"]
gen = torch.Generator(device=dev).manual_seed(1)
sets = []
for _ in range(2):
    sets.append(([torch.randn(chunk, C, h, w, device=dev, generator=gen).relu_() for (C, h, w) in layers],
                 [torch.randn(chunk, C, h, w, device=dev, generator=gen).mul_(1e-3) for (C, h, w) in layers]))
calls = {"n": 0}
def hooks(idx):
    a, g = sets[calls["n"] % 2]; calls["n"] += 1
    n = len(idx); return [x[:n] for x in a], [x[:n] for x in g]
keep = {}
for kw in (dict(streams=2, count_foreground=True), dict(streams=2, count_foreground=False), dict(streams=1, count_foreground=False), dict(streams=1, count_foreground=True), dict(streams=2, count_foreground=False, keep_largest_masks=True)):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = generate_pseudo_masks_sharded(hooks, 3680, (S, S), chunk=chunk, rank=0, world=1, sink=lambda i, m: keep.__setitem__("m", m), **kw)
        torch.cuda.synchronize(); t1 = time.perf_counter()
    print(kw, f"{(t1 - t0) * 1e3:.1f} ms per pass", r["counters"])
# raw loop of the fused call only
torch.cuda.synchronize(); t0 = time.perf_counter()
for c in range(29):
    a, g = sets[c % 2]
    WF.layercam_fused(a, g, (S, S), thresh=0.3, want_cam=False)
torch.cuda.synchronize(); print(f"29 plain layercam_fused calls: {(time.perf_counter() - t0) * 1e3:.1f} ms")
t0 = time.perf_counter()
for c in range(29):
    hooks(list(range(128)))
print(f"29 hook_source calls: {(time.perf_counter() - t0) * 1e3:.1f} ms")

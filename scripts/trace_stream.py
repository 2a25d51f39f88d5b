"""Debug aid: per-group, per-tile phase timestamps of weak_loss_stream_kernel.  Needs a library built with
WSDL_NVCC_EXTRA=-DWSDL_PS_TRACE (scripts/gpu_trace.sh swaps it in)."""
import ctypes, numpy as np, torch, sys
from weaklysuperviseddl_b200 import functional as WF, _native
sys.path.insert(0, "tests")
from helpers import smooth_images
gen = torch.Generator().manual_seed(1)
B = 32
logits = torch.randn(B, 2, 224, 224, generator=gen).cuda()
img = smooth_images(gen, B, 224, 224).cuda()
lib = _native.lib()
lib.wsdl_sg_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.zero_()
    torch.cuda.synchronize()
    WF.pairwise_dual_loss_and_grad(logits, img)
torch.cuda.synchronize()
N = 148 * 3 * 8 * 8
buf = (ctypes.c_ulonglong * N)()
lib.wsdl_sg_trace_read(buf, N)
a = np.frombuffer(buf, dtype=np.uint64).reshape(148, 3, 8, 8).astype(np.int64)
t0 = a[a > 0].min()
names = ["ticket", "tile landed", "converted", "marched", "heads barrier", "heads emitted", "band done", "tile end"]
valid = a[:, :, :, 7] > 0
# keep only entries of the last launch (timestamps within 200 us of the max)
recent = a[:, :, :, 0] >= a.max() - 200000
ok = valid & recent
t0 = a[:, :, :, 0][ok].min()
T = (a - t0) / 1000.0
print(f"tiles traced {ok.sum()}, span {T[:, :, :, 7][ok].max():.1f} us")
for n in range(4):
    sel = ok[:, :, n]
    if not sel.any():
        continue
    print(f"-- tile #{n} of a group: {sel.sum()} tiles; phase durations med / max (us), median end time")
    for i in range(1, 8):
        d = (T[:, :, n, i] - T[:, :, n, i - 1])[sel]
        print(f"   {names[i]:14s} {np.median(d):6.2f} {d.max():6.2f}   @ {np.median(T[:, :, n, i][sel]):6.2f}")
ends = np.where(ok, T[:, :, :, 7], 0).max(axis=(1, 2))
print(f"per-SM finish: min {ends.min():.1f} med {np.median(ends):.1f} max {ends.max():.1f}; tiles per CTA min {ok.sum(axis=(1,2)).min()} max {ok.sum(axis=(1,2)).max()}")
c = int(np.argmax(ends))
print(f"CTA {c} timeline (us): group: [ticket | landed | converted | marched | end] per tile")
for g in range(3):
    print(f"  g{g}: " + "  ".join("[" + " ".join(f"{T[c, g, n, i]:5.1f}" for i in (0, 1, 2, 3, 7)) + "]" for n in range(8) if ok[c, g, n]))

"""Debug aid: timestamps of pairwise_pipe_kernel per CTA / role / tile (library built with -DWSDL_PS_TRACE)."""
import ctypes, numpy as np, torch, sys
from weaklysuperviseddl_b200 import functional as WF, _native
sys.path.insert(0, "tests")
from helpers import smooth_images
gen = torch.Generator().manual_seed(1)
B = 32
logits = torch.randn(B, 2, 224, 224, generator=gen).cuda()
img = smooth_images(gen, B, 224, 224).cuda()
probs = torch.softmax(logits, 1)
lib = _native.lib()
lib.wsdl_pp_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
N = 512 * 2 * 8 * 4
for name, args in (("cut", (logits, img, 5, 0.05, None, True, True, False)), ("boundary", (probs, img, 5, 0.1, 5.0, False, False, True))):
    for _ in range(3):
        flush.zero_()
        WF.pairwise_loss_and_grad(*args)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * N)()
    lib.wsdl_pp_trace_read(buf, N)
    a = np.frombuffer(buf, dtype=np.uint64).reshape(512, 2, 8, 4).astype(np.int64)
    live = a[:, 0, 0, 0] > 0
    a = a[live]
    t0 = a[a > 0].min()
    T = np.where(a > 0, (a - t0) / 1000.0, np.nan)
    print(f"== {name}: {len(a)} CTAs, span {np.nanmax(T):.1f} us")
    mn = ["ready", "marched-loop", "band", "handed"]; hn = ["landed", "converted", "-", "-"]
    for j in range(5):
        if np.isnan(T[:, 0, j, 0]).all():
            break
        m = T[:, 0, j, :]; h = T[:, 1, j, :]
        print(f"  tile {j}: M " + " ".join(f"{mn[k]} {np.nanmedian(m[:,k]):5.1f}" for k in range(4)) +
              "   | H " + " ".join(f"{hn[k]} {np.nanmedian(h[:,k]):5.1f}" for k in range(4)) + f"   (n={np.sum(~np.isnan(m[:,0]))})")
    c = 0
    print("  CTA 0 M:", " | ".join(" ".join(f"{T[c,0,j,k]:5.1f}" for k in range(4)) for j in range(5)))
    print("  CTA 0 H:", " | ".join(" ".join(f"{T[c,1,j,k]:5.1f}" for k in range(4)) for j in range(5)))

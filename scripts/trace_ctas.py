"""Debug aid: per-CTA phase timestamps of pairwise_sym_kernel.  Needs a library built with
WSDL_NVCC_EXTRA=-DWSDL_PS_TRACE (python -m weaklysuperviseddl_b200.build --force)."""
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
lib.wsdl_ps_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
N = 8192
for name, args in (("cut", (logits, img, 5, 0.05, None, True, True, False)), ("boundary", (probs, img, 5, 0.1, 5.0, False, False, True))):
    for _ in range(3):
        flush.zero_()
        WF.pairwise_loss_and_grad(*args)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (N * 8))()
    lib.wsdl_ps_trace_read(buf, N)
    a = np.frombuffer(buf, dtype=np.uint64).reshape(N, 8).astype(np.int64)
    a = a[a[:, 4] > 0]
    a = a[a[:, 0] >= a[:, 0].max() - 200000]  # this launch only (stale rows of a larger earlier grid are older)
    t0 = a[:, 0].min()
    T = a[:, :5] - t0
    sm = a[:, 7]
    print(f"== {name}: {len(a)} CTAs, kernel span {T[:,4].max()/1000:.1f} us; last start {T[:,0].max()/1000:.2f} us")
    names = ["start", "staged", "xfix", "march", "heads+partial"]
    for i in range(1, 5):
        d = (T[:, i] - T[:, i - 1]) / 1000.0
        print(f"  {names[i]:14s}: phase dur min/med/mean/max {d.min():5.2f} {np.median(d):5.2f} {d.mean():5.2f} {d.max():5.2f}")
    d = (T[:, 4] - T[:, 0]) / 1000.0
    print(f"  CTA lifetime  : min/med/mean/max {d.min():5.2f} {np.median(d):5.2f} {d.mean():5.2f} {d.max():5.2f}")
    per_sm_end = np.zeros(148); per_sm_n = np.zeros(148, dtype=int)
    for i in range(len(a)):
        per_sm_end[sm[i]] = max(per_sm_end[sm[i]], T[i, 4]); per_sm_n[sm[i]] += 1
    print(f"  per-SM finish us: min {per_sm_end.min()/1000:.1f} med {np.median(per_sm_end)/1000:.1f} max {per_sm_end.max()/1000:.1f}; CTAs per SM min {per_sm_n.min()} max {per_sm_n.max()}")
    order = np.argsort(T[:, 0])
    print("  start times (us) of every 64th CTA in start order:", " ".join(f"{T[i,0]/1000:.1f}" for i in order[::64]))

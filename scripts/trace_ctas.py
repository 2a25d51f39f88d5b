"""Debug aid: per-CTA start/end times of pairwise_sym_kernel (library built with -DWSDL_PS_TRACE)."""
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
for name, args in (("cut", (logits, img, 5, 0.05, None, True, True, False)), ("boundary", (probs, img, 5, 0.1, 5.0, False, False, True))):
    for _ in range(3):
        WF.pairwise_loss_and_grad(*args)
    torch.cuda.synchronize()
    n = 600
    buf = (ctypes.c_ulonglong * (n * 4))()
    lib.wsdl_ps_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
    rc = lib.wsdl_ps_trace_read(buf, n)
    a = np.frombuffer(buf, dtype=np.uint64).reshape(n, 4).astype(np.int64)
    a = a[a[:, 1] > 0]
    t0 = a[:, 0].min()
    st, en, sm = a[:, 0] - t0, a[:, 1] - t0, a[:, 2]
    print(name, "ctas", len(a), "start min/med/max ns", st.min(), int(np.median(st)), st.max(), "| end min/med/max", en.min(), int(np.median(en)), en.max(),
          "| dur min/med/max", (en - st).min(), int(np.median(en - st)), (en - st).max())
    per_sm = np.bincount(sm, minlength=148)
    print("  ctas per SM: min", per_sm[per_sm > 0].min(), "max", per_sm.max(), "SMs used", (per_sm > 0).sum())
    order = np.argsort(en)
    for i in list(order[:3]) + list(order[-6:]):
        print("   cta", i, "sm", sm[i], "start", st[i], "end", en[i], "dur", en[i] - st[i])
    # duration by cta position within image column
    d = en - st
    idx = np.arange(len(a))
    print("  mean dur by (cta % 4.571 column position) top-slow ctas idx:", sorted(order[-20:].tolist()))
    if name == "cut":
        print("  dur(us) of ctas 0..59:", " ".join(f"{x/1000:.0f}" for x in d[:60]))
        print("  corr kcycles of ctas 0..59:", " ".join(f"{x/1000:.1f}" for x in a[:60,3]))
        # per-SM sum of durations
        mx = np.zeros(148); 
        for i in range(len(a)): mx[sm[i]] = max(mx[sm[i]], en[i])
        print("  per-SM finish us: min", mx.min()/1000, "med", np.median(mx)/1000, "max", mx.max()/1000)

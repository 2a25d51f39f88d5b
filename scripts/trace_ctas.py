"""Debug aid: per-warp phase timestamps of pairwise_sym_kernel.  Needs a library built with
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
N = 4096
names = ["start", "tma landed", "early rows", "all rows", "neighbour ok", "step0", "step1", "step2", "step3", "march end (dual S=5: + step4)",
         "barrier", "heads", "band pass", "end"]
cases = [("cut", (logits, img, 5, 0.05, None, True, True, False)), ("boundary", (probs, img, 5, 0.1, 5.0, False, False, True)),
         ("dual", None)]
if len(sys.argv) > 1:
    cases = [c for c in cases if c[0] in sys.argv[1:]]
for name, args in cases:
    for _ in range(3):
        flush.zero_()
        if args is None:
            WF.pairwise_dual_loss_and_grad(logits, img)
        else:
            WF.pairwise_loss_and_grad(*args)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (N * 64))()
    lib.wsdl_ps_trace_read(buf, N)
    a = np.frombuffer(buf, dtype=np.uint64).reshape(N, 4, 16).astype(np.int64)
    a = a[a[:, 0, 13] > 0]
    a = a[a[:, 0, 0] >= a[:, 0, 0].max() - 200000]
    t0 = a[:, :, 0].min()
    T = a[:, :, :14] - t0            # [cta][warp][slot]
    sm = a[:, 0, 15]
    print(f"== {name}: {len(a)} CTAs, kernel span {T[:,:,13].max()/1000:.1f} us; last start {T[:,:,0].max()/1000:.2f} us")
    first = T[:, 0, 0] < 2000        # first-wave CTAs
    for label, sel in (("first wave", first), ("later", ~first)):
        if not sel.any():
            continue
        print(f"  -- {label}: {sel.sum()} CTAs; per-warp phase durations (us) med / mean / max, then median end time")
        for i in range(1, 14):
            d = (T[sel][:, :, i] - T[sel][:, :, i - 1]) / 1000.0
            ok = (T[sel][:, :, i] > 0) & (T[sel][:, :, i - 1] > 0)
            if not ok.any():
                continue
            d = d[ok]
            print(f"     {names[i]:13s} {np.median(d):6.2f} {d.mean():6.2f} {d.max():6.2f}   @ {np.median(T[sel][:, :, i][ok])/1000:6.2f}")
    life = (T[:, :, 13].max(axis=1) - T[:, :, 0].min(axis=1)) / 1000.0
    print(f"  CTA lifetime  : min/med/mean/max {life.min():5.2f} {np.median(life):5.2f} {life.mean():5.2f} {life.max():5.2f}")
    per_sm_end = np.zeros(148); per_sm_n = np.zeros(148, dtype=int)
    for i in range(len(a)):
        per_sm_end[sm[i]] = max(per_sm_end[sm[i]], T[i, :, 13].max()); per_sm_n[sm[i]] += 1
    print(f"  per-SM finish us: min {per_sm_end.min()/1000:.1f} med {np.median(per_sm_end)/1000:.1f} max {per_sm_end.max()/1000:.1f}; CTAs per SM min {per_sm_n.min()} max {per_sm_n.max()}")
    # one SM's timeline
    k = int(np.argmax(per_sm_end))
    rows = [i for i in range(len(a)) if sm[i] == k]
    rows.sort(key=lambda i: T[i, 0, 0])
    print(f"  SM {k} timeline (us): cta: start | tma | rows | march begin..end (warp 0..3) | end")
    for i in rows:
        print(f"    {T[i,0,0]/1000:5.1f} | {T[i,0,1]/1000:5.1f} | {T[i,:,3].max()/1000:5.1f} | " +
              " ".join(f"{T[i,w,4]/1000:4.1f}..{T[i,w,9]/1000:4.1f}" for w in range(4)) + f" | {T[i,:,13].max()/1000:5.1f}")

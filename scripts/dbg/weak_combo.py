import sys, torch, subprocess, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
if len(sys.argv) > 1:
    from helpers import smooth_images
    from weaklysuperviseddl_b200 import functional as WF
    ld, idt, lab = sys.argv[1:4]
    gen = torch.Generator().manual_seed(9)
    B, H, W = 3, 96, 128
    logits = torch.randn(B, 2, H, W, generator=gen) * 2
    img = smooth_images(gen, B, H, W)
    labels = (torch.rand(B, H, W, generator=gen) > 0.5).long()
    lg = logits.to(torch.bfloat16) if ld == "bf16" else logits
    im = (img * 255).round().to(torch.uint8) if idt == "u8" else img
    lb = None if lab == "none" else (labels.to(torch.uint8) if lab == "u8" else labels)
    out = WF.weak_loss_and_grad(lg.cuda(), im.cuda(), lb.cuda() if lb is not None else None)
    torch.cuda.synchronize()
    print("OK", ld, idt, lab, out[0].item())
else:
    for ld in ("f32", "bf16"):
        for idt in ("f32", "u8"):
            for lab in ("none", "u8", "i64"):
                r = subprocess.run([sys.executable, __file__, ld, idt, lab], capture_output=True, text=True, timeout=120,
                                   env=dict(os.environ, CUDA_LAUNCH_BLOCKING="1"))
                print(ld, idt, lab, "->", (r.stdout.strip().splitlines() or ["-"])[-1], "|", (r.stderr.strip().splitlines() or ["-"])[-1][:150])

"""Small fixed workload for ncu captures of the pairwise kernels at BASELINE configs[1] size (32x2x224x224): the
two-regulariser launch (pairwise_dual_kernel), the one-launch training loss (weak_loss_stream_kernel<1>: f32 + labels;
<2>: bf16 logits + u8 images + labels) and keep_largest at 128 x 512^2."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import smooth_images
from weaklysuperviseddl_b200 import functional as WF
gen = torch.Generator().manual_seed(1)
B, H, W = 32, 224, 224
logits = torch.randn(B, 2, H, W, generator=gen).cuda()
img = smooth_images(gen, B, H, W).cuda()
labels = (torch.rand(B, H, W, generator=gen) > 0.5).to(torch.uint8).cuda()
img8 = (img * 255).round().to(torch.uint8)
go_c, go_b = torch.full((1,), 0.1, device="cuda"), torch.full((B,), 0.5 / B, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for rep in range(3):
    flush.zero_()
    WF.pairwise_dual_loss_and_grad(logits, img, grad_out_cut=go_c, grad_out_bnd=go_b)
    flush.zero_()
    WF.weak_loss_and_grad(logits, img, None, 0.0, go_c, go_b)          # <0>: f32, no labels
    flush.zero_()
    WF.weak_loss_and_grad(logits, img, labels, 1.0, go_c, go_b)        # <1>: f32 + labels
    flush.zero_()
    WF.weak_loss_and_grad(logits.to(torch.bfloat16), img, labels, 1.0, go_c, go_b)  # <3>: the bf16 training step
    flush.zero_()
    WF.weak_loss_and_grad(logits.to(torch.bfloat16), img8, labels, 1.0, go_c, go_b)
g = torch.Generator(device="cuda").manual_seed(0)
f = torch.nn.functional.avg_pool2d(torch.rand(128, 1, 512 + 32, 512 + 32, device="cuda", generator=g), 33, stride=1)[:, 0]
m = (f > f.mean()).to(torch.uint8).contiguous()
fc = torch.nn.functional.interpolate(torch.rand(128, 1, 16, 16, device="cuda", generator=g), size=(512, 512), mode="bilinear")[:, 0]
mc = (fc > 0.5).to(torch.uint8).contiguous()  # CAM-like masks: smooth outlines
for _ in range(2):
    WF.keep_largest(m)
    WF.keep_largest(mc)
torch.cuda.synchronize()
print("ok")

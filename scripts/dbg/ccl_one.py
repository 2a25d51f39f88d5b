import os, sys, torch
sys.path.insert(0, ".")
from weaklysuperviseddl_b200 import functional as WF
B, S = 128, 512
g = torch.Generator(device="cuda").manual_seed(0)
f = torch.nn.functional.avg_pool2d(torch.rand(B, 1, S + 32, S + 32, device="cuda", generator=g), 33, stride=1)[:, 0]
m = (f > f.mean()).to(torch.uint8).contiguous()
for _ in range(3):
    out = WF.keep_largest(m)
torch.cuda.synchronize()
print("ok", out.float().mean().item())

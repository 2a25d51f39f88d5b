import os, sys, subprocess
code = r'''
import os, torch
from weaklysuperviseddl_b200 import functional as WF
torch.manual_seed(0)
v = torch.randn(2,2,40,36, device="cuda"); im = torch.rand(2,3,40,36, device="cuda")
try:
    l, g = WF.pairwise_loss_and_grad(v, im)
    torch.cuda.synchronize()
    print("OK", l.item(), g.abs().max().item())
except Exception as e:
    print("FAIL", str(e)[:120].replace("\n"," "))
'''
for env in ({"WSDL_PAIRWISE_NO_TMA":"1"}, {"WSDL_PAIRWISE_DEBUG":"1"}, {"WSDL_PAIRWISE_DEBUG":"3"}, {"WSDL_PAIRWISE_DEBUG":"2"}, {}):
    e = dict(os.environ); e.update(env); e["CUDA_LAUNCH_BLOCKING"]="1"
    r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, timeout=120)
    print(env, r.stdout.strip()[-200:], r.stderr.strip()[-200:] if r.returncode else "")

"""GPU: the composed training loss of BASELINE config 4 (CE on pseudo-labels + cut + boundary), single GPU against
the oracle composition, and a 2-rank NCCL DistributedDataParallel step when two GPUs are visible."""
import os
import socket
import sys

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from helpers import smooth_images
from oracle import wsdl_oracle as O

pytestmark = pytest.mark.gpu


class TinySeg(nn.Module):
    def __init__(self):
        super().__init__()
        self.c1 = nn.Conv2d(3, 8, 3, padding=1)
        self.c2 = nn.Conv2d(8, 2, 3, padding=1)

    def forward(self, x):
        return {"out": self.c2(torch.relu(self.c1(x)))}


def _oracle_total(logits, img, labels, lc, lb):
    x = logits.detach().clone().requires_grad_(True)
    ce = F.cross_entropy(x, labels)
    cut = O.cut_loss(x, img, sigma_color=0.05, window_size=5)
    p = torch.softmax(x, dim=1)
    bnd = torch.stack([O.boundary_loss(p[b], img[b], sigma_color=0.1, sigma_space=5, window_size=5)
                       for b in range(x.shape[0])]).mean()
    total = ce + lc * cut + lb * bnd
    total.backward()
    return total.detach(), x.grad


def test_composed_loss_matches_oracle():
    from weaklysuperviseddl_b200 import WeakSupervisionLoss

    gen = torch.Generator().manual_seed(8)
    B, H, W = 3, 40, 52
    logits = torch.randn(B, 2, H, W, generator=gen)
    img = smooth_images(gen, B, H, W)
    labels = (torch.rand(B, H, W, generator=gen) > 0.5).long()
    ref_total, ref_grad = _oracle_total(logits.double(), img.double(), labels, 0.1, 0.5)
    x = logits.cuda().requires_grad_(True)
    total, parts = WeakSupervisionLoss()(x, img.cuda(), labels.cuda())
    total.backward()
    assert abs(total.item() - ref_total.item()) <= 1e-5 * abs(ref_total.item())
    assert (x.grad.cpu().double() - ref_grad).abs().max().item() <= 1e-5 * ref_grad.abs().max().item()
    assert set(parts) == {"ce", "cut", "boundary"}


def test_train_step_bf16_autocast_reduces_loss():
    from weaklysuperviseddl_b200 import WeakSupervisionLoss

    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(9)
    net = TinySeg().cuda()
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    crit = WeakSupervisionLoss()
    img = smooth_images(gen, 4, 64, 64).cuda()
    labels = (img.mean(1) > img.mean()).long()
    losses = []
    for _ in range(8):
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = net(img)["out"]
        total, _ = crit(out, img, labels)
        total.backward()
        opt.step()
        losses.append(total.item())
    assert losses[-1] < losses[0]


def _ddp_worker(rank, world, port, ret):
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from weaklysuperviseddl_b200 import WeakSupervisionLoss

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        torch.manual_seed(0)
        net = DDP(TinySeg().cuda(rank), device_ids=[rank])
        gen = torch.Generator().manual_seed(10)
        img_all = smooth_images(gen, 2 * world, 48, 48)
        labels_all = (img_all.mean(1) > img_all.mean()).long()
        img, labels = img_all[rank::world].cuda(rank), labels_all[rank::world].cuda(rank)  # image sharding
        total, _ = WeakSupervisionLoss()(net(img)["out"], img, labels)
        total.backward()
        g = torch.cat([p.grad.flatten() for p in net.parameters()])
        gathered = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(gathered, g)
        if rank == 0:
            ret["same"] = all(torch.equal(gathered[0], t) for t in gathered)
            ret["norm"] = g.norm().item()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_ddp_nccl_step_two_gpus():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ddp_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["same"] and ret["norm"] > 0

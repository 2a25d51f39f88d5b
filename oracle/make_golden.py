"""TEST INFRASTRUCTURE ONLY -- regenerates tests/golden/*.npz from the REAL reference.

Run in the authoring container (needs /root/reference):

    python oracle/make_golden.py

The reference has no golden vectors of its own, so these
fixtures are outputs of the reference functions themselves on seeded inputs.
They travel to the GPU box, where /root/reference does not exist.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


class TinyCAMNet(nn.Module):
    """Small stand-in with the FrozenResNetCAM contract (ClassificationModel.py:18-41):
    hookable attributes layer2/3/4, forward -> (logits, [f2, f3, f4]).  Some
    activations are negative on purpose (no ReLU after layer3) to pin the
    reference's relu(grad*act) form (LayerCAM.py:57)."""

    def __init__(self, num_classes=5):
        super().__init__()
        self.layer0 = nn.Sequential(nn.Conv2d(3, 8, 3, 2, 1), nn.ReLU())
        self.layer1 = nn.Sequential(nn.Conv2d(8, 12, 3, 2, 1), nn.ReLU())
        self.layer2 = nn.Sequential(nn.Conv2d(12, 16, 3, 2, 1), nn.ReLU())
        self.layer3 = nn.Sequential(nn.Conv2d(16, 24, 3, 2, 1))
        self.layer4 = nn.Sequential(nn.Conv2d(24, 40, 3, 1, 1), nn.ReLU())
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(40, num_classes)

    def forward(self, x):
        f2 = self.layer2(self.layer1(self.layer0(x)))
        f3 = self.layer3(f2)
        f4 = self.layer4(f3)
        return self.fc(self.avgpool(f4).flatten(1)), [f2, f3, f4]


class FixedLogitsNet(nn.Module):
    """model(x)['out'] contract of torchvision segmentation nets (SegmentationModel.py:86-88)."""

    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(3, 2, 5, padding=2)

    def forward(self, x):
        return {"out": self.conv(x) * 4.0}


def smooth_images(gen, B, H, W):
    """box-filtered noise rescaled per image to [0,1]."""
    raw = torch.rand(B, 3, H + 16, W + 16, generator=gen)
    img = torch.nn.functional.avg_pool2d(raw, 9, stride=1)[..., 4 : H + 4, 4 : W + 4]
    lo = img.amin(dim=(1, 2, 3), keepdim=True)
    hi = img.amax(dim=(1, 2, 3), keepdim=True)
    return ((img - lo) / (hi - lo)).contiguous()


def main():
    R = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    gen = torch.Generator().manual_seed(1234)

    # ---- LayerCAM (both generator variants) on a tiny hookable model, 224x224 output ----
    torch.manual_seed(7)
    net = TinyCAMNet().eval()
    state = {k: v.numpy() for k, v in net.state_dict().items()}
    images = torch.rand(2, 3, 96, 80, generator=gen)  # the reference resizes to 224x224 regardless (LayerCAM.py:69)
    labels = torch.tensor([1, 4])
    rec = {"images": images.numpy(), "labels": labels.numpy()}
    for tag, cls, layers in (("main", "LayerCAMGenerator", ["layer3", "layer4"]), ("variant", "LayerCAMGeneratorVariant", ["layer2", "layer3", "layer4"])):
        fresh = TinyCAMNet().eval()
        fresh.load_state_dict(net.state_dict())
        g = R[cls](fresh, layers)
        for alpha in ((1.0, 0.5) if tag == "main" else (1.0, 2.0)):
            cams = []
            for i in range(images.shape[0]):
                if tag == "main":
                    cam = g.generate(images[i].clone(), alpha, class_idx=labels[i : i + 1])
                else:
                    cam = g.generate(images[i].clone(), class_idx=labels[i : i + 1], alpha=alpha)
                cams.append(cam[0].numpy())
                if alpha == 1.0:
                    for n in layers:
                        rec[f"{tag}_act_{n}_{i}"] = g.activations[n].detach().numpy()
                        rec[f"{tag}_grad_{n}_{i}"] = g.gradients[n].detach().numpy()
            rec[f"{tag}_cam_alpha{alpha}"] = np.stack(cams)
        # class_idx=None path (argmax)
        if tag == "main":
            rec["main_cam_argmax"] = g.generate(images[0].clone(), 1.0)[0].numpy()
    np.savez_compressed(os.path.join(OUT, "layercam_tiny.npz"), **rec, **{f"w::{k}": v for k, v in state.items()})

    # ---- pairwise losses ----
    B, C, H, W = 2, 2, 40, 36
    logits = torch.randn(B, C, H, W, generator=gen)
    img = smooth_images(gen, B, H, W)
    rec = {"logits": logits.numpy(), "images": img.numpy()}
    cut = R["LocalNormalizedCutLoss"](sigma_color=0.05, window_size=5)
    x = logits.clone().requires_grad_(True)
    val = cut(x, img)
    val.backward()
    rec["cut_loss"], rec["cut_grad"] = val.detach().numpy(), x.grad.numpy()
    x = logits[0].clone().requires_grad_(True)  # 3-D auto-batch path (CutLoss.py:72-74)
    val = cut(x, img[0])
    val.backward()
    rec["cut3d_loss"], rec["cut3d_grad"] = val.detach().numpy(), x.grad.numpy()
    cut7 = R["LocalNormalizedCutLoss"](sigma_color=0.2, window_size=3)
    x = logits.clone().requires_grad_(True)
    val = cut7(x, img)
    val.backward()
    rec["cut_w3_loss"], rec["cut_w3_grad"] = val.detach().numpy(), x.grad.numpy()
    # three classes
    logits3 = torch.randn(1, 3, 17, 23, generator=gen)
    img3 = smooth_images(gen, 1, 17, 23)
    x = logits3.clone().requires_grad_(True)
    val = cut(x, img3)
    val.backward()
    rec.update(logits3=logits3.numpy(), images3=img3.numpy(), cut_c3_loss=val.detach().numpy(), cut_c3_grad=x.grad.numpy())

    bnd = R["ConstrainToBoundaryLossSingle"](sigma_color=0.1, sigma_space=5, window_size=5)
    probs = torch.softmax(logits, dim=1)
    bl, bg = [], []
    for b in range(B):
        x = probs[b].clone().requires_grad_(True)
        val = bnd(x, img[b])
        val.backward()
        bl.append(val.detach().numpy())
        bg.append(x.grad.numpy())
    rec["boundary_loss"], rec["boundary_grad"] = np.stack(bl), np.stack(bg)
    rec["affinities_batched"] = torch.stack(R["compute_affinities"](img, 0.1, 5, 5), dim=1).squeeze(2).numpy()  # (B,24,H,W)
    rec["affinities_single"] = torch.stack(bnd.compute_affinities_single(img[1], 0.1, 5, 5), dim=0).squeeze(1).numpy()  # (24,H,W)
    np.savez_compressed(os.path.join(OUT, "pairwise.npz"), **rec)

    # ---- refine_pseudo_mask (reference function, fake segmentation net) ----
    torch.manual_seed(11)
    seg = FixedLogitsNet().eval()
    H = W = 32
    image = smooth_images(gen, 1, H, W)[0]
    mask = ((torch.rand(H, W, generator=gen) > 0.5).long() * 255)
    with torch.no_grad():
        S = torch.softmax(seg(image.unsqueeze(0))["out"], dim=1)
    rec = {"image": image.numpy(), "mask": mask.numpy(), "S": S.numpy()}
    for steps, lam, thr, lr in ((1, 0.1, 0.5, 1e-2), (10, 0.1, 0.3, 1e-4), (20, 0.1, 0.5, 1e-2)):
        out = R["refine_pseudo_mask"](seg, image, mask, lambda_boundary=lam, threshold=thr, lr=lr, num_steps=steps)
        rec[f"refined_s{steps}"] = out.numpy()
    np.savez_compressed(os.path.join(OUT, "refine.npz"), **rec)

    # ---- IoU / accuracy (ExtraUtilities.py:4-21) ----
    pm = (torch.rand(50, 60, generator=gen) > 0.4).long()
    tm = (torch.rand(50, 60, generator=gen) > 0.5).long()
    iou, acc = R["compute_iou_and_acc"](pm, tm)
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), pred=pm.numpy(), true=tm.numpy(), iou=np.float64(iou), acc=np.float64(acc))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

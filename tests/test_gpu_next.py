"""GPU: the steps either side of the hot path: batched on-device refinement, label assembly without
the PNG round trip, fused IoU / accuracy counters."""
import os

import numpy as np
import pytest
import torch

from oracle import wsdl_oracle as O
from oracle.make_golden import FixedLogitsNet

pytestmark = pytest.mark.gpu


def test_refine_batched_matches_per_image_loop(golden_dir):
    """refine_pseudo_masks_batched == the reference loop (refine_pseudo_mask drop-in) image by image, and == the
    fixture the real reference produced."""
    from weaklysuperviseddl_b200.AlternatingDirectionCutLoss import refine_pseudo_mask, refine_pseudo_masks_batched

    d = np.load(os.path.join(golden_dir, "refine.npz"))
    torch.manual_seed(11)
    seg = FixedLogitsNet().eval().cuda()
    image, mask = torch.from_numpy(d["image"]), torch.from_numpy(d["mask"])
    gen = torch.Generator().manual_seed(5)
    # three different problems: the fixture image, a flipped copy, a noisy copy with a different initial mask
    images = torch.stack([image, image.flip(-1), (image + 0.05 * torch.rand(image.shape, generator=gen)).clamp(0, 1)])
    masks = torch.stack([mask, mask.flip(-1), torch.roll(mask, 3, 0)])
    for steps, lam, thr, lr in ((1, 0.1, 0.5, 1e-2), (10, 0.1, 0.3, 1e-4), (20, 0.1, 0.5, 1e-2)):
        out, X, Xf = refine_pseudo_masks_batched(seg, images.cuda(), masks.cuda(), lambda_boundary=lam, threshold=thr, lr=lr,
                                                 num_steps=steps, return_state=True)
        assert out.shape == (3,) + tuple(mask.shape)
        assert (out[0].cpu().numpy() != d[f"refined_s{steps}"]).mean() <= 0.002  # the real reference's result
        for b in range(3):
            one = refine_pseudo_mask(seg, images[b], masks[b].cuda(), lambda_boundary=lam, threshold=thr, lr=lr,
                                     num_steps=steps)
            assert (out[b] != one).float().mean().item() <= 0.002, (steps, b)
        # against the fp32 oracle port of the whole loop started from the same S, on the soft state (the fp64-arbitrated
        # bound lives in test_gpu_configs.py::test_refine_soft_state_vs_fp64)
        with torch.no_grad():
            S = torch.softmax(seg(images.cuda())["out"], dim=1).cpu()
        _, Xo, Xfo = O.refine_from_probs(S[:1], images[0], masks[0], lam, thr, lr, steps, return_state=True)
        assert (Xf[0].cpu() - Xfo[0]).abs().max().item() <= 1e-5


def test_labels_from_masks_equals_png_round_trip():
    from weaklysuperviseddl_b200 import functional as WF

    gen = torch.Generator().manual_seed(3)
    for (B, H, W, S) in ((2, 224, 224, 256), (1, 512, 512, 256), (3, 100, 37, 256), (1, 256, 256, 256), (2, 7, 300, 64)):
        m = (torch.rand(B, H, W, generator=gen) > 0.6).to(torch.uint8)
        lab = WF.labels_from_masks(m.cuda(), S)
        assert lab.dtype == torch.int64 and lab.shape == (B, S, S)
        for b in range(B):
            ref = O.png_mask_to_labels(O.mask_to_png_array(m[b].numpy()), S)
            assert np.array_equal(lab[b].cpu().numpy(), ref), (B, H, W, S, b)
    one = WF.labels_from_masks(m[0].cuda(), 64)
    assert one.shape == (64, 64) and torch.equal(one, lab[0])


@pytest.mark.parametrize("dtype", [torch.uint8, torch.int32, torch.int64, torch.float32])
def test_iou_acc_counts(dtype):
    from weaklysuperviseddl_b200 import functional as WF
    from weaklysuperviseddl_b200.ExtraUtilities import compute_iou_and_acc, compute_iou_and_acc_batched

    gen = torch.Generator().manual_seed(4)
    pred = (torch.rand(5, 97, 131, generator=gen) > 0.5).to(dtype)
    true = (torch.rand(5, 97, 131, generator=gen) > 0.4).to(dtype)
    true[4] = 0  # empty ground truth and
    pred[4] = 0  # empty prediction: iou = 0 / 1e-8 = 0, acc = 1
    c = WF.iou_acc_counts(pred.cuda(), true.cuda()).cpu()
    for b in range(5):
        pf, tf = pred[b] > 0, true[b] > 0
        assert c[b].tolist() == [int((pf & tf).sum()), int((pf | tf).sum()), int((pred[b] == true[b]).sum())]
        iou, acc = compute_iou_and_acc(pred[b].cuda(), true[b].cuda())
        ro, ra = O.iou_and_acc(pred[b], true[b])
        assert iou == ro and acc == ra
    bi, ba = compute_iou_and_acc_batched(pred.cuda(), true.cuda())
    assert abs(bi[0].item() - O.iou_and_acc(pred[0], true[0])[0]) < 1e-12 and ba[4].item() == 1.0 and bi[4].item() == 0.0

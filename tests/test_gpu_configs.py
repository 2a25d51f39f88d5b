"""GPU parity at the BASELINE.json configurations that round 1 left untested: config-5 geometry (4 backbone stages at
1024x1024), keep_largest at 1024x1024, the full 32x2x224x224 batch of configs[1] through the fused cut + boundary
launch, refinement against an fp64 run, and the two boundary callables that only had signature / shape checks
(evaluate_layercam_on_test_set, generate_bg_cam)."""
import os

import numpy as np
import pytest
import torch

from helpers import assert_cam_close, assert_masks_match, smooth_images, synth_act_grad
from oracle import wsdl_oracle as O
from oracle.make_golden import FixedLogitsNet, TinyCAMNet

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def WF():
    from weaklysuperviseddl_b200 import functional

    return functional


# ------------------------------------------------------------------------------------------------ config 5
def test_layercam_config5_geometry(WF):
    """BASELINE config 5: layers 1-4 of a ResNet-50 at 1024x1024 input (256x256^2, 512x128^2, 1024x64^2, 2048x64^2
    hooks; the layer-1 low-resolution map alone is 256 KB, larger than a CTA's shared memory), B = 2, against the
    fp64 oracle; masks bit-exact outside the 1e-6 band, band pixels counted by the kernel == counted here."""
    gen = torch.Generator(device="cuda").manual_seed(55)
    layers = [(256, 256, 256), (512, 128, 128), (1024, 64, 64), (2048, 64, 64)]
    B, S = 2, 1024
    acts = [torch.randn(B, C, h, w, device="cuda", generator=gen) for (C, h, w) in layers]  # negatives kept
    grads = [torch.randn(B, C, h, w, device="cuda", generator=gen) * 1e-3 for (C, h, w) in layers]
    cam, mask, near = WF.layercam_fused(acts, grads, (S, S), thresh=0.3)
    ref64 = O.layercam_from_hooks([a.cpu() for a in acts], [g.cpu() for g in grads], (S, S), dtype=torch.float64)
    assert_cam_close(cam, ref64, "config-5 geometry vs fp64 oracle")
    inside = assert_masks_match(mask, ref64, 0.3)
    assert int(near.item()) == int(((cam - 0.3).abs() < 1e-6).sum().item())
    assert abs(int(near.item()) - inside) <= 8 + inside // 4  # fp32 vs fp64 view of the same 1e-6 band
    # the variant arithmetic (per-layer alpha between two normalisations) at the same geometry, one image
    cam_v, _, _ = WF.layercam_fused([a[:1] for a in acts], [g[:1] for g in grads], (S, S), alpha=2.0, alpha_mode=1)
    ref_v = O.layercam_from_hooks([a[:1].cpu() for a in acts], [g[:1].cpu() for g in grads], (S, S), alpha=2.0,
                                  alpha_mode=1, dtype=torch.float64)
    assert_cam_close(cam_v, ref_v, "config-5 geometry, variant")


def test_near_count_exact_at_224(WF):
    """at out_w = 224 the second warp of a row block has lanes past the right edge; the band counter must
    still be exact (it used to shuffle with exited lanes)."""
    gen = torch.Generator().manual_seed(8)
    a, g = synth_act_grad(gen, 3, 64, 14, 14)
    for band in (1e-6, 1e-3, 5e-2):
        cam, mask, near = WF.layercam_fused([a.cuda()], [g.cuda()], (224, 224), thresh=0.3, near_band=band)
        assert int(near.item()) == int(((cam - 0.3).abs() < band).sum().item()), band
    cam, mask, near = WF.layercam_fused([a.cuda()], [g.cuda()], (50, 37), thresh=0.3, near_band=5e-2)  # ragged width
    assert int(near.item()) == int(((cam - 0.3).abs() < 5e-2).sum().item())


def test_keep_largest_1024(WF):
    rng = np.random.default_rng(21)
    blobs = torch.nn.functional.avg_pool2d(torch.from_numpy(rng.random((1, 1, 1024 + 24, 1024 + 24))).float(), 25, 1)[0, 0]
    cases = [(blobs.numpy() > np.quantile(blobs.numpy(), q)).astype(np.uint8) for q in (0.45, 0.7)]
    cases.append((rng.random((1024, 1024)) < 0.58).astype(np.uint8))  # near the 8-connectivity percolation threshold
    m = torch.from_numpy(np.stack(cases)).cuda()
    out, area = WF.keep_largest(m, return_area=True)
    for b, c in enumerate(cases):
        ref = O.keep_largest(c)
        assert np.array_equal(out[b].cpu().numpy(), ref), b
        assert int(area[b].item()) == int(ref.sum())


# ------------------------------------------------------------------------------------------------ configs[1]
def test_dual_full_bench_batch_vs_closed_form(WF):
    """The whole 32x2x224x224 batch of BASELINE configs[1] through ONE fused cut + boundary launch, every image
    against the fp64 closed form of the reference's loss and autograd gradient (oracle.pairwise_closed_form, itself
    proven equal to autograd through the op-for-op restatement in tests/test_oracle.py)."""
    B, H, W = 32, 224, 224
    gen = torch.Generator().manual_seed(1)
    logits = torch.randn(B, 2, H, W, generator=gen)
    img = smooth_images(gen, B, H, W)
    go_b = torch.rand(B, generator=gen) + 0.5
    lc, lb, g = WF.pairwise_dual_loss_and_grad(logits.cuda(), img.cuda(), 0.05, 0.1, 5.0, 5,
                                               grad_out_cut=None, grad_out_bnd=go_b.cuda())
    lc, lb, g = lc.cpu().double(), lb.cpu().double(), g.cpu().double()
    p = torch.softmax(logits.double(), 1)
    cut_sum, worst = 0.0, 0.0
    for b in range(B):
        l_c, g_c = O.pairwise_closed_form(logits[b].numpy(), img[b].numpy(), 0.05, None, 5, True, True)
        l_b, g_p = O.pairwise_closed_form(p[b].numpy(), img[b].numpy(), 0.1, 5.0, 5, False, False)
        g_p = torch.from_numpy(g_p)
        g_b = p[b] * (g_p - (p[b] * g_p).sum(0, keepdim=True))  # through softmax(logits)
        ref = torch.from_numpy(g_c) / B + go_b[b].double() * g_b  # cut: mean over the batch (kappa has 1/B)
        cut_sum += l_c
        assert abs(lb[b].item() - l_b) <= 1e-5 * abs(l_b), (b, lb[b].item(), l_b)
        scale = ref.abs().max().item()
        err = (g[b] - ref).abs().max().item()
        worst = max(worst, err / scale)
        assert err <= 1e-5 * scale, (b, err, scale)
    assert abs(lc.item() - cut_sum / B) <= 1e-5 * abs(cut_sum / B)
    print(f"full batch: worst gradient error {worst:.2e} of each image's gradient scale")


# ------------------------------------------------------------------------------------------------ refinement
@pytest.mark.parametrize("steps,lam,thr,lr", [(1, 0.1, 0.5, 1e-2), (10, 0.1, 0.3, 1e-4), (20, 0.1, 0.5, 1e-2)])
def test_refine_soft_state_vs_fp64(golden_dir, steps, lam, thr, lr):
    """refine_pseudo_masks_batched against an fp64 run of the reference loop (oracle.refine_from_probs, the port of
    AlternatingDirectionCutLoss.py:709-767) started from the SAME network output S.  fp32 Adam allows this much: the
    reference's own fp32 arithmetic drifts from the fp64 run by 5e-8 / 3e-7 / 1.2e-6 in X after 1 / 10 / 20 steps
    (measured on the fixture), so X and softmax(X) are held to 1e-5 absolute (north_star's 1e-5 on values of order 1)
    and the thresholded mask may differ only where |softmax(X)[1] - threshold| is inside that bound."""
    from weaklysuperviseddl_b200.AlternatingDirectionCutLoss import refine_pseudo_masks_batched

    d = np.load(os.path.join(golden_dir, "refine.npz"))
    torch.manual_seed(11)
    seg = FixedLogitsNet().eval().cuda()
    image, mask = torch.from_numpy(d["image"]), torch.from_numpy(d["mask"])
    gen = torch.Generator().manual_seed(5)
    images = torch.stack([image, image.flip(-1), (image + 0.05 * torch.rand(image.shape, generator=gen)).clamp(0, 1)])
    masks = torch.stack([mask, mask.flip(-1), torch.roll(mask, 3, 0)])
    out, X, Xf = refine_pseudo_masks_batched(seg, images.cuda(), masks.cuda(), lambda_boundary=lam, threshold=thr,
                                             lr=lr, num_steps=steps, return_state=True)
    with torch.no_grad():
        S = torch.softmax(seg(images.cuda())["out"], dim=1).cpu()
    tol = 1e-5
    for b in range(3):
        ref, X64, Xf64 = O.refine_from_probs(S[b:b + 1].double(), images[b].double(), masks[b], lam, thr, lr, steps,
                                             return_state=True)
        ex = (X[b].cpu().double() - X64[0]).abs().max().item()
        ef = (Xf[b].cpu().double() - Xf64[0]).abs().max().item()
        assert ex <= tol and ef <= tol, (steps, b, ex, ef)
        flips = out[b].cpu() != ref
        assert not (flips & ((Xf64[0, 1] - thr).abs() > tol)).any(), (steps, b)
    # and the fixture the REAL reference produced for image 0 (its own fp32 run, S from a CPU convolution)
    assert (out[0].cpu().numpy() != d[f"refined_s{steps}"]).mean() <= 0.002


def test_refine_graph_replay_equals_direct_launches(golden_dir):
    """The captured step loop (one CUDA graph per shape / hyper-parameter set, replayed on later calls) gives the same
    bits as launching the 2 * steps + 1 kernels directly, call after call, and B = 1 equals the batched run's image."""
    from weaklysuperviseddl_b200.AlternatingDirectionCutLoss import refine_pseudo_mask, refine_pseudo_masks_batched

    torch.manual_seed(11)
    seg = FixedLogitsNet().eval().cuda()
    gen = torch.Generator().manual_seed(6)
    images = smooth_images(gen, 4, 64, 80).cuda()
    masks = ((torch.rand(4, 64, 80, generator=gen) > 0.5).long() * 255).cuda()
    a = refine_pseudo_masks_batched(seg, images, masks, num_steps=7, return_state=True, use_graph=True)
    b = refine_pseudo_masks_batched(seg, images, masks, num_steps=7, return_state=True, use_graph=False)
    c = refine_pseudo_masks_batched(seg, images, masks, num_steps=7, return_state=True, use_graph=True)  # cached graph
    for x, y, z in zip(a, b, c):
        assert torch.equal(x, y) and torch.equal(x, z)
    assert a[0].shape == (4, 64, 80) and set(a[0].unique().tolist()) <= {0.0, 1.0}
    one = refine_pseudo_mask(seg, images[2], masks[2], num_steps=7)
    assert one.shape == (64, 80) and (one != a[0][2]).float().mean().item() <= 1e-3  # S of a batch of one: cuDNN may differ in the last bit


# ------------------------------------------------------------------------------------------------ boundary callables
class _Loader:
    """Batch-size-1 loader in the reference's format: (img (1,3,H,W), (label (1,), true_mask (1,1,h,w)))
    (LayerCAM.py:96-99; the Oxford-IIIT-Pet wrapper with PILToTensor masks)."""

    def __init__(self, items):
        self.items = items

    def __iter__(self):
        return iter(self.items)


def test_evaluate_layercam_on_test_set_functional(golden_dir, capsys):
    """LayerCAM.py:84-130 end to end on a 12-item loader: only the first 11 items count (:119-120), the truth is
    (1,h,w) so the prediction always goes through the nearest resize (:109-112) and the IoU through broadcasting
    (ExtraUtilities.py:11-15).  Expected values: the oracle path on the hooks the generator saw."""
    from weaklysuperviseddl_b200.LayerCAM import LayerCAMGenerator, evaluate_layercam_on_test_set

    lc = np.load(os.path.join(golden_dir, "layercam_tiny.npz"))
    net = TinyCAMNet().eval()
    net.load_state_dict({k[3:]: torch.from_numpy(lc[k]) for k in lc.files if k.startswith("w::")})
    net = net.cuda()
    gen_cam = LayerCAMGenerator(net, ["layer3", "layer4"])
    g = torch.Generator().manual_seed(31)
    items = []
    for i in range(12):
        img = torch.rand(1, 3, 96, 80, generator=g)
        label = torch.randint(0, 5, (1,), generator=g)
        hw = (224, 224) if i % 3 else (200, 180)  # some truths need a real resize
        yy, xx = torch.meshgrid(torch.arange(hw[0]), torch.arange(hw[1]), indexing="ij")
        cy, cx, r = hw[0] * (0.3 + 0.4 * torch.rand(1, generator=g)), hw[1] * (0.3 + 0.4 * torch.rand(1, generator=g)), 60
        truth = (((yy - cy) ** 2 + (xx - cx) ** 2) < r * r).long()
        truth[:4] = 2  # trimap-style other value: only == 1 is foreground (LayerCAM.py:99)
        items.append((img, (label, truth[None, None])))
    res = evaluate_layercam_on_test_set(gen_cam, _Loader(items), alpha=1.0, cam_thresh=0.3)
    assert set(res) == {"layercam_fg_iou", "layercam_fg_acc"}
    assert "Evaluation of CAMs on test set" in capsys.readouterr().out
    ious, accs, band = [], [], 0
    for img, (label, truth) in items[:11]:
        acts, grads = gen_cam._forward_backward(img[0].cuda(), label.cuda())
        cam = O.layercam_from_hooks([a.cpu() for a in acts], [x.cpu() for x in grads], (224, 224), dtype=torch.float64)[0]
        band += int(((cam - 0.3).abs() < 1e-6).sum())
        pred = torch.from_numpy(O.threshold_mask(cam, 0.3)).long()
        t = (truth[0] == 1).long()
        if pred.shape != t.shape:
            pred = torch.nn.functional.interpolate(pred[None, None].float(), size=t.shape[-2:], mode="nearest").squeeze().long()
        iou, acc = O.iou_and_acc(pred, t)
        ious.append(iou)
        accs.append(acc)
    slack = 0.0 if band == 0 else 1e-4
    assert abs(res["layercam_fg_iou"] - sum(ious) / 11) <= slack
    assert abs(res["layercam_fg_acc"] - sum(accs) / 11) <= slack
    # a loader shorter than 11 items is averaged over what it has
    res3 = evaluate_layercam_on_test_set(gen_cam, _Loader(items[:3]))
    assert abs(res3["layercam_fg_iou"] - sum(ious[:3]) / 3) <= slack


def test_compute_iou_and_acc_broadcasts_like_the_reference():
    """(H,W) prediction against a (1,H,W) truth, as the reference's loader produces."""
    from weaklysuperviseddl_b200.ExtraUtilities import compute_iou_and_acc

    g = torch.Generator().manual_seed(2)
    pred = (torch.rand(97, 131, generator=g) > 0.5).long()
    true = (torch.rand(1, 97, 131, generator=g) > 0.4).long()
    assert compute_iou_and_acc(pred.cuda(), true.cuda()) == O.iou_and_acc(pred, true)
    assert compute_iou_and_acc(pred.cuda()[None], true.cuda()[0]) == O.iou_and_acc(pred[None], true[0])
    with pytest.raises(RuntimeError):
        compute_iou_and_acc(pred.cuda(), true.cuda()[:, :50])


def test_generate_bg_cam_vs_oracle(golden_dir):
    """AlternatingDirectionCutLoss.py:296-318 against oracle.bg_cam_from_fg on the CAMs the generator returns, for the
    reference's default exponent and another one, with several candidate classes (max over dim 0)."""
    from weaklysuperviseddl_b200.AlternatingDirectionCutLoss import LayerCAMGenerator as Variant

    lc = np.load(os.path.join(golden_dir, "layercam_tiny.npz"))
    net = TinyCAMNet().eval()
    net.load_state_dict({k[3:]: torch.from_numpy(lc[k]) for k in lc.files if k.startswith("w::")})
    gen_cam = Variant(net.cuda(), ["layer2", "layer3", "layer4"])
    images = torch.from_numpy(lc["images"]).cuda()
    labels = torch.from_numpy(lc["labels"]).cuda()
    for alpha in (2.0, 0.5):
        m_bg, m_fg = gen_cam.generate_bg_cam(images[0].clone(), labels[0:1], alpha=alpha)
        cams = gen_cam.generate(images[0].clone(), labels[0:1])
        r_bg, r_fg = O.bg_cam_from_fg(cams.cpu().double(), alpha=alpha)
        assert m_bg.shape == (224, 224) and m_fg.shape == (224, 224)
        assert (m_bg.cpu().double() - r_bg).abs().max().item() <= 1e-6
        assert (m_fg.cpu().double() - r_fg).abs().max().item() <= 1e-6
        assert m_bg.min().item() >= 0.0 and m_bg.max().item() <= 1.0


# ------------------------------------------------------------------------------------------------ config 3
def test_sharded_pseudo_masks_on_device(WF):
    """generate_pseudo_masks_sharded with the fused launch: four simulated ranks (explicit rank / world, no process
    group) partition 21 images; their union equals the single-rank run bit for bit and the masks equal the oracle's
    outside the threshold band."""
    from weaklysuperviseddl_b200.PsuedoMasks import generate_pseudo_masks_sharded

    layers = ((64, 16, 16), (128, 8, 8))

    def hooks(indices):
        acts, grads = [], []
        for (C, h, w) in layers:
            a, g = [], []
            for i in indices:
                gen = torch.Generator().manual_seed(i)  # seed = image index
                a.append(torch.randn(C, h, w, generator=gen).relu())
                g.append(torch.randn(C, h, w, generator=gen) * 1e-3)
            acts.append(torch.stack(a).cuda())
            grads.append(torch.stack(g).cuda())
        return acts, grads

    single = generate_pseudo_masks_sharded(hooks, 21, (96, 96), chunk=8, rank=0, world=1, keep_largest_masks=True, count_foreground=True)
    assert single["masks"].shape == (21, 96, 96) and single["counters"]["masks"] == 21
    fg = 0
    for r in range(4):
        part = generate_pseudo_masks_sharded(hooks, 21, (96, 96), chunk=3, rank=r, world=4, keep_largest_masks=True, count_foreground=True, streams=2)
        assert part["indices"].tolist() == list(range(r, 21, 4))
        for j, i in enumerate(part["indices"].tolist()):
            assert torch.equal(part["masks"][j], single["masks"][i]), (r, i)
        fg += part["counters"]["foreground_pixels"]
    assert fg == single["counters"]["foreground_pixels"] == int(single["masks"].sum().item())
    acts, grads = hooks([5])
    ref = O.layercam_from_hooks([a.cpu() for a in acts], [g.cpu() for g in grads], (96, 96), dtype=torch.float64)[0]
    plain = generate_pseudo_masks_sharded(hooks, 6, (96, 96), chunk=4, rank=0, world=1)["masks"][5]
    assert_masks_match(plain, ref, 0.3)

"""Pins oracle/wsdl_oracle.py (the CPU restatement) against (a) the committed golden vectors that
oracle/make_golden.py produced by running the REAL reference functions, and (b) when /root/reference is
mounted (authoring container only), the reference functions themselves on fresh seeded inputs.
CPU only -- part of `-m "not gpu"`."""
import os

import numpy as np
import pytest
import torch

from helpers import smooth_images
from oracle import ref_loader
from oracle import wsdl_oracle as O

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted here")


@pytest.fixture(scope="module")
def lc(golden_dir):
    return np.load(os.path.join(golden_dir, "layercam_tiny.npz"))


@pytest.fixture(scope="module")
def pw(golden_dir):
    return np.load(os.path.join(golden_dir, "pairwise.npz"))


# ------------------------------------------------------------ golden vectors: LayerCAM
@pytest.mark.parametrize("alpha", [1.0, 0.5])
def test_layercam_main_golden(lc, alpha):
    for i in range(2):
        acts = [torch.from_numpy(lc[f"main_act_{n}_{i}"]) for n in ("layer3", "layer4")]
        grads = [torch.from_numpy(lc[f"main_grad_{n}_{i}"]) for n in ("layer3", "layer4")]
        cam = O.layercam_from_hooks(acts, grads, (224, 224), alpha=alpha, alpha_mode=0)
        ref = torch.from_numpy(lc[f"main_cam_alpha{alpha}"][i])
        assert torch.equal(cam[0], ref), "fp32 oracle must reproduce the reference bit for bit on CPU"
        cam2 = O.layercam_from_hooks(acts, grads, (224, 224), alpha=alpha, alpha_mode=0, use_torch_interpolate=False)
        assert (cam2[0] - ref).abs().max() <= 2.4e-7  # own bilinear restatement: within 2 ulp at 1.0


@pytest.mark.parametrize("alpha", [1.0, 2.0])
def test_layercam_variant_golden(lc, alpha):
    names = ("layer2", "layer3", "layer4")
    for i in range(2):
        acts = [torch.from_numpy(lc[f"variant_act_{n}_{i}"]) for n in names]
        grads = [torch.from_numpy(lc[f"variant_grad_{n}_{i}"]) for n in names]
        cam = O.layercam_from_hooks(acts, grads, (224, 224), alpha=alpha, alpha_mode=1)
        assert torch.equal(cam[0], torch.from_numpy(lc[f"variant_cam_alpha{alpha}"][i]))


def test_layercam_fp64_close_to_fp32(lc):
    acts = [torch.from_numpy(lc[f"main_act_{n}_0"]) for n in ("layer3", "layer4")]
    grads = [torch.from_numpy(lc[f"main_grad_{n}_0"]) for n in ("layer3", "layer4")]
    c32 = O.layercam_from_hooks(acts, grads, (224, 224))
    c64 = O.layercam_from_hooks(acts, grads, (224, 224), dtype=torch.float64)
    assert (c32.double() - c64).abs().max() < 1e-6


def test_threshold_idiom():
    cam = torch.tensor([[0.0, 0.29999, 0.3, 0.31, float("nan"), 1.0, -0.0]])
    m = O.threshold_mask(cam, 0.3)
    assert m.tolist() == [[0, 0, 1, 1, 0, 1, 0]]
    m0 = O.threshold_mask(cam, 0.0)  # thr=0: cam > 0 still drops exact zeros
    assert m0.tolist() == [[0, 1, 1, 1, 0, 1, 0]]


# ------------------------------------------------------------ golden vectors: pairwise
def test_cut_loss_golden(pw):
    logits, img = torch.from_numpy(pw["logits"]), torch.from_numpy(pw["images"])
    val, g = O.loss_and_grad(O.cut_loss, logits, img)
    assert abs(val.item() - pw["cut_loss"].item()) <= 1e-6 * abs(pw["cut_loss"].item())
    assert np.abs(g.numpy() - pw["cut_grad"]).max() <= 1e-5 * np.abs(pw["cut_grad"]).max()
    val, g = O.loss_and_grad(O.cut_loss, logits[0], img[0])
    assert abs(val.item() - pw["cut3d_loss"].item()) <= 1e-6 * abs(pw["cut3d_loss"].item())
    assert np.abs(g.numpy() - pw["cut3d_grad"]).max() <= 1e-5 * np.abs(pw["cut3d_grad"]).max()
    val, g = O.loss_and_grad(O.cut_loss, logits, img, sigma_color=0.2, window_size=3)
    assert abs(val.item() - pw["cut_w3_loss"].item()) <= 1e-6 * abs(pw["cut_w3_loss"].item())
    assert np.abs(g.numpy() - pw["cut_w3_grad"]).max() <= 1e-5 * np.abs(pw["cut_w3_grad"]).max()
    val, g = O.loss_and_grad(O.cut_loss, torch.from_numpy(pw["logits3"]), torch.from_numpy(pw["images3"]))
    assert abs(val.item() - pw["cut_c3_loss"].item()) <= 1e-6 * abs(pw["cut_c3_loss"].item())
    assert np.abs(g.numpy() - pw["cut_c3_grad"]).max() <= 1e-5 * np.abs(pw["cut_c3_grad"]).max()


def test_boundary_loss_golden(pw):
    logits, img = torch.from_numpy(pw["logits"]), torch.from_numpy(pw["images"])
    probs = torch.softmax(logits, dim=1)
    for b in range(2):
        val, g = O.loss_and_grad(O.boundary_loss, probs[b], img[b])
        assert abs(val.item() - pw["boundary_loss"][b]) <= 1e-6 * abs(pw["boundary_loss"][b])
        assert np.abs(g.numpy() - pw["boundary_grad"][b]).max() <= 1e-5 * np.abs(pw["boundary_grad"][b]).max()


def test_affinities_golden(pw):
    img = torch.from_numpy(pw["images"])
    a = torch.stack(O.affinities(img, 0.1, 5, 5), dim=1).squeeze(2)
    assert torch.equal(a, torch.from_numpy(pw["affinities_batched"]))
    s = torch.stack(O.affinities(img[1], 0.1, 5, 5), dim=0).squeeze(1)
    assert torch.equal(s, torch.from_numpy(pw["affinities_single"]))


@pytest.mark.parametrize("case", ["cut", "cut_w3", "cut_c3", "boundary"])
def test_closed_form_matches_reference_autograd(pw, case):
    """The gather-form gradient the CUDA kernel implements, incl. reflect borders."""
    if case == "boundary":
        probs = torch.softmax(torch.from_numpy(pw["logits"]), dim=1).numpy()
        for b in range(2):
            L, g = O.pairwise_closed_form(probs[b], pw["images"][b], 0.1, 5, 5, False, False)
            assert abs(L - pw["boundary_loss"][b]) <= 2e-6 * abs(pw["boundary_loss"][b])
            assert np.abs(g - pw["boundary_grad"][b]).max() <= 1e-5 * np.abs(pw["boundary_grad"][b]).max()
        return
    logits = pw["logits3"] if case == "cut_c3" else pw["logits"]
    imgs = pw["images3"] if case == "cut_c3" else pw["images"]
    sc, win = (0.2, 3) if case == "cut_w3" else (0.05, 5)
    B = logits.shape[0]
    tot, grads = 0.0, []
    for b in range(B):
        L, g = O.pairwise_closed_form(logits[b], imgs[b], sc, None, win, True, True)
        tot += L / B
        grads.append(g / B)
    assert abs(tot - pw[f"{case}_loss"].item()) <= 2e-6 * abs(pw[f"{case}_loss"].item())
    ref = pw[f"{case}_grad"]
    assert np.abs(np.stack(grads) - ref).max() <= 1e-5 * np.abs(ref).max()


def test_refine_golden(golden_dir):
    d = np.load(os.path.join(golden_dir, "refine.npz"))
    S, image, mask = torch.from_numpy(d["S"]), torch.from_numpy(d["image"]), torch.from_numpy(d["mask"])
    for steps, lam, thr, lr in ((1, 0.1, 0.5, 1e-2), (10, 0.1, 0.3, 1e-4), (20, 0.1, 0.5, 1e-2)):
        out = O.refine_from_probs(S, image, mask, lambda_boundary=lam, threshold=thr, lr=lr, num_steps=steps)
        ref = d[f"refined_s{steps}"]
        assert (out.numpy() != ref).sum() == 0


def test_metrics_golden(golden_dir):
    d = np.load(os.path.join(golden_dir, "metrics.npz"))
    iou, acc = O.iou_and_acc(torch.from_numpy(d["pred"]), torch.from_numpy(d["true"]))
    assert iou == float(d["iou"]) and acc == float(d["acc"])


# ------------------------------------------------------------ keep_largest (skimage absent: scipy vs cv2)
def test_keep_largest_scipy_vs_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for density in (0.3, 0.5, 0.62):
        m = (rng.random((61, 47)) < density).astype(np.uint8)
        ours = O.keep_largest(m)
        n, lab, stats, _ = cv2.connectedComponentsWithStats(m, connectivity=8)
        areas = stats[1:, cv2.CC_STAT_AREA]
        best = 1 + int(np.argmax(areas))
        assert np.array_equal(ours, (lab == best).astype(np.uint8))
    z = np.zeros((5, 5), np.uint8)
    assert O.keep_largest(z) is z  # empty -> input returned (PsuedoMasks.py:18-19)
    diag = np.eye(6, dtype=np.uint8)  # 8-connectivity joins a diagonal
    assert O.keep_largest(diag).sum() == 6


def test_keep_largest_tie_takes_first_in_raster_order():
    m = np.zeros((5, 9), np.uint8)
    m[3, 0:2] = 1   # area 2, first pixel at row 3
    m[1, 5:7] = 1   # area 2, first pixel at row 1 -> label 1 -> wins
    out = O.keep_largest(m)
    assert out[1, 5:7].sum() == 2 and out[3].sum() == 0


def test_keep_largest_two_independent_restatements_agree():
    """skimage cannot be had (absent here and from the wheelhouse), so keep_largest parity is UNPINNED against it;
    what we can pin: the scipy-based restatement and the from-scratch two-pass restatement of skimage's own
    union-find labelling (min raster index = root, labels in raster order of the first pixel) agree on labels
    AND on the kept component, ties included."""
    from scipy import ndimage

    rng = np.random.default_rng(17)
    cases = [(rng.random((41, 57)) < d).astype(np.uint8) for d in (0.2, 0.45, 0.55, 0.62, 0.9)]
    tie = np.zeros((9, 12), np.uint8)
    tie[6, 0:3] = 1   # three components of area 3; the one whose first pixel comes first in raster order wins
    tie[2, 8:11] = 1
    tie[4:7, 5] = 1
    cases.append(tie)
    snake = np.zeros((12, 12), np.uint8)  # U shape: the two arms get provisional labels that must be merged late
    snake[0:10, 1] = 1
    snake[0:10, 9] = 1
    snake[10, 1:10] = 1
    cases.append(snake)
    checker = (np.indices((10, 10)).sum(0) % 2).astype(np.uint8)  # 8-connectivity joins the whole checkerboard
    cases.append(checker)
    for m in cases:
        lab_a = O.label_two_pass(m)
        lab_b, _ = ndimage.label(m != 0, structure=np.ones((3, 3), dtype=np.int32))
        assert np.array_equal(lab_a, lab_b)  # same components AND same label numbering (raster order)
        assert np.array_equal(O.keep_largest_two_pass(m), O.keep_largest(m))
    assert O.keep_largest_two_pass(tie)[2, 8:11].sum() == 3
    z = np.zeros((4, 4), np.uint8)
    assert O.keep_largest_two_pass(z) is z
    assert O.label_two_pass(checker).max() == 1


def test_png_label_roundtrip():
    m = (np.random.default_rng(0).random((224, 224)) > 0.5).astype(np.uint8)
    png = O.mask_to_png_array(m)
    assert set(np.unique(png)) <= {0, 255} and png.shape == (224, 224, 3)
    lab = O.png_mask_to_labels(png, 256)
    assert lab.shape == (256, 256) and set(np.unique(lab)) <= {0, 1}
    src = np.floor((np.arange(256) + 0.5) * 224 / 256).astype(int)
    assert np.array_equal(lab, m[src][:, src].astype(np.int64))


# ------------------------------------------------------------ live reference (authoring container only)
@needs_ref
def test_oracle_vs_live_reference_pairwise():
    R = ref_loader.load()
    gen = torch.Generator().manual_seed(99)
    logits = torch.randn(2, 2, 21, 19, generator=gen)
    img = smooth_images(gen, 2, 21, 19)
    x = logits.clone().requires_grad_(True)
    v = R["LocalNormalizedCutLoss"](0.05, 5)(x, img)
    v.backward()
    ov, og = O.loss_and_grad(O.cut_loss, logits, img)
    assert abs(ov.item() - v.item()) <= 1e-6 * abs(v.item())
    assert (og - x.grad).abs().max() <= 1e-5 * x.grad.abs().max()
    p = torch.softmax(logits, 1)[0]
    x = p.clone().requires_grad_(True)
    v = R["ConstrainToBoundaryLossSingle"](0.1, 5, 5)(x, img[0])
    v.backward()
    ov, og = O.loss_and_grad(O.boundary_loss, p, img[0])
    assert abs(ov.item() - v.item()) <= 1e-6 * abs(v.item())
    assert (og - x.grad).abs().max() <= 1e-5 * x.grad.abs().max()


@needs_ref
def test_oracle_vs_live_reference_layercam():
    from oracle.make_golden import TinyCAMNet

    R = ref_loader.load()
    torch.manual_seed(3)
    net = TinyCAMNet().eval()
    g = R["LayerCAMGenerator"](net, ["layer2", "layer4"])
    img = torch.rand(3, 64, 72)
    cam = g.generate(img.clone(), 0.5, class_idx=torch.tensor([2]))
    acts = [g.activations[n].detach() for n in ("layer2", "layer4")]
    grads = [g.gradients[n].detach() for n in ("layer2", "layer4")]
    assert torch.equal(cam, O.layercam_from_hooks(acts, grads, (224, 224), alpha=0.5))

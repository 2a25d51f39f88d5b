"""GPU parity of the one-launch weak-supervision loss (wsdl_weak_loss_fwd_bwd, csrc/pairwise_stream.cu): cross-entropy
on the pseudo-labels + cut loss + per-image boundary loss, forward values and d total / d logits, against the fp64
oracle composition of the reference's own functions (oracle.cut_loss = AlternatingDirectionCutLoss.py:71-105,
oracle.boundary_loss = AlternatingDirectionBoundaryLoss.py:20-44, F.cross_entropy = SegmentationModel.py:96-113).
Tolerance: 1e-5 relative (north_star); bf16 / u8 inputs: on the values as given (the arithmetic stays fp32)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import assert_grad_close, assert_loss_close, smooth_images
from oracle import wsdl_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def WF():
    from weaklysuperviseddl_b200 import functional

    return functional


def oracle_total(logits, images, labels, lam_ce, go_cut, go_bnd, ignore_index=-100):
    """fp64 autograd through the op-for-op restatement: returns (total, ce, cut, bnd[B], d total / d logits)."""
    x = logits.double().clone().requires_grad_(True)
    img = images.double()
    B = x.shape[0]
    ce = F.cross_entropy(x, labels.long(), ignore_index=ignore_index) if labels is not None else x.new_zeros(())
    cut = O.cut_loss(x, img, sigma_color=0.05, window_size=5)
    p = torch.softmax(x, 1)
    bnd = torch.stack([O.boundary_loss(p[b], img[b], sigma_color=0.1, sigma_space=5, window_size=5) for b in range(B)])
    total = lam_ce * ce + go_cut * cut + (go_bnd.double() * bnd).sum()
    (g,) = torch.autograd.grad(total, x)
    return total.detach(), ce.detach(), cut.detach(), bnd.detach(), g


def closed_form_total(logits, images, labels, lam_ce, go_cut, go_bnd, ignore_index=-100):
    """The same numbers from the fp64 closed form (fast enough for full-size batches)."""
    B = logits.shape[0]
    x = logits.double()
    p = torch.softmax(x, 1)
    g = torch.zeros_like(x)
    cut_sum, bnd = 0.0, []
    for b in range(B):
        l_c, g_c = O.pairwise_closed_form(x[b].numpy(), images[b].double().numpy(), 0.05, None, 5, True, True)
        l_b, g_p = O.pairwise_closed_form(p[b].numpy(), images[b].double().numpy(), 0.1, 5.0, 5, False, False)
        g_p = torch.from_numpy(g_p)
        g[b] = go_cut * torch.from_numpy(g_c) / B + go_bnd[b].double() * p[b] * (g_p - (p[b] * g_p).sum(0, keepdim=True))
        cut_sum += l_c
        bnd.append(l_b)
    ce = x.new_zeros(())
    if labels is not None:
        xr = x.clone().requires_grad_(True)
        ce = F.cross_entropy(xr, labels.long(), ignore_index=ignore_index)
        g = g + lam_ce * torch.autograd.grad(ce, xr)[0]
        ce = ce.detach()
    bnd = torch.tensor(bnd, dtype=torch.float64)
    total = lam_ce * ce + go_cut * cut_sum / B + (go_bnd.double() * bnd).sum()
    return total, ce, torch.tensor(cut_sum / B), bnd, g


def _labels(gen, B, H, W, dtype=torch.int64, p_ignore=0.0, ignore_index=-100):
    lab = (torch.rand(B, H, W, generator=gen) > 0.45).to(torch.int64)
    if p_ignore:
        lab[torch.rand(B, H, W, generator=gen) < p_ignore] = ignore_index
    return lab.to(dtype)


SHAPES = [(2, 64, 48), (1, 224, 224), (3, 40, 124), (2, 6, 8), (1, 7, 228), (5, 33, 60), (1, 100, 64), (2, 256, 256),
          (1, 38, 12), (4, 77, 68)]


@pytest.mark.parametrize("case", range(len(SHAPES)))
def test_weak_loss_vs_oracle(WF, case):
    B, H, W = SHAPES[case]
    gen = torch.Generator().manual_seed(300 + case)
    logits = torch.randn(B, 2, H, W, generator=gen) * (1.0 + case % 3)
    img = smooth_images(gen, B, H, W)
    labels = _labels(gen, B, H, W, p_ignore=0.1 if case % 2 else 0.0)
    go_c = torch.tensor([0.1 + 0.05 * case])
    go_b = torch.rand(B, generator=gen) + 0.25
    lam_ce = 1.0 if case % 3 else 0.7
    total, ce, cut, bnd, g = WF.weak_loss_and_grad(logits.cuda(), img.cuda(), labels.cuda(), lam_ce, go_c.cuda(), go_b.cuda())
    ref = (oracle_total if H * W * B <= 20000 else closed_form_total)(logits, img, labels, lam_ce, go_c.item(), go_b)
    assert_loss_close(total, ref[0], "total")
    assert_loss_close(ce, ref[1], "cross-entropy")
    assert_loss_close(cut, ref[2], "cut")
    assert_loss_close(bnd, ref[3], "boundary")
    assert_grad_close(g, ref[4], f"d total / d logits {SHAPES[case]}")
    # forward only gives the same values; without labels the CE term is absent
    t2, ce2, cut2, bnd2, none = WF.weak_loss_and_grad(logits.cuda(), img.cuda(), labels.cuda(), lam_ce, go_c.cuda(), go_b.cuda(),
                                                      want_grad=False)
    assert none is None and torch.equal(t2, total) and torch.equal(bnd2, bnd)
    t3, ce3, cut3, bnd3, g3 = WF.weak_loss_and_grad(logits.cuda(), img.cuda(), None, 0.0, go_c.cuda(), go_b.cuda())
    assert ce3 is None and torch.equal(cut3, cut) and torch.equal(bnd3, bnd)
    # the two-regulariser entry (one tile per CTA) runs the same march: equal to rounding
    lc, lb, gd = WF.pairwise_dual_loss_and_grad(logits.cuda(), img.cuda(), grad_out_cut=go_c.cuda(), grad_out_bnd=go_b.cuda())
    assert abs(lc.item() - cut.item()) <= 2e-6 * abs(cut.item()) and (lb - bnd).abs().max() <= 2e-6 * bnd.abs().max()
    assert (gd - g3).abs().max().item() <= 2e-6 * g3.abs().max().item()


def test_element_wise_gradient_error_is_reported(WF, capsys):
    """The gradient is compared relative to max|grad| (terms cancel, element-wise relative error is meaningless where the
    gradient crosses zero); the element-wise worst case over entries above 1 % of the scale is printed and bounded too."""
    gen = torch.Generator().manual_seed(5)
    logits = torch.randn(2, 2, 96, 96, generator=gen)
    img = smooth_images(gen, 2, 96, 96)
    labels = _labels(gen, 2, 96, 96)
    go_c, go_b = torch.tensor([0.1]), torch.full((2,), 0.25)
    _, _, _, _, g = WF.weak_loss_and_grad(logits.cuda(), img.cuda(), labels.cuda(), 1.0, go_c.cuda(), go_b.cuda())
    ref = closed_form_total(logits, img, labels, 1.0, 0.1, go_b)[4]
    big = ref.abs() > 0.01 * ref.abs().max()
    rel = ((g.cpu().double() - ref).abs() / ref.abs())[big].max().item()
    print(f"element-wise worst relative gradient error over entries > 1 % of the scale: {rel:.2e}")
    assert rel <= 1e-4


@pytest.mark.parametrize("lab_dtype", [torch.uint8, torch.int64])
def test_bf16_logits_u8_images_u8_labels(WF, lab_dtype):
    """Config 4's real dtypes: the network's bf16 output, the dataset's 8-bit images, byte labels.  Arithmetic is fp32
    on the values as given: the result must equal the f32 path on the up-converted inputs bit for bit (losses) and
    to bf16 rounding (gradient)."""
    gen = torch.Generator().manual_seed(9)
    B, H, W = 3, 96, 128
    logits = (torch.randn(B, 2, H, W, generator=gen) * 2).to(torch.bfloat16)
    img8 = (smooth_images(gen, B, H, W) * 255).round().to(torch.uint8)
    labels = _labels(gen, B, H, W, dtype=lab_dtype)
    go_c, go_b = torch.tensor([0.1]).cuda(), torch.full((B,), 0.5 / B).cuda()
    out8 = WF.weak_loss_and_grad(logits.cuda(), img8.cuda(), labels.cuda(), 1.0, go_c, go_b)
    out32 = WF.weak_loss_and_grad(logits.float().cuda(), (img8.float() / 255).cuda(), labels.long().cuda(), 1.0, go_c, go_b)
    for a, b in zip(out8[:4], out32[:4]):
        assert torch.equal(a, b)
    assert out8[4].dtype == torch.bfloat16
    assert torch.equal(out8[4], out32[4].to(torch.bfloat16))  # same fp32 gradient, rounded once
    # bf16 logits + f32 images + labels: the training step's own types have a compiled variant of the kernel
    outb = WF.weak_loss_and_grad(logits.cuda(), (img8.float() / 255).cuda(), labels.cuda(), 1.0, go_c, go_b)
    for a, b in zip(outb[:4], out32[:4]):
        assert torch.equal(a, b)
    assert torch.equal(outb[4], out8[4])
    ref = closed_form_total(logits.float(), img8.float() / 255, labels, 1.0, 0.1, go_b.cpu())
    assert_loss_close(out8[0], ref[0], "total (bf16 / u8 inputs)")
    assert_grad_close(out32[4], ref[4], "gradient (rounded inputs)")


def test_ignore_index_variants(WF):
    gen = torch.Generator().manual_seed(12)
    B, H, W = 2, 48, 64
    logits = torch.randn(B, 2, H, W, generator=gen)
    img = smooth_images(gen, B, H, W)
    go_c, go_b = torch.tensor([0.0]).cuda(), torch.zeros(B).cuda()   # isolate the cross-entropy term
    lab = _labels(gen, B, H, W, p_ignore=0.3)
    x = logits.double().requires_grad_(True)
    ce_ref = F.cross_entropy(x, lab, ignore_index=-100)
    (g_ref,) = torch.autograd.grad(ce_ref, x)
    total, ce, _, _, g = WF.weak_loss_and_grad(logits.cuda(), img.cuda(), lab.cuda(), 1.0, go_c, go_b)
    assert_loss_close(ce, ce_ref, "CE with ignored labels")
    assert_loss_close(total, ce_ref, "total == CE when the pairwise weights are 0")
    assert_grad_close(g, g_ref, "CE gradient with ignored labels")
    lab8 = lab.clone()
    lab8[lab8 == -100] = 255
    _, ce8, _, _, g8 = WF.weak_loss_and_grad(logits.cuda(), img.cuda(), lab8.to(torch.uint8).cuda(), 1.0, go_c, go_b,
                                              ignore_index=255)
    assert torch.equal(ce8, ce) and torch.equal(g8, g)
    everything = torch.full((B, H, W), -100, dtype=torch.int64)
    _, ce_nan, _, _, _ = WF.weak_loss_and_grad(logits.cuda(), img.cuda(), everything.cuda(), 1.0, go_c, go_b)
    assert torch.isnan(ce_nan).all()  # F.cross_entropy: mean over zero labels


def test_module_fast_path_and_fallback(WF):
    """WeakSupervisionLoss: one launch when the tensors allow it, the composed launches otherwise -- same numbers."""
    from weaklysuperviseddl_b200 import WeakSupervisionLoss

    crit = WeakSupervisionLoss()
    gen = torch.Generator().manual_seed(21)
    for (B, H, W) in ((2, 64, 64), (2, 50, 45)):  # 45 columns: rows TMA cannot address -> fallback
        logits = torch.randn(B, 2, H, W, generator=gen)
        img = smooth_images(gen, B, H, W)
        labels = _labels(gen, B, H, W)
        assert WF.weak_loss_supported(logits.cuda(), img.cuda(), labels.cuda()) == (W % 4 == 0)
        x = logits.cuda().requires_grad_(True)
        total, parts = crit(x, img.cuda(), labels.cuda())
        (3.0 * total).backward()
        ref = oracle_total(logits, img, labels, 1.0, 0.1, torch.full((B,), 0.5 / B))
        assert_loss_close(total, ref[0], "module total")
        assert_loss_close(parts["ce"], ref[1])
        assert_loss_close(parts["cut"], ref[2])
        assert_loss_close(parts["boundary"], ref[3].mean())
        assert_grad_close(x.grad, 3.0 * ref[4], "module gradient (upstream 3)")
    # channels-last network output (what a channels_last DeepLab hands over) and autocast dtype
    logits = torch.randn(2, 2, 64, 64, generator=gen).cuda().to(memory_format=torch.channels_last).to(torch.bfloat16)
    x = logits.clone().requires_grad_(True)
    total, _ = crit(x, smooth_images(gen, 2, 64, 64).cuda(), _labels(gen, 2, 64, 64).cuda())
    total.backward()
    assert x.grad.dtype == torch.bfloat16 and x.grad.shape == x.shape and torch.isfinite(x.grad.float()).all()


def test_repeatable_under_load_and_queue_order(WF):
    """Tiles are drawn from a global queue by whichever group is free: the per-tile partials and the fixed-order final
    sum must make the result independent of that order -- bit-identical over repeated launches, also while another
    stream keeps SMs busy, also on two streams at once with their own workspaces."""
    gen = torch.Generator().manual_seed(33)
    side = torch.cuda.Stream()
    busy = torch.randn(2048, 2048, device="cuda")
    for (B, H, W) in ((8, 224, 224), (3, 100, 124), (1, 256, 256), (6, 30, 64)):
        logits = torch.randn(B, 2, H, W, generator=gen).cuda()
        img = smooth_images(gen, B, H, W).cuda()
        labels = _labels(gen, B, H, W).cuda()
        first = WF.weak_loss_and_grad(logits, img, labels)
        for rep in range(6):
            if rep % 2:
                with torch.cuda.stream(side):
                    for _ in range(4):
                        busy = busy @ busy * 1e-3
            again = WF.weak_loss_and_grad(logits, img, labels)
            for a, b in zip(first, again):
                assert torch.equal(a, b), (B, H, W, rep)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            other = WF.weak_loss_and_grad(logits, img, labels)
        here = WF.weak_loss_and_grad(logits, img, labels)
        torch.cuda.synchronize()
        for a, b, c in zip(first, other, here):
            assert torch.equal(a, b) and torch.equal(a, c)


def test_count_valid_labels(WF):
    from weaklysuperviseddl_b200 import _native

    lib = _native.lib()
    gen = torch.Generator().manual_seed(2)
    for n, dtype, ign in ((1000, torch.int64, -100), (224 * 224 * 32, torch.uint8, 255), (7, torch.int64, 1)):
        lab = torch.randint(0, 2, (n,), generator=gen)
        lab[torch.rand(n, generator=gen) < 0.2] = ign
        lab = lab.to(dtype).cuda()
        scratch = torch.empty(2, dtype=torch.int64, device="cuda")
        inv = torch.empty(1, device="cuda")
        rc = lib.wsdl_count_valid_labels(lab.data_ptr(), 3 if dtype == torch.uint8 else 4, n, ign, scratch.data_ptr(),
                                         inv.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        assert inv.item() == np.float32(1.0) / np.float32(int((lab != ign).sum().item()))

"""GPU parity: fused cut / boundary loss forward+backward through the C ABI against the oracle (fp64 arbiter)
and the committed reference outputs; loss and gradient within 1e-5 relative (north_star)."""
import os

import numpy as np
import pytest
import torch

from helpers import assert_grad_close, assert_loss_close, smooth_images
from oracle import wsdl_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def WF():
    from weaklysuperviseddl_b200 import functional

    return functional


@pytest.fixture(scope="module")
def W():
    import weaklysuperviseddl_b200

    return weaklysuperviseddl_b200


@pytest.fixture(scope="module")
def pw(golden_dir):
    return np.load(os.path.join(golden_dir, "pairwise.npz"))


def test_cut_loss_golden(W, pw):
    logits = torch.from_numpy(pw["logits"]).cuda().requires_grad_(True)
    img = torch.from_numpy(pw["images"]).cuda()
    loss = W.LocalNormalizedCutLoss(0.05, 5)(logits, img)
    assert loss.dim() == 0
    loss.backward()
    assert_loss_close(loss, pw["cut_loss"], "cut loss")
    assert_grad_close(logits.grad, pw["cut_grad"], "cut grad")
    x = torch.from_numpy(pw["logits"][0]).cuda().requires_grad_(True)   # 3-D auto-batch (CutLoss.py:72-74)
    loss = W.LocalNormalizedCutLoss(0.05, 5)(x, img[0])
    loss.backward()
    assert_loss_close(loss, pw["cut3d_loss"])
    assert_grad_close(x.grad, pw["cut3d_grad"])
    x = torch.from_numpy(pw["logits"]).cuda().requires_grad_(True)      # window 3, sigma 0.2
    loss = W.LocalNormalizedCutLoss(0.2, 3)(x, img)
    loss.backward()
    assert_loss_close(loss, pw["cut_w3_loss"])
    assert_grad_close(x.grad, pw["cut_w3_grad"])
    x = torch.from_numpy(pw["logits3"]).cuda().requires_grad_(True)     # three classes
    loss = W.LocalNormalizedCutLoss(0.05, 5)(x, torch.from_numpy(pw["images3"]).cuda())
    loss.backward()
    assert_loss_close(loss, pw["cut_c3_loss"])
    assert_grad_close(x.grad, pw["cut_c3_grad"])


def test_boundary_loss_golden(W, pw):
    probs = torch.softmax(torch.from_numpy(pw["logits"]), dim=1)
    img = torch.from_numpy(pw["images"]).cuda()
    mod = W.ConstrainToBoundaryLossSingle(0.1, 5, 5)
    for b in range(2):
        x = probs[b].cuda().requires_grad_(True)
        loss = mod(x, img[b])
        assert loss.dim() == 0
        loss.backward()
        assert_loss_close(loss, pw["boundary_loss"][b])
        assert_grad_close(x.grad, pw["boundary_grad"][b])
    x = probs.cuda().requires_grad_(True)                                # batched extension: (B,) losses
    losses = mod(x, img)
    losses.sum().backward()
    assert_loss_close(losses, pw["boundary_loss"])
    assert_grad_close(x.grad, pw["boundary_grad"])


def test_affinities_golden(W, pw):
    img = torch.from_numpy(pw["images"]).cuda()
    out = W.compute_affinities(img, 0.1, 5, 5)
    assert len(out) == 24 and out[0].shape == (2, 1, 40, 36)
    ours = torch.stack(out, dim=1).squeeze(2).cpu().numpy()
    ref = pw["affinities_batched"]
    assert np.abs(ours - ref).max() <= 1e-5 * np.abs(ref).max()
    rel = np.abs(ours - ref) / np.maximum(ref, 1e-30)
    assert rel[ref > 1e-30].max() <= 1e-5
    single = W.ConstrainToBoundaryLossSingle.compute_affinities_single(img[1], 0.1, 5, 5)
    assert len(single) == 24 and single[0].shape == (1, 40, 36)
    s = torch.stack(single, 0).squeeze(1).cpu().numpy()
    assert np.abs(s - pw["affinities_single"]).max() <= 1e-5 * np.abs(pw["affinities_single"]).max()


CASES = [
    # B, C, H, W, window, sigma_color, sigma_space, softmax, divc, per_image
    (2, 2, 64, 64, 5, 0.05, None, True, True, False),     # all-interior + border tiles
    (1, 2, 224, 224, 5, 0.05, None, True, True, False),   # config-2 image size, cut
    (3, 2, 224, 224, 5, 0.1, 5.0, False, False, True),    # config-2 image size, boundary
    (2, 2, 45, 70, 5, 0.1, 5.0, False, False, True),      # ragged tiles
    (1, 3, 33, 31, 5, 0.05, None, True, True, False),     # three classes
    (1, 5, 20, 50, 5, 0.08, 2.0, True, False, False),     # runtime-C kernel, softmax + spatial
    (2, 2, 19, 23, 3, 0.2, None, True, True, False),      # window 3
    (1, 2, 40, 40, 7, 0.1, 3.0, False, True, True),       # window 7
    (1, 1, 16, 16, 5, 0.1, None, False, True, False),     # single class, no softmax
    (1, 2, 3, 3, 5, 0.1, 5.0, True, True, False),         # smallest legal image for pad=2
    (1, 8, 12, 9, 5, 0.1, None, True, True, False),       # max classes
    # pair-symmetric kernel (window 5, C <= 2): border multiplicities, band columns, ragged widths
    (1, 2, 6, 6, 5, 0.1, 1.0, True, True, False),         # smallest image of the sym kernel: every pixel is in a band
    (1, 2, 6, 6, 5, 0.1, None, False, False, True),
    (2, 2, 7, 9, 5, 0.1, 0.7, False, False, True),        # small gamma: gamma^4 ~ 0.017
    (1, 2, 50, 62, 5, 0.1, 5.0, False, False, True),      # right band straddles two column tiles (59 | 60, 61)
    (1, 2, 50, 61, 5, 0.05, None, True, True, False),     # odd width: scalar staging / stores
    (2, 1, 41, 63, 5, 0.1, 2.0, False, True, False),      # one stored channel without softmax
    (1, 2, 100, 120, 5, 0.1, 5.0, True, False, True),     # softmax + spatial term, two full tiles
    (1, 2, 90, 122, 5, 0.1, 5.0, False, False, True),     # 2-column last tile: its left halo centres feed tile 1
    (3, 2, 39, 44, 5, 0.05, None, True, True, False),     # several images per CTA range, rows split mid-image
    (1, 2, 256, 256, 5, 0.1, None, True, True, False),    # the reference's own call (refine_pseudo_mask, CutLoss.py:745)
    (1, 2, 300, 520, 5, 0.1, 5.0, False, False, True),    # tall/wide: many row blocks and column tiles, boundary form
    (1, 2, 38, 60, 5, 0.1, 5.0, False, False, True),      # exactly one 38-row block and one 60-column tile
    (2, 2, 40, 64, 5, 0.05, None, True, True, False),     # a 2-row last block, a 4-column last tile
    (1, 1, 77, 60, 5, 0.1, 2.0, True, True, False),       # two full blocks + 1 row
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_against_fp64_oracle(WF, case):
    B, C, H, W_, win, sc, ss, sm, divc, per = CASES[case]
    gen = torch.Generator().manual_seed(500 + case)
    vals = torch.randn(B, C, H, W_, generator=gen)
    if not sm:
        vals = torch.softmax(vals * 2, dim=1)
    img = smooth_images(gen, B, H, W_)
    loss, grad = WF.pairwise_loss_and_grad(vals.cuda(), img.cuda(), win, sc, ss, sm, divc, per)
    ref_l, ref_g = [], []
    for b in range(B):
        L, g = O.pairwise_closed_form(vals[b].numpy(), img[b].numpy(), sc, ss, win, sm, divc)
        ref_l.append(L)
        ref_g.append(g)
    ref_g = np.stack(ref_g)
    if per:
        assert_loss_close(loss, np.array(ref_l), f"case {case} loss")
    else:
        assert_loss_close(loss, np.mean(ref_l), f"case {case} loss")
        ref_g = ref_g / B
    assert_grad_close(grad, ref_g, f"case {case} grad")


def test_closed_form_oracle_is_the_reference_math():
    """The fp64 arbiter used above == autograd through the op-for-op restatement (pinned to the reference)."""
    gen = torch.Generator().manual_seed(9)
    vals = torch.randn(1, 2, 17, 13, generator=gen)
    img = smooth_images(gen, 1, 17, 13)
    v, g = O.loss_and_grad(O.cut_loss, vals, img, dtype=torch.float64)
    L, gc = O.pairwise_closed_form(vals[0].numpy(), img[0].numpy(), 0.05, None, 5, True, True)
    assert abs(L - v.item()) < 1e-12 * abs(v.item()) + 1e-18
    assert np.abs(gc - g[0].numpy()).max() < 1e-12 * np.abs(gc).max()


def test_upstream_gradient_and_forward_only(WF, W):
    gen = torch.Generator().manual_seed(21)
    vals = torch.randn(2, 2, 48, 40, generator=gen).cuda()
    img = smooth_images(gen, 2, 48, 40).cuda()
    l1, g1 = WF.pairwise_loss_and_grad(vals, img, 5, 0.05, None, True, True, False)
    go = torch.tensor([0.37], device="cuda")
    l2, g2 = WF.pairwise_loss_and_grad(vals, img, 5, 0.05, None, True, True, False, grad_out=go)
    assert torch.equal(l1, l2)
    assert (g2 - 0.37 * g1).abs().max() <= 1e-6 * g1.abs().max()
    with torch.no_grad():
        l3 = W.LocalNormalizedCutLoss()(vals, img)
    assert torch.equal(l3.reshape(1), l1)
    # autograd: loss * lambda, backward twice (retain_graph) must not double-scale
    x = vals.clone().requires_grad_(True)
    loss = W.LocalNormalizedCutLoss()(x, img) * 0.37
    loss.backward(retain_graph=True)
    first = x.grad.clone()
    x.grad = None
    loss.backward()
    assert torch.equal(first, x.grad)
    assert (first - g2).abs().max() <= 1e-6 * g1.abs().max()
    # per-image upstream gradients
    p = torch.softmax(vals, 1)
    gob = torch.tensor([2.0, -0.5], device="cuda")
    lb, gb = WF.pairwise_loss_and_grad(p, img, 5, 0.1, 5.0, False, False, True, grad_out=gob)
    lb1, gb1 = WF.pairwise_loss_and_grad(p, img, 5, 0.1, 5.0, False, False, True)
    assert (gb - gb1 * gob.view(2, 1, 1, 1)).abs().max() <= 1e-6 * gb1.abs().max()


def test_deterministic(WF):
    gen = torch.Generator().manual_seed(2)
    vals = torch.randn(4, 2, 224, 224, generator=gen).cuda()
    img = smooth_images(gen, 4, 224, 224).cuda()
    a = WF.pairwise_loss_and_grad(vals, img)
    b = WF.pairwise_loss_and_grad(vals, img)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_config2_full_size_properties(WF):
    """32x2x224x224 (BASELINE config 2): properties that need no oracle at full size + one image vs fp64."""
    gen = torch.Generator().manual_seed(1)
    B, C, H, W_ = 32, 2, 224, 224
    logits = torch.randn(B, C, H, W_, generator=gen).cuda()
    img = smooth_images(gen, B, H, W_).cuda()
    loss, grad = WF.pairwise_loss_and_grad(logits, img, 5, 0.05, None, True, True, False)
    assert loss.item() > 0
    # batch mean == mean of per-image losses (the reference's .mean() over b,h,w)
    per = torch.stack([WF.pairwise_loss_and_grad(logits[b:b + 1], img[b:b + 1])[0] for b in range(B)])
    assert abs(per.mean().item() - loss.item()) <= 1e-6 * loss.item()
    # softmax shift invariance: adding a per-pixel constant to all logits changes nothing
    shift = torch.randn(B, 1, H, W_, generator=torch.Generator().manual_seed(3)).cuda()
    loss_s, grad_s = WF.pairwise_loss_and_grad(logits + shift, img)
    assert abs(loss_s.item() - loss.item()) <= 2e-6 * loss.item()
    # gradient of a softmax-composed loss sums to zero over classes
    assert grad.sum(dim=1).abs().max().item() <= 1e-6 * grad.abs().max().item()
    # constant image => affinity 1 everywhere; uniform predictions => zero loss and gradient
    l0, g0 = WF.pairwise_loss_and_grad(torch.zeros_like(logits), torch.full_like(img, 0.5))
    assert l0.item() == 0.0 and g0.abs().max().item() == 0.0
    b = 17
    L, g = O.pairwise_closed_form(logits[b].cpu().numpy(), img[b].cpu().numpy(), 0.05, None, 5, True, True)
    assert_loss_close(per[b], L)
    assert_grad_close(grad[b] * B, g)
    # boundary loss on the same batch, one launch for 32 images
    probs = torch.softmax(logits, 1)
    lb, gb = WF.pairwise_loss_and_grad(probs, img, 5, 0.1, 5.0, False, False, True)
    L, g = O.pairwise_closed_form(probs[b].cpu().numpy(), img[b].cpu().numpy(), 0.1, 5.0, 5, False, False)
    assert_loss_close(lb[b], L)
    assert_grad_close(gb[b], g)


def test_prepared_workspace_is_self_cleaning(WF):
    """wsdl_pairwise_fwd_bwd_prepared on one workspace, launch after launch (different kernels in between), must
    equal wsdl_pairwise_fwd_bwd (which zeroes the ticket itself) bit for bit."""
    from weaklysuperviseddl_b200 import _native

    lib = _native.lib()
    gen = torch.Generator().manual_seed(77)
    shapes = [(2, 2, 64, 64, 5), (1, 3, 33, 31, 5), (2, 2, 19, 23, 3), (2, 2, 64, 64, 5), (1, 2, 224, 224, 5)]
    nmax = max(lib.wsdl_pairwise_workspace_bytes(B, H, W_) for (B, C, H, W_, win) in shapes)
    ws = torch.empty(nmax, dtype=torch.uint8, device="cuda").fill_(0xAB)  # garbage until initialised
    sp = torch.cuda.current_stream().cuda_stream
    assert lib.wsdl_pairwise_workspace_init(ws.data_ptr(), nmax, sp) == 0
    for (B, C, H, W_, win) in shapes * 2:
        vals = torch.randn(B, C, H, W_, generator=gen).cuda()
        img = smooth_images(gen, B, H, W_).cuda()
        out = []
        for prepared in (True, False):
            loss = torch.empty(1, device="cuda")
            grad = torch.empty_like(vals)
            w = ws if prepared else torch.empty(nmax, dtype=torch.uint8, device="cuda").fill_(0xCD)
            fn = lib.wsdl_pairwise_fwd_bwd_prepared if prepared else lib.wsdl_pairwise_fwd_bwd
            rc = fn(vals.data_ptr(), img.data_ptr(), B, C, H, W_, win, 0.05, 0.0, 1, 1, 0, None, loss.data_ptr(),
                    grad.data_ptr(), w.data_ptr(), nmax, sp)
            assert rc == 0
            out.append((loss, grad))
        assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
    torch.cuda.synchronize()


def test_module_under_cuda_graph_capture(W):
    """The modules can be captured in a CUDA graph (workspace allocated inside the capture is not cached) and the
    replay reproduces the eager result."""
    gen = torch.Generator().manual_seed(31)
    vals = torch.randn(2, 2, 48, 64, generator=gen).cuda()
    img = smooth_images(gen, 2, 48, 64).cuda()
    crit = W.LocalNormalizedCutLoss()
    with torch.no_grad():
        eager = crit(vals, img).clone()
    static_in = vals.clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        crit(static_in, img)  # warm-up on the side stream
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g), torch.no_grad():
        out = crit(static_in, img)
    for scale in (1.0, 0.5):
        static_in.copy_(vals * scale)
        g.replay()
        with torch.no_grad():
            ref = crit(vals * scale, img)
        assert torch.equal(out.reshape(-1), ref.reshape(-1))
    assert torch.equal(eager.reshape(-1), crit(vals, img).detach().reshape(-1))


DUAL_CASES = [(2, 64, 64), (1, 224, 224), (2, 45, 70), (1, 6, 6), (3, 39, 44), (1, 100, 122), (2, 50, 61), (1, 256, 256),
              (2, 300, 520), (5, 7, 9),
              # row blocks of the 5-row-segment geometry (38 owned rows): exactly one block and one column tile, a 2-row
              # last block, two full blocks + 1 row
              (1, 38, 60), (2, 40, 64), (1, 77, 60),
              # the fused kernel's own geometry (segments of up to 6 rows, 46 owned rows per block): exactly one block, a
              # 2-row last block, two full blocks + 1 row; 3-row segments (heads in the unused tail of the planes)
              (1, 46, 60), (2, 48, 64), (1, 93, 60), (2, 22, 30)]


@pytest.mark.parametrize("case", range(len(DUAL_CASES)))
def test_dual_launch_equals_the_two_losses(WF, case):
    """wsdl_pairwise_dual_fwd_bwd == cut loss on the logits + boundary loss on softmax(logits), values and the
    gradient w.r.t. the logits (with upstream gradients), against the separate launches and the fp64 oracle."""
    B, H, W_ = DUAL_CASES[case]
    gen = torch.Generator().manual_seed(900 + case)
    logits = torch.randn(B, 2, H, W_, generator=gen)
    img = smooth_images(gen, B, H, W_)
    go_c = torch.tensor([0.7])
    go_b = torch.rand(B, generator=gen) + 0.5
    lc, lb, g = WF.pairwise_dual_loss_and_grad(logits.cuda(), img.cuda(), 0.05, 0.1, 5.0, 5, go_c.cuda(), go_b.cuda())
    # fp64 oracle, image by image
    ref_g = np.zeros((B, 2, H, W_))
    ref_lc, ref_lb = [], []
    for b in range(B):
        Lc, gc = O.pairwise_closed_form(logits[b].numpy(), img[b].numpy(), 0.05, None, 5, True, True)
        p = torch.softmax(logits[b].double(), 0).numpy()
        Lb, gp = O.pairwise_closed_form(p, img[b].numpy(), 0.1, 5.0, 5, False, False)
        gz = p * (gp - (p * gp).sum(0, keepdims=True))  # through the softmax
        ref_lc.append(Lc)
        ref_lb.append(Lb)
        ref_g[b] = go_c.item() * gc / B + go_b[b].item() * gz
    assert_loss_close(lc, np.mean(ref_lc), f"dual case {case} cut loss")
    assert_loss_close(lb, np.array(ref_lb), f"dual case {case} boundary loss")
    assert_grad_close(g, ref_g, f"dual case {case} grad")
    # and the separate launches agree with the fused one to rounding
    l1, _ = WF.pairwise_loss_and_grad(logits.cuda(), img.cuda(), 5, 0.05, None, True, True, False)
    l2, _ = WF.pairwise_loss_and_grad(torch.softmax(logits.cuda(), 1), img.cuda(), 5, 0.1, 5.0, False, False, True)
    assert abs(l1.item() - lc.item()) <= 2e-6 * abs(l1.item())
    assert (l2 - lb).abs().max().item() <= 2e-6 * l2.abs().max().item()
    # forward only, no upstream gradients
    lc2, lb2, none = WF.pairwise_dual_loss_and_grad(logits.cuda(), img.cuda(), want_grad=False)
    assert none is None and torch.equal(lc2, lc) and torch.equal(lb2, lb)


def test_repeatability_stress(WF):
    """Races (named-barrier hand-offs, carries added after the barrier, band pass, loss ticket) would show up as
    run-to-run differences: every variant must reproduce itself bit for bit over repeated launches, also while another
    stream keeps the GPU busy."""
    gen = torch.Generator().manual_seed(123)
    side = torch.cuda.Stream()
    busy_a = torch.randn(2048, 2048, device="cuda")
    shapes = [(4, 2, 224, 224), (2, 2, 45, 70), (3, 2, 100, 122), (1, 2, 256, 256), (6, 2, 30, 64)]
    for (B, C, H, W_) in shapes:
        logits = torch.randn(B, C, H, W_, generator=gen).cuda()
        probs = torch.softmax(logits, 1)
        img = smooth_images(gen, B, H, W_).cuda()
        ref = None
        for rep in range(6):
            if rep % 2:
                with torch.cuda.stream(side):
                    for _ in range(4):
                        busy_a = busy_a @ busy_a * 1e-3
            cut = WF.pairwise_loss_and_grad(logits, img, 5, 0.05, None, True, True, False)
            bnd = WF.pairwise_loss_and_grad(probs, img, 5, 0.1, 5.0, False, False, True)
            dual = WF.pairwise_dual_loss_and_grad(logits, img)
            cur = [t.clone() for t in (*cut, *bnd, *dual)]
            if ref is None:
                ref = cur
            else:
                for a, b in zip(ref, cur):
                    assert torch.equal(a, b), (B, H, W_, rep)
    torch.cuda.synchronize()


def test_prepared_workspaces_survive_interleaved_kernels_and_shapes(WF):
    """The functional layer caches prepared workspaces.  Calls that take different kernels (pair-symmetric: window 5;
    generic: window 3 / three classes) and different shapes, interleaved on one stream, must each reproduce what they
    returned the first time: no kernel may find another call's partials where it expects its own empty result slots."""
    gen = torch.Generator().manual_seed(321)
    cases = []
    for (B, C, H, W_, window) in [(2, 2, 64, 64, 5), (2, 2, 64, 64, 3), (1, 3, 64, 64, 5), (3, 2, 40, 96, 5),
                                  (3, 2, 40, 96, 3), (1, 2, 224, 224, 5), (2, 2, 64, 64, 7)]:
        vals = torch.randn(B, C, H, W_, generator=gen).cuda()
        img = smooth_images(gen, B, H, W_).cuda()
        cases.append((vals, img, window))
    first = {}
    for rep in range(3):
        for i, (vals, img, window) in enumerate(cases):
            out = WF.pairwise_loss_and_grad(vals, img, window, 0.05, None, True, True, False)
            outs = [t.clone() for t in out]
            if vals.shape[1] == 2 and window == 5:
                outs += [t.clone() for t in WF.pairwise_dual_loss_and_grad(vals, img)]
            if i in first:
                for a, b in zip(first[i], outs):
                    assert torch.equal(a, b), (i, rep)
            else:
                first[i] = outs
    torch.cuda.synchronize()


def _eager_pair_loss(p, img, sigma_color, sigma_space, mean_over_classes):
    """The reference's formula with plain torch ops, any device / dtype: 24 reflect-shifted copies."""
    import torch.nn.functional as F

    H, W_ = p.shape[-2:]
    pp, ip = F.pad(p, (2,) * 4, mode="reflect"), F.pad(img, (2,) * 4, mode="reflect")
    total = 0.0
    for dy in range(-2, 3):
        for dx in range(-2, 3):
            if dy == 0 and dx == 0:
                continue
            ps, isf = pp[:, :, 2 + dy:2 + dy + H, 2 + dx:2 + dx + W_], ip[:, :, 2 + dy:2 + dy + H, 2 + dx:2 + dx + W_]
            e = -((img - isf) ** 2).sum(1, keepdim=True) / (2 * sigma_color ** 2)
            if sigma_space:
                e = e - (dx * dx + dy * dy) / (2 * sigma_space ** 2)
            d2 = (p - ps) ** 2
            total = total + (torch.exp(e) * (d2.mean(1, keepdim=True) if mean_over_classes else d2.sum(1, keepdim=True))).mean()
    return total / 24


def test_large_image_against_eager_fp64(WF):
    """1024 x 1024 (the largest size BASELINE names): 27 row blocks x 18 column tiles per image; fused launch and the
    single-loss launches against an fp64 eager evaluation of the reference's formula on the GPU."""
    gen = torch.Generator().manual_seed(21)
    B, H, W_ = 2, 1024, 1024
    logits = torch.randn(B, 2, H, W_, generator=gen).cuda()
    img = smooth_images(gen, B, H, W_).cuda()
    x = logits.double().requires_grad_(True)
    p = torch.softmax(x, 1)
    cut = _eager_pair_loss(p, img.double(), 0.05, None, True)
    bnd = torch.stack([_eager_pair_loss(p[b:b + 1], img[b:b + 1].double(), 0.1, 5.0, False) for b in range(B)])
    go_b = torch.tensor([0.5, 1.5], dtype=torch.float64, device="cuda")
    (0.7 * cut + (go_b * bnd).sum()).backward()
    lc, lb, g = WF.pairwise_dual_loss_and_grad(logits, img, 0.05, 0.1, 5.0, 5, torch.tensor([0.7]).cuda(), go_b.float())
    assert abs(lc.item() - cut.item()) <= 1e-5 * abs(cut.item())
    assert (lb.double() - bnd.detach()).abs().max().item() <= 1e-5 * bnd.abs().max().item()
    assert (g.double() - x.grad).abs().max().item() <= 1e-5 * x.grad.abs().max().item()
    l1, g1 = WF.pairwise_loss_and_grad(logits, img, 5, 0.05, None, True, True, False)
    assert abs(l1.item() - cut.item()) <= 1e-5 * abs(cut.item())

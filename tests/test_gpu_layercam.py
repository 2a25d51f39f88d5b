"""GPU parity: fused LayerCAM -> normalise -> upsample -> fuse -> threshold (+ keep_largest) through the C ABI
against the oracle and the committed reference outputs.  CAMs within 1e-5 relative (north_star); masks
bit-exact outside the 1e-6 threshold band, pixels inside counted."""
import os

import numpy as np
import pytest
import torch

from helpers import assert_cam_close, assert_masks_match, synth_act_grad
from oracle import wsdl_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def WF():
    from weaklysuperviseddl_b200 import functional

    return functional


@pytest.fixture(scope="module")
def lc(golden_dir):
    return np.load(os.path.join(golden_dir, "layercam_tiny.npz"))


def _cuda(ts):
    return [t.cuda() for t in ts]


@pytest.mark.parametrize("alpha", [1.0, 0.5])
def test_golden_main(WF, lc, alpha):
    names = ("layer3", "layer4")
    for i in range(2):
        acts = [torch.from_numpy(lc[f"main_act_{n}_{i}"]) for n in names]
        grads = [torch.from_numpy(lc[f"main_grad_{n}_{i}"]) for n in names]
        cam, mask, near = WF.layercam_fused(_cuda(acts), _cuda(grads), (224, 224), alpha=alpha, thresh=0.3)
        ref = lc[f"main_cam_alpha{alpha}"][i]
        assert_cam_close(cam[0], ref, f"golden main alpha={alpha} img {i}")
        ref64 = O.layercam_from_hooks(acts, grads, (224, 224), alpha=alpha, dtype=torch.float64)[0]
        inside = assert_masks_match(mask[0], ref64, 0.3)
        assert int(near.item()) <= inside + 8  # the kernel's own band count is of the same order


@pytest.mark.parametrize("alpha", [1.0, 2.0])
def test_golden_variant(WF, lc, alpha):
    names = ("layer2", "layer3", "layer4")
    for i in range(2):
        acts = [torch.from_numpy(lc[f"variant_act_{n}_{i}"]) for n in names]
        grads = [torch.from_numpy(lc[f"variant_grad_{n}_{i}"]) for n in names]
        cam, _, _ = WF.layercam_fused(_cuda(acts), _cuda(grads), (224, 224), alpha=alpha, alpha_mode=1)
        assert_cam_close(cam[0], lc[f"variant_cam_alpha{alpha}"][i], f"golden variant alpha={alpha} img {i}")


SHAPES = [
    # (B, [(C,h,w)...], out)  -- config 1 (224: layer3+4), config 3 (512), 4-stage, odd sizes, downsampling
    (8, [(1024, 14, 14), (2048, 14, 14)], (224, 224)),
    (3, [(1024, 32, 32), (2048, 32, 32)], (512, 512)),
    (2, [(64, 64, 64), (128, 32, 32), (256, 16, 16), (512, 16, 16)], (256, 256)),
    (1, [(37, 7, 9), (5, 3, 3)], (224, 224)),          # hw not a multiple of 4 -> scalar path
    (2, [(16, 40, 40)], (24, 31)),                      # downsampling, single layer
    (1, [(8, 6, 5), (8, 12, 10), (8, 3, 3), (4, 5, 5), (4, 2, 2)], (64, 48)),  # 5 layers -> generic upsample
    (1, [(2048, 1, 1), (3, 14, 14)], (224, 224)),       # constant map: max-min = 0 -> 0/1e-8
]


@pytest.mark.parametrize("case", range(len(SHAPES)))
@pytest.mark.parametrize("alpha,mode", [(1.0, 0), (0.5, 0), (2.0, 1), (1.7, 0)])
def test_random_shapes(WF, case, alpha, mode):
    B, layers, out = SHAPES[case]
    gen = torch.Generator().manual_seed(100 + case)
    acts, grads = [], []
    for (C, h, w) in layers:
        a, g = synth_act_grad(gen, B, C, h, w)
        acts.append(a)
        grads.append(g)
    thr = 0.3
    cam, mask, near = WF.layercam_fused(_cuda(acts), _cuda(grads), out, alpha=alpha, alpha_mode=mode, thresh=thr)
    ref32 = O.layercam_from_hooks(acts, grads, out, alpha=alpha, alpha_mode=mode)
    ref64 = O.layercam_from_hooks(acts, grads, out, alpha=alpha, alpha_mode=mode, dtype=torch.float64)
    assert_cam_close(cam, ref32, f"case {case} vs fp32 oracle")
    assert_cam_close(cam, ref64, f"case {case} vs fp64 oracle")
    assert_masks_match(mask, ref64, thr)
    # mask-only call (the fused pseudo-mask path: no CAM in HBM) gives the same bits
    _, mask2, _ = WF.layercam_fused(_cuda(acts), _cuda(grads), out, alpha=alpha, alpha_mode=mode, thresh=thr,
                                    want_cam=False)
    assert torch.equal(mask, mask2)


# small batches take the one-launch cluster path (csrc/layercam.cu: layercam_cluster_kernel): 16 CTAs per image meet
# through distributed shared memory.  Eligible: <= 4 layers, low-resolution maps of <= 1024 pixels (a multiple of the
# 128-bit vector), <= 12 MB of hooks per image, B <= 16.
CLUSTER_SHAPES = [
    (1, [(1024, 14, 14), (2048, 14, 14)], (224, 224)),                       # the reference's own call: one image
    (8, [(1024, 14, 14), (2048, 14, 14)], (224, 224)),                       # BASELINE config 1
    (5, [(512, 28, 28), (1024, 14, 14), (2048, 14, 14)], (224, 224)),        # the variant's three layers
    (2, [(64, 16, 16)], (100, 75)),                                          # single layer, ragged output
    (4, [(32, 32, 32), (64, 16, 16), (128, 8, 8), (256, 8, 8)], (256, 256)), # four stages
    (16, [(8, 2, 2), (40, 4, 4)], (33, 1021)),                               # tiny maps, widest output, fewer channels than ranks
    (3, [(1, 8, 8), (2, 4, 4), (700, 12, 12)], (64, 64)),                    # one channel: most ranks of the cluster idle
]


@pytest.mark.parametrize("case", range(len(CLUSTER_SHAPES)))
@pytest.mark.parametrize("alpha,mode", [(1.0, 0), (0.5, 0), (2.0, 1)])
def test_small_batch_cluster_path(WF, case, alpha, mode):
    B, layers, out = CLUSTER_SHAPES[case]
    gen = torch.Generator().manual_seed(500 + case)
    acts, grads = [], []
    for (C, h, w) in layers:
        a, g = synth_act_grad(gen, B, C, h, w)
        acts.append(a)
        grads.append(g)
    cam, mask, near = WF.layercam_fused(_cuda(acts), _cuda(grads), out, alpha=alpha, alpha_mode=mode, thresh=0.3)
    ref64 = O.layercam_from_hooks(acts, grads, out, alpha=alpha, alpha_mode=mode, dtype=torch.float64)
    assert_cam_close(cam, ref64, f"cluster case {case}")
    assert_masks_match(mask, ref64, 0.3)
    assert int(near.item()) == int(((cam - 0.3).abs() < 1e-6).sum().item())
    _, mask2, _ = WF.layercam_fused(_cuda(acts), _cuda(grads), out, alpha=alpha, alpha_mode=mode, thresh=0.3, want_cam=False)
    assert torch.equal(mask, mask2)
    cam3, _, _ = WF.layercam_fused(_cuda(acts), _cuda(grads), out, alpha=alpha, alpha_mode=mode)
    assert torch.equal(cam, cam3)  # deterministic: peers are added in rank order


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_half_inputs_cluster_path(WF, dtype):
    gen = torch.Generator().manual_seed(17)
    acts, grads = [], []
    for (C, h, w) in [(256, 16, 16), (512, 8, 8)]:
        a, g = synth_act_grad(gen, 3, C, h, w)
        acts.append(a.to(dtype))
        grads.append((g * 100).to(dtype))
    cam, _, _ = WF.layercam_fused(_cuda(acts), _cuda(grads), (128, 128))
    ref = O.layercam_from_hooks([a.float() for a in acts], [g.float() for g in grads], (128, 128), dtype=torch.float64)
    assert_cam_close(cam, ref, str(dtype))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_half_inputs(WF, dtype):
    gen = torch.Generator().manual_seed(7)
    acts, grads = [], []
    for (C, h, w) in [(256, 32, 32), (512, 14, 14)]:
        a, g = synth_act_grad(gen, 2, C, h, w)
        acts.append(a.to(dtype))
        grads.append((g * 100).to(dtype))
    cam, _, _ = WF.layercam_fused(_cuda(acts), _cuda(grads), (128, 128))
    # reference semantics on the rounded inputs, accumulated in fp32
    ref = O.layercam_from_hooks([a.float() for a in acts], [g.float() for g in grads], (128, 128), dtype=torch.float64)
    assert_cam_close(cam, ref, str(dtype))


def test_deterministic_and_batch_independent(WF):
    gen = torch.Generator().manual_seed(11)
    a, g = synth_act_grad(gen, 6, 1024, 14, 14)
    a2, g2 = synth_act_grad(gen, 6, 2048, 14, 14)
    A, G = _cuda([a, a2]), _cuda([g, g2])
    c1, m1, _ = WF.layercam_fused(A, G, (224, 224), thresh=0.3)
    c2, m2, _ = WF.layercam_fused(A, G, (224, 224), thresh=0.3)
    assert torch.equal(c1, c2) and torch.equal(m1, m2)
    # image 4 alone == image 4 inside the batch up to the summation split (different plan): still within tolerance
    c4, _, _ = WF.layercam_fused([x[4:5].clone() for x in A], [x[4:5].clone() for x in G], (224, 224))
    assert_cam_close(c4[0], c1[4], "batch independence")


def test_threshold_edge_values(WF):
    cam = torch.tensor([0.0, 0.29999, 0.3, 0.3000001, 0.31, float("nan"), 1.0, -0.0], device="cuda")
    for thr in (0.3, 0.0):
        m, near = WF.threshold_mask(cam, thr)
        assert m.cpu().numpy().tolist() == O.threshold_mask(cam.cpu()[None], thr)[0].tolist()
    m, near = WF.threshold_mask(cam, 0.3)
    assert int(near.item()) == 2  # 0.3 and 0.3000001


def test_full_size_properties(WF):
    """Config-3 sized batch (512x512): properties that need no oracle at full size."""
    gen = torch.Generator(device="cuda").manual_seed(0)
    B = 16
    acts = [torch.randn(B, 1024, 32, 32, device="cuda", generator=gen).relu_(),
            torch.randn(B, 2048, 32, 32, device="cuda", generator=gen).relu_()]
    grads = [torch.randn(B, 1024, 32, 32, device="cuda", generator=gen) * 1e-3,
             torch.randn(B, 2048, 32, 32, device="cuda", generator=gen) * 1e-3]
    cam, mask, near = WF.layercam_fused(acts, grads, (512, 512), thresh=0.3)
    assert cam.min().item() >= 0.0 and cam.max().item() <= 1.0 + 1e-6
    assert torch.equal(mask, ((cam >= 0.3) & (cam > 0)).to(torch.uint8))           # mask is the threshold of the CAM
    assert int(near.item()) == int(((cam - 0.3).abs() < 1e-6).sum().item())
    perm = torch.randperm(B, device="cuda")                                         # per-image independence
    cam_p, _, _ = WF.layercam_fused([a[perm] for a in acts], [g[perm] for g in grads], (512, 512))
    assert torch.equal(cam_p, cam[perm])
    cam_s, _, _ = WF.layercam_fused(acts, [g * 8.0 for g in grads], (512, 512))   # scale invariance of min-max
    assert (cam_s - cam).abs().max().item() < 1e-5
    # spot-check two images against the fp64 oracle
    for b in (0, B - 1):
        ref = O.layercam_from_hooks([a[b:b + 1].cpu() for a in acts], [g[b:b + 1].cpu() for g in grads], (512, 512),
                                    dtype=torch.float64)
        assert_cam_close(cam[b], ref[0], f"full-size image {b}")
        assert_masks_match(mask[b], ref[0], 0.3)


# ------------------------------------------------------------------ keep_largest
def test_keep_largest_matches_oracle(WF):
    rng = np.random.default_rng(3)
    cases = []
    for density, shape in ((0.3, (64, 64)), (0.5, (224, 224)), (0.62, (97, 131)), (0.45, (512, 512)), (0.95, (40, 33))):
        cases.append((rng.random(shape) < density).astype(np.uint8))
    cases.append(np.zeros((32, 32), np.uint8))          # empty
    cases.append(np.ones((50, 20), np.uint8))           # full
    cases.append(np.eye(48, dtype=np.uint8))            # diagonal: 8-connectivity
    tie = np.zeros((5, 9), np.uint8)
    tie[3, 0:2] = 1
    tie[1, 5:7] = 1
    cases.append(tie)
    spiral = np.zeros((65, 65), np.uint8)               # long thin component: deep union-find chains
    for k in range(0, 32, 2):
        spiral[k, k:65 - k] = 1
        spiral[k:65 - k, 64 - k] = 1
        spiral[64 - k, k:65 - k] = 1
        spiral[k + 2:65 - k, k] = 1
    cases.append(spiral)
    cb = np.zeros((70, 100), np.uint8)                  # 256 components per 32x32 tile, 16 runs in every other row
    cb[::2, ::2] = 1
    cb[10, 10:15] = 1                                   # ... and one of five pixels that wins
    cases.append(cb)
    stripes = np.zeros((64, 96), np.uint8)              # 16 runs per row, every run spans the tile seams
    stripes[:, ::2] = 1
    stripes[31:33, 40:47] = 1
    cases.append(stripes)
    anti = np.zeros((96, 96), np.uint8)                 # a diagonal that crosses tile corners up-right
    anti[np.arange(96), 95 - np.arange(96)] = 1
    cases.append(anti)
    yy, xx = np.mgrid[0:200, 0:208]                     # CAM-like blobs over several tiles (smooth outlines)
    cases.append((((yy - 60) ** 2 + (xx - 70) ** 2 < 50 ** 2) | ((yy - 150) ** 2 / 4 + (xx - 150) ** 2 < 30 ** 2)).astype(np.uint8))
    for m in cases:
        out, area = WF.keep_largest(torch.from_numpy(m).cuda(), return_area=True)
        ref = O.keep_largest(m)
        assert np.array_equal(out.cpu().numpy(), ref), f"shape {m.shape}"
        assert int(area.item()) == int(ref.sum())


def test_keep_largest_batched_and_idempotent(WF):
    rng = np.random.default_rng(4)
    m = (rng.random((5, 224, 224)) < 0.55).astype(np.uint8)
    out = WF.keep_largest(torch.from_numpy(m).cuda())
    for b in range(5):
        assert np.array_equal(out[b].cpu().numpy(), O.keep_largest(m[b]))
    assert torch.equal(WF.keep_largest(out), out)

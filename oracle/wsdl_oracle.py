"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this file, and only as the *checker* or the
*timed CPU baseline*; the product package (`weaklysuperviseddl_b200/`) never
does.  Parity status: PINNED (except `keep_largest`: UNPINNED, skimage cannot be had -- see
`label_two_pass`) -- `tests/test_oracle.py` runs every
function below against the real reference functions loaded from /root/reference
(in the authoring container) and against the fixtures in `tests/golden/` that
`oracle/make_golden.py` produced from the real reference (those travel to the
GPU box).  The reference itself holds no tests or golden vectors.

Each function cites the reference lines it restates (paths relative to
/root/reference/TraditionalModel).  Every function takes `dtype`: float32
reproduces the reference's arithmetic op for op (same ATen op sequence, so it
is also an honest CPU timing of the reference path); float64 is the arbiter for
near-threshold pixels and for 1e-5 relative checks.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# Stage 1: LayerCAM (LayerCAM.py:52-76; variant AlternatingDirectionCutLoss.py:262-286)
# --------------------------------------------------------------------------------------


def lowres_cam(act: torch.Tensor, grad: torch.Tensor) -> torch.Tensor:
    """relu(sum_c relu(grad*act))  -- LayerCAM.py:57-59.  (B,C,h,w) -> (B,h,w)."""
    return torch.relu(torch.relu(grad * act).sum(dim=1))


def minmax_rows_(cam: torch.Tensor) -> torch.Tensor:
    """Per-image in-place min-max: c -= min; c /= (max(c) + 1e-8)  -- LayerCAM.py:62-67.

    NB the max is taken *after* the subtraction, and the 1e-8 is added in the
    tensor's dtype (a no-op in fp32 unless max is tiny)."""
    for b in range(cam.shape[0]):
        row = cam[b]
        row -= row.min()
        row /= (row.max() + 1e-8)
    return cam


def bilinear_axis(in_size: int, out_size: int, dtype: torch.dtype):
    """Source indices / lambdas of torch's align_corners=False bilinear resize
    (ATen/native/UpSample.h area_pixel_compute_source_index + guard_index_and_lambda):
    scale = in/out in the op-math type; src = max(scale*(dst+0.5)-0.5, 0);
    i0 = floor(src); i1 = i0 + (i0 < in-1); l1 = src - i0; l0 = 1 - l1."""
    one = torch.ones((), dtype=dtype)
    scale = (one * in_size) / out_size
    dst = torch.arange(out_size, dtype=dtype)
    src = torch.clamp(scale * (dst + 0.5) - 0.5, min=0.0)
    i0 = src.floor().to(torch.int64).clamp(max=in_size - 1)
    i1 = i0 + (i0 < in_size - 1).to(torch.int64)
    l1 = (src - i0.to(dtype)).clamp(min=0.0, max=1.0)
    l0 = 1.0 - l1
    return i0, i1, l0, l1


def bilinear_resize(x: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """(B,h,w) -> (B,out_h,out_w); restates F.interpolate(mode='bilinear',
    align_corners=False) as used at LayerCAM.py:69 (size there is (224,224))."""
    _, h, w = x.shape
    y0, y1, ly0, ly1 = bilinear_axis(h, out_h, x.dtype)
    x0, x1, lx0, lx1 = bilinear_axis(w, out_w, x.dtype)
    top = x[:, y0][:, :, x0] * lx0 + x[:, y0][:, :, x1] * lx1
    bot = x[:, y1][:, :, x0] * lx0 + x[:, y1][:, :, x1] * lx1
    return top * ly0[None, :, None] + bot * ly1[None, :, None]


def layercam_from_hooks(
    acts: Sequence[torch.Tensor],
    grads: Sequence[torch.Tensor],
    out_size: Tuple[int, int] = (224, 224),
    alpha: float = 1.0,
    alpha_mode: int = 0,
    dtype: torch.dtype = torch.float32,
    use_torch_interpolate: bool = True,
) -> torch.Tensor:
    """Everything `generate` does after the backbone fwd/bwd.

    alpha_mode 0: LayerCAM.py:52-76  (mean of layers -> clamp(0) ** alpha)
    alpha_mode 1: AlternatingDirectionCutLoss.py:262-286 (per layer: normalise,
                  ** alpha, normalise again; plain mean; no final clamp/pow)
    Returns (B, out_h, out_w)."""
    per_layer: List[torch.Tensor] = []
    for a, g in zip(acts, grads):
        a = a.detach().to(dtype)
        g = g.detach().to(dtype)
        cam = lowres_cam(a, g)
        if alpha_mode == 0:
            minmax_rows_(cam)
        else:
            for b in range(cam.shape[0]):
                row = cam[b]
                row -= row.min()
                row /= (row.max() + 1e-8)
                row = row ** alpha
                row -= row.min()
                row /= (row.max() + 1e-8)
                cam[b] = row
        if use_torch_interpolate:
            up = F.interpolate(cam.unsqueeze(1), size=tuple(out_size), mode="bilinear", align_corners=False).squeeze(1)
        else:
            up = bilinear_resize(cam, out_size[0], out_size[1])
        per_layer.append(up)
    fused = sum(per_layer) / len(per_layer)  # python sum: 0 + l0 + l1 ... (LayerCAM.py:74)
    if alpha_mode == 0:
        fused = fused.clamp(min=0.0) ** alpha  # LayerCAM.py:76
    return fused


def bg_cam_from_fg(all_cams: torch.Tensor, alpha: float = 2.0, size: Tuple[int, int] = (224, 224)):
    """generate_bg_cam tail, AlternatingDirectionCutLoss.py:305-318: max over dim 0,
    m_bg = 1 - clamp(1-max,0)**alpha, both bilinearly resized to `size`."""
    max_obj, _ = all_cams.max(dim=0)
    m_bg = 1.0 - ((1.0 - max_obj).clamp(min=0.0) ** alpha)
    rs = lambda t: F.interpolate(t[None, None], size=tuple(size), mode="bilinear", align_corners=False).squeeze()
    return rs(m_bg), rs(max_obj)


# --------------------------------------------------------------------------------------
# Stage 2: pseudo-mask thresholding / label assembly (PsuedoMasks.py:15-21,59-74)
# --------------------------------------------------------------------------------------


def threshold_mask(cam: torch.Tensor, cam_thresh: float) -> np.ndarray:
    """cam[cam < thr] = 0; mask = (cam > 0) as uint8  -- PsuedoMasks.py:59-62."""
    cam = cam.clone()
    cam[cam < cam_thresh] = 0.0
    return (cam.cpu().numpy() > 0).astype(np.uint8)


def keep_largest(mask: np.ndarray) -> np.ndarray:
    """PsuedoMasks.py:15-21 with skimage.measure.label's defaults restated:
    full (8-) connectivity, background 0, labels in raster order of first pixel;
    `max(regions, key=area)` returns the FIRST maximum, i.e. the lowest label.
    skimage is not installed here, so this restatement is pinned only against
    scipy.ndimage, cv2 and the independent two-pass restatement of skimage's own
    algorithm below (`keep_largest_two_pass`), which all agree -- parity with
    skimage proper is UNPINNED (see `label_two_pass`)."""
    from scipy import ndimage

    labeled, n = ndimage.label(mask != 0, structure=np.ones((3, 3), dtype=np.int32))
    if n == 0:
        return mask
    areas = np.bincount(labeled.ravel(), minlength=n + 1)[1:]
    best = int(np.argmax(areas)) + 1  # argmax returns the first maximum
    return (labeled == best).astype(np.uint8)


def label_two_pass(mask: np.ndarray) -> np.ndarray:
    """SECOND, independent restatement of `skimage.measure.label(mask)` with its defaults (connectivity = ndim
    -> 8-connectivity in 2-D, background 0), written from skimage's published algorithm (measure/_ccomp.pyx:
    `_label` -- a raster scan that joins every foreground pixel with its already-visited neighbours W, NW, N, NE
    in a union-find forest whose `join_trees` always makes the SMALLER raster index the root -- followed by
    `resolve_labels`, a second raster scan that hands out consecutive labels 1, 2, ... the moment it meets a
    root).  Because a component's root is its minimum raster index, labels come out in raster order of each
    component's first pixel.  Pure numpy/Python, no scipy: it shares no code with `keep_largest` above.

    skimage itself is absent from the reference tree, from this image and from the offline wheelhouse
    (`pip download scikit-image --no-index --find-links /opt/wheelhouse` finds no distribution), and the reference
    pins no version (no requirements file): parity of keep_largest is therefore UNPINNED against skimage proper
    and rests on two independent restatements that must agree (tests/test_oracle.py)."""
    m = np.asarray(mask) != 0
    H, W = m.shape
    parent = np.arange(H * W, dtype=np.int64)

    def find(i):
        r = i
        while parent[r] != r:
            r = parent[r]
        while parent[i] != r:  # path compression
            parent[i], i = r, parent[i]
        return r

    def join(a, b):
        ra, rb = find(a), find(b)
        if ra < rb:
            parent[rb] = ra
        elif rb < ra:
            parent[ra] = rb

    for y in range(H):
        for x in range(W):
            if not m[y, x]:
                continue
            i = y * W + x
            if x > 0 and m[y, x - 1]:
                join(i, i - 1)
            if y > 0:
                if x > 0 and m[y - 1, x - 1]:
                    join(i, i - W - 1)
                if m[y - 1, x]:
                    join(i, i - W)
                if x + 1 < W and m[y - 1, x + 1]:
                    join(i, i - W + 1)
    out = np.zeros(H * W, dtype=np.int64)
    nxt = 0
    flat = m.ravel()
    for i in range(H * W):
        if not flat[i]:
            continue
        r = find(i)
        if r == i:
            nxt += 1
            out[i] = nxt
        else:
            out[i] = out[r]  # r < i: already resolved
    return out.reshape(H, W)


def keep_largest_two_pass(mask: np.ndarray) -> np.ndarray:
    """PsuedoMasks.py:15-21 on top of `label_two_pass`: `regionprops` lists regions in label order and
    `max(regions, key=lambda r: r.area)` keeps the FIRST region of maximal area (Python's max); an empty
    labelling returns the input object itself (:18-19)."""
    labeled = label_two_pass(mask)
    n = int(labeled.max())
    if n == 0:
        return mask
    best, best_area = 0, -1
    for lab in range(1, n + 1):  # label order, strict > keeps the first maximum
        area = int((labeled == lab).sum())
        if area > best_area:
            best, best_area = lab, area
    return (labeled == best).astype(np.uint8)


def mask_to_png_array(mask: np.ndarray) -> np.ndarray:
    """What torchvision.utils.save_image writes for a {0,1} float mask of shape
    (1,H,W) (PsuedoMasks.py:67-69): mul(255).add_(0.5).clamp_(0,255) -> uint8,
    replicated to 3 channels.  Returns (H,W,3) uint8 in {0,255}."""
    g = np.clip(mask.astype(np.float32) * 255.0 + 0.5, 0, 255).astype(np.uint8)
    return np.repeat(g[:, :, None], 3, axis=2)


def png_mask_to_labels(png_rgb: np.ndarray, size: int = 256) -> np.ndarray:
    """SegmentationDataset.py:21,26,35 + SegmentationModel.py:100: convert('L'),
    NEAREST resize to (size,size), int64, clamp(max=1).  For a grey RGB image
    'L' is the channel value itself.  PIL NEAREST picks src = floor((dst+0.5)*in/out)."""
    from PIL import Image

    img = Image.fromarray(png_rgb).convert("L").resize((size, size), Image.NEAREST)
    lab = np.array(img).astype(np.int64)
    return np.minimum(lab, 1)


def image_minmax01(img: torch.Tensor) -> torch.Tensor:
    """(img - min) / (max - min) over the whole image -- PsuedoMasks.py:72-73."""
    return (img - img.min()) / (img.max() - img.min())


def iou_and_acc(pred_mask: torch.Tensor, true_mask: torch.Tensor):
    """ExtraUtilities.py:4-21."""
    pf = pred_mask > 0
    tf = true_mask > 0
    inter = (pf & tf).sum().item()
    union = (pf | tf).sum().item()
    correct = (pred_mask == true_mask).sum().item()
    return inter / (union + 1e-8), correct / true_mask.numel()


# --------------------------------------------------------------------------------------
# Stage 3: pairwise regularisers
# --------------------------------------------------------------------------------------


def window_offsets(window_size: int) -> List[Tuple[int, int]]:
    """dy outer, dx inner, centre skipped (AlternatingDirectionCutLoss.py:87-90)."""
    pad = window_size // 2
    return [(dy, dx) for dy in range(-pad, pad + 1) for dx in range(-pad, pad + 1) if (dy, dx) != (0, 0)]


def _shifted(padded: torch.Tensor, pad: int, dy: int, dx: int, H: int, W: int) -> torch.Tensor:
    return padded[..., pad + dy : pad + dy + H, pad + dx : pad + dx + W]


def affinities(
    image: torch.Tensor, sigma_color: float = 0.1, sigma_space: Optional[float] = 5, window_size: int = 5
) -> List[torch.Tensor]:
    """AlternatingDirectionCutLoss.py:612-637 (batched, (B,3,H,W) -> 24 x (B,1,H,W)) and
    AlternatingDirectionBoundaryLoss.py:46-70 ((3,H,W) -> 24 x (1,H,W)).
    sigma_space=None drops the spatial term (the cut loss' affinity, :95-96)."""
    single = image.dim() == 3
    img = image.unsqueeze(0) if single else image
    H, W = img.shape[-2:]
    pad = window_size // 2
    img_p = F.pad(img, (pad, pad, pad, pad), mode="reflect")
    out = []
    for dy, dx in window_offsets(window_size):
        d2 = (img - _shifted(img_p, pad, dy, dx, H, W)).pow(2).sum(dim=1, keepdim=True)
        if sigma_space is None:
            w = torch.exp(-d2 / (2 * sigma_color ** 2))
        else:
            w = torch.exp(-d2 / (2 * sigma_color ** 2) - (dx ** 2 + dy ** 2) / (2 * sigma_space ** 2))
        out.append(w.squeeze(0) if single else w)
    return out


def cut_loss(preds: torch.Tensor, images: torch.Tensor, sigma_color: float = 0.05, window_size: int = 5) -> torch.Tensor:
    """LocalNormalizedCutLoss.forward, AlternatingDirectionCutLoss.py:71-105.
    preds are LOGITS (softmax inside, :78); result divided by K*C (:105)."""
    if preds.dim() == 3:
        preds, images = preds.unsqueeze(0), images.unsqueeze(0)
    _, C, H, W = preds.shape
    pad = window_size // 2
    p = F.softmax(preds, dim=1)
    p_p = F.pad(p, (pad, pad, pad, pad), mode="reflect")
    i_p = F.pad(images, (pad, pad, pad, pad), mode="reflect")
    total, k = 0.0, 0
    for dy, dx in window_offsets(window_size):
        q = _shifted(p_p, pad, dy, dx, H, W)
        d2 = (images - _shifted(i_p, pad, dy, dx, H, W)).pow(2).sum(dim=1, keepdim=True)
        aff = torch.exp(-d2 / (2 * sigma_color ** 2))
        for c in range(C):
            total = total + (aff * (p[:, c : c + 1] - q[:, c : c + 1]) ** 2).mean()
        k += 1
    return total / (k * C)


def boundary_loss(
    probs: torch.Tensor, image: torch.Tensor, sigma_color: float = 0.1, sigma_space: float = 5, window_size: int = 5
) -> torch.Tensor:
    """ConstrainToBoundaryLossSingle.forward, AlternatingDirectionBoundaryLoss.py:20-44.
    probs (C,H,W) are used as given (no softmax); result divided by K only (:43)."""
    _, H, W = probs.shape
    pad = window_size // 2
    p_p = F.pad(probs.unsqueeze(0), (pad, pad, pad, pad), mode="reflect").squeeze(0)
    affs = affinities(image, sigma_color, sigma_space, window_size)
    total = 0.0
    for k, (dy, dx) in enumerate(window_offsets(window_size)):
        diff = (probs - _shifted(p_p, pad, dy, dx, H, W)).pow(2).sum(dim=0)
        total = total + (affs[k].squeeze(0) * diff).mean()
    return total / len(affs)


def loss_and_grad(fn, preds: torch.Tensor, *args, dtype=torch.float32, **kw):
    """Runs `fn` under autograd on CPU; returns (loss, dloss/dpreds)."""
    x = preds.detach().to(dtype).clone().requires_grad_(True)
    others = [a.detach().to(dtype) if torch.is_tensor(a) else a for a in args]
    val = fn(x, *others, **kw)
    (g,) = torch.autograd.grad(val, x)
    return val.detach(), g


# ---- closed form used to design the CUDA kernel (checked against autograd in tests) ----


def _reflect(i: int, n: int) -> int:
    if i < 0:
        return -i
    if i >= n:
        return 2 * (n - 1) - i
    return i


def _axis_multiplicity(n: int, pad: int, s1) -> np.ndarray:
    """m[a, e+pad] = sum_{d in [-pad,pad]} [reflect(a+d) == a+e] * s1(d)   (0 when a+e is outside)."""
    m = np.zeros((n, 2 * pad + 1), dtype=np.float64)
    for a in range(n):
        for d in range(-pad, pad + 1):
            b = _reflect(a + d, n)
            e = b - a
            if -pad <= e <= pad:
                m[a, e + pad] += s1(d)
            else:  # cannot happen for n > pad (reflection lands within pad of a)
                raise AssertionError("reflection left the window")
    return m


def pairwise_closed_form(
    values: np.ndarray,
    image: np.ndarray,
    sigma_color: float,
    sigma_space: Optional[float],
    window_size: int,
    inner_softmax: bool,
    divide_by_c: bool,
):
    """Gather-form loss and gradient for ONE image in fp64 numpy:

        L    = kappa * sum_z sum_y M(z,y) k(z,y) |p(z)-p(y)|^2
        dL/dp(z) = 2 kappa * sum_y [M(z,y) + M(y,z)] k(z,y) (p(z)-p(y))

    with k the colour kernel, M(z,y) = my(zy,yy)*mx(zx,yx) the (separable)
    number of window offsets d with reflect(z+d) == y, weighted by the spatial
    Gaussian.  values (C,H,W); image (3,H,W).  Returns (loss, grad (C,H,W))."""
    v = values.astype(np.float64)
    img = image.astype(np.float64)
    C, H, W = v.shape
    pad = window_size // 2
    if inner_softmax:
        e = np.exp(v - v.max(axis=0, keepdims=True))
        p = e / e.sum(axis=0, keepdims=True)
    else:
        p = v
    if sigma_space is None:
        s1 = lambda d: 1.0
    else:
        s1 = lambda d: math.exp(-(d * d) / (2.0 * sigma_space ** 2))
    my = _axis_multiplicity(H, pad, s1)
    mx = _axis_multiplicity(W, pad, s1)
    K = window_size * window_size - 1
    kappa = 1.0 / (K * H * W * (C if divide_by_c else 1))
    loss = 0.0
    g = np.zeros_like(p)
    inv = 1.0 / (2.0 * sigma_color ** 2)
    for ey in range(-pad, pad + 1):
        for ex in range(-pad, pad + 1):
            if ey == 0 and ex == 0:
                continue
            zy0, zy1 = max(0, -ey), min(H, H - ey)
            zx0, zx1 = max(0, -ex), min(W, W - ex)
            if zy0 >= zy1 or zx0 >= zx1:
                continue
            zs = (slice(zy0, zy1), slice(zx0, zx1))
            ys = (slice(zy0 + ey, zy1 + ey), slice(zx0 + ex, zx1 + ex))
            kz = np.exp(-((img[(slice(None),) + zs] - img[(slice(None),) + ys]) ** 2).sum(0) * inv)
            dp = p[(slice(None),) + zs] - p[(slice(None),) + ys]
            m_f = my[zy0:zy1, ey + pad][:, None] * mx[zx0:zx1, ex + pad][None, :]
            m_b = my[zy0 + ey : zy1 + ey, -ey + pad][:, None] * mx[zx0 + ex : zx1 + ex, -ex + pad][None, :]
            loss += (m_f * kz * (dp ** 2).sum(0)).sum()
            g[(slice(None),) + zs] += (m_f + m_b) * kz * dp
    loss *= kappa
    g *= 2.0 * kappa
    if inner_softmax:
        g = p * (g - (p * g).sum(axis=0, keepdims=True))
    return loss, g


# --------------------------------------------------------------------------------------
# "next" row: alternating-direction refinement (AlternatingDirectionCutLoss.py:709-767)
# --------------------------------------------------------------------------------------


def refine_from_probs(
    S: torch.Tensor,
    image: torch.Tensor,
    mask: torch.Tensor,
    lambda_boundary: float = 0.1,
    threshold: float = 0.5,
    lr: float = 1e-2,
    num_steps: int = 20,
    sigma_color: float = 0.1,
    window_size: int = 5,
    return_state: bool = False,
):
    """refine_pseudo_mask after the network call: S = softmax(model(image)['out']) is
    given ((1,2,H,W)); mask holds {0,255}.  Adam on X (torch defaults), KL(batchmean)
    + dynamically weighted cut loss, threshold on softmax(X)[0,1]."""
    mask = (mask == 255).long()
    X = F.one_hot(mask, num_classes=2).permute(2, 0, 1).to(S.dtype).unsqueeze(0).clone().requires_grad_(True)
    opt = torch.optim.Adam([X], lr=lr)
    for _ in range(num_steps):
        opt.zero_grad()
        Xn = F.softmax(X, dim=1)
        l_kl = F.kl_div((Xn + 1e-8).log(), S, reduction="batchmean")
        l_b = cut_loss(Xn[0], image, sigma_color=sigma_color, window_size=window_size)
        lam = lambda_boundary * (l_kl.item() / (l_b.item() + 1e-6))
        (l_kl + lam * l_b).backward()
        opt.step()
    Xf = F.softmax(X, dim=1)
    refined = (Xf[0, 1] > threshold).float()
    if return_state:
        return refined, X.detach(), Xf.detach()
    return refined

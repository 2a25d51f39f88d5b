"""Shared test helpers: synthetic inputs of SURVEY.md 8(d) and comparison rules of BASELINE.json's north_star."""
import numpy as np
import torch


def smooth_images(gen, B, H, W, device="cpu"):
    """box-filtered noise rescaled per image to [0,1] (iid noise gives ~0 affinity)."""
    raw = torch.rand(B, 3, H + 16, W + 16, generator=gen)
    img = torch.nn.functional.avg_pool2d(raw, 9, stride=1)[..., 4 : H + 4, 4 : W + 4]
    lo = img.amin(dim=(1, 2, 3), keepdim=True)
    hi = img.amax(dim=(1, 2, 3), keepdim=True)
    return ((img - lo) / (hi - lo)).contiguous().to(device)


def synth_act_grad(gen, B, C, h, w, negative_acts=True):
    """config-3 style hooks: act ~ relu(randn) (some negatives kept on request to pin relu(grad*act)),
    grad ~ randn * 1e-3."""
    act = torch.randn(B, C, h, w, generator=gen)
    if not negative_acts:
        act = act.relu()
    grad = torch.randn(B, C, h, w, generator=gen) * 1e-3
    return act, grad


CAM_RTOL = 1e-5  # north_star: CAM maps within 1e-5 relative (fp32)
CAM_ATOL = 1e-6  # CAMs live in [0,1]; 1e-6 is also the threshold band
LOSS_RTOL = 1e-5  # loss values / gradients within 1e-5 relative


def assert_cam_close(ours, ref, what="cam"):
    ours = torch.as_tensor(ours).double().cpu()
    ref = torch.as_tensor(ref).double().cpu()
    err = (ours - ref).abs()
    tol = CAM_ATOL + CAM_RTOL * ref.abs()
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} px off, max abs err {err.max().item():.3e}"


def assert_masks_match(mask, cam_ref64, thresh, near_count=None, band=1e-6, what="mask"):
    """Bit-exact outside the 1e-6 band around the threshold; pixels inside are counted (north_star)."""
    mask = torch.as_tensor(mask).cpu().to(torch.uint8)
    cam = torch.as_tensor(cam_ref64).double().cpu()
    ref = ((cam >= thresh) & (cam > 0)).to(torch.uint8)
    inside = (cam - thresh).abs() < band
    diff = (mask != ref) & ~inside
    assert not diff.any(), f"{what}: {int(diff.sum())} mismatching pixels outside the threshold band"
    return int(inside.sum())


def assert_grad_close(ours, ref, what="grad"):
    """1e-5 relative to the gradient's scale (element-wise relative error is meaningless where terms cancel)."""
    ours = torch.as_tensor(ours).double().cpu()
    ref = torch.as_tensor(ref).double().cpu()
    scale = ref.abs().max().item()
    err = (ours - ref).abs().max().item()
    assert err <= LOSS_RTOL * scale + 1e-30, f"{what}: max abs err {err:.3e} vs scale {scale:.3e} (rel {err / max(scale, 1e-300):.3e})"


def assert_loss_close(ours, ref, what="loss"):
    o = float(torch.as_tensor(ours).double().reshape(-1)[0]) if torch.as_tensor(ours).numel() == 1 else None
    if o is None:
        ours = torch.as_tensor(ours).double().cpu()
        ref = torch.as_tensor(ref).double().cpu()
        rel = ((ours - ref).abs() / ref.abs().clamp_min(1e-300)).max().item()
    else:
        r = float(torch.as_tensor(ref).double().reshape(-1)[0])
        rel = abs(o - r) / max(abs(r), 1e-300)
    assert rel <= LOSS_RTOL, f"{what}: relative error {rel:.3e}"

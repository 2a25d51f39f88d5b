"""Drop-ins for the hot-path symbols of the reference's TraditionalModel/AlternatingDirectionCutLoss.py
(the rest of that file is a notebook script: data download, training loops, DenseCRF -- out of scope)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as WF
from .LayerCAM import LayerCAMGenerator as _BaseLayerCAMGenerator


class LocalNormalizedCutLoss(nn.Module):
    """reference AlternatingDirectionCutLoss.py:65-105: preds are LOGITS (softmax inside), images (B,3,H,W);
    3-D inputs are auto-batched (:72-74); returns a 0-dim tensor differentiable w.r.t. preds."""

    def __init__(self, sigma_color=0.05, window_size=5):
        super().__init__()
        self.sigma_color = sigma_color
        self.window_size = window_size

    def forward(self, preds, images):
        if preds.dim() == 3:
            preds = preds.unsqueeze(0)
            images = images.unsqueeze(0)
        loss = WF.pairwise_loss(preds, images, window_size=self.window_size, sigma_color=self.sigma_color,
                                sigma_space=None, inner_softmax=True, divide_by_c=True, per_image=False)
        return loss.reshape(())


def compute_affinities(image, sigma_color=0.1, sigma_space=5, window_size=5):
    """reference AlternatingDirectionCutLoss.py:612-637: (B,3,H,W) -> list of K (B,1,H,W)."""
    out = WF.affinities(image, sigma_color, sigma_space, window_size)  # (K,B,1,H,W)
    return [out[k] for k in range(out.shape[0])]


class LayerCAMGenerator(_BaseLayerCAMGenerator):
    """The notebook's variant (AlternatingDirectionCutLoss.py:216-318): alpha applied per layer between two
    normalisations (:271-279), keyword order generate(images, class_idx=None, alpha=1.0)."""

    alpha_mode = 1

    def generate(self, images, class_idx=None, alpha=1.0):
        return super().generate(images, alpha=alpha, class_idx=class_idx)

    def generate_bg_cam(self, image_tensor, valid_class_indices, alpha=2.0):
        """:296-318 -- max over the returned CAMs, m_bg = 1 - clamp(1-max,0)**alpha, both resized."""
        all_cams = self.generate(image_tensor, valid_class_indices)
        max_obj_cam, _ = all_cams.max(dim=0)
        m_bg = 1.0 - ((1.0 - max_obj_cam).clamp(min=0.0) ** alpha)
        size = self.output_size
        m_bg_resized = F.interpolate(m_bg[None, None], size=size, mode='bilinear', align_corners=False).squeeze()
        max_obj_cam_resized = F.interpolate(max_obj_cam[None, None], size=size, mode='bilinear',
                                            align_corners=False).squeeze()
        return m_bg_resized, max_obj_cam_resized


def refine_pseudo_mask(model, image, mask, lambda_boundary=0.1, threshold=0.5, lr=1e-2, num_steps=20,
                       sigma_color=0.1, window_size=5):
    """reference AlternatingDirectionCutLoss.py:709-767 with the cut loss on the fused kernel.  Same optimiser
    (Adam on X), same dynamic lambda (two host reads per step, :748), same threshold on softmax(X)[0,1]."""
    device = next(model.parameters()).device
    image = image.to(device)

    model.eval()
    with torch.no_grad():
        input_tensor = image.unsqueeze(0)
        S = model(input_tensor)['out']
        S = F.softmax(S, dim=1)

    num_classes = 2
    mask = (mask == 255).long()
    X_init = F.one_hot(mask.long(), num_classes=num_classes).permute(2, 0, 1).float()
    X = X_init.unsqueeze(0).to(device).requires_grad_(True)
    optimizer_X = torch.optim.Adam([X], lr=lr)

    criterion_boundary = LocalNormalizedCutLoss(sigma_color=sigma_color, window_size=window_size)

    for step in range(num_steps):
        optimizer_X.zero_grad()
        X_norm = F.softmax(X, dim=1)
        loss_kl = F.kl_div((X_norm + 1e-8).log(), S, reduction='batchmean')
        loss_boundary = criterion_boundary(X_norm[0], input_tensor[0])
        lambda_boundary_dynamic = lambda_boundary * (loss_kl.item() / (loss_boundary.item() + 1e-6))
        loss = loss_kl + lambda_boundary_dynamic * loss_boundary
        loss.backward()
        optimizer_X.step()

    X_final = F.softmax(X, dim=1)
    pseudo_mask_refined = (X_final[0, 1] > threshold).float()
    return pseudo_mask_refined


def refine_pseudo_masks_batched(model, images, masks, lambda_boundary=0.1, threshold=0.5, lr=1e-2, num_steps=20,
                                sigma_color=0.1, window_size=5, return_state=False):
    """`refine_pseudo_mask` (reference AlternatingDirectionCutLoss.py:709-767) for a whole batch at once, with
    nothing on the host inside the step loop.

    images (B,3,H,W), masks (B,H,W) holding {0,255}.  Every image keeps its own problem exactly as in the
    reference -- its own KL term (batchmean over a batch of one), its own cut loss (the inner softmax of
    LocalNormalizedCutLoss included, :745 -> :78) and its own dynamic weight lambda * KL / (cut + 1e-6) (:748) --
    but the weight stays on the device: it is handed to the fused cut-loss launch as the per-image upstream gradient,
    so the two `.item()` syncs per step and image of the reference disappear.  Adam on X with torch's defaults,
    written out element-wise (identical for every image since Adam is element-wise).  Returns (B,H,W) float {0,1}."""
    device = next(model.parameters()).device
    images = images.to(device).float()
    model.eval()
    with torch.no_grad():
        S = F.softmax(model(images)['out'], dim=1)
    B = images.shape[0]
    m = (masks.to(device) == 255).long()
    X = F.one_hot(m, num_classes=2).permute(0, 3, 1, 2).float().contiguous()
    exp_avg = torch.zeros_like(X)
    exp_avg_sq = torch.zeros_like(X)
    beta1, beta2, eps = 0.9, 0.999, 1e-8
    lam = torch.full((B,), float(lambda_boundary), device=device)
    for step in range(1, num_steps + 1):
        Xn = F.softmax(X, dim=1)
        # KL(S || Xn), reduction 'batchmean' with a batch of one: the sum over the image
        logq = (Xn + 1e-8).log()
        kl = (torch.xlogy(S, S) - S * logq).flatten(1).sum(1)                       # (B,)
        # cut loss of every image and its gradient w.r.t. Xn: one fused launch for the whole batch
        loss_b, g_b = WF.pairwise_loss_and_grad(Xn, images, window_size, sigma_color, None, True, True, True)
        w = lam * kl / (loss_b + 1e-6)                                               # (B,), stays on the device
        # d total / d Xn = -S / (Xn + 1e-8) + w * d cut / d Xn ; then through the outer softmax
        gXn = -S / (Xn + 1e-8) + w.view(B, 1, 1, 1) * g_b
        gX = Xn * (gXn - (gXn * Xn).sum(dim=1, keepdim=True))
        # Adam (torch defaults: betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad)
        exp_avg.mul_(beta1).add_(gX, alpha=1 - beta1)
        exp_avg_sq.mul_(beta2).addcmul_(gX, gX, value=1 - beta2)
        bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
        denom = (exp_avg_sq.sqrt() / (bc2 ** 0.5)).add_(eps)
        X.addcdiv_(exp_avg, denom, value=-lr / bc1)
    Xf = F.softmax(X, dim=1)
    refined = (Xf[:, 1] > threshold).float()
    if return_state:
        return refined, X, Xf
    return refined

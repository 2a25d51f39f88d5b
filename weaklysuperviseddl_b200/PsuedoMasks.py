"""Drop-in for the reference's TraditionalModel/PsuedoMasks.py (sic): pseudo-mask thresholding, largest-component
filter and label assembly, with LayerCAM -> threshold fused on the GPU (csrc/layercam.cu) and the connected
components on the GPU as well (csrc/ccl.cu).  One device->host copy per loader batch instead of one per image."""
from __future__ import annotations

import os

import numpy as np
import torch

from . import functional as WF

CONTENT_ROOT = "/content"  # the reference hard-codes Colab's /content (PsuedoMasks.py:31-32)
MAX_IMAGES = 500           # PsuedoMasks.py:49-50


def delete_dir_recursive(path):
    """PsuedoMasks.py:5-12."""
    if not os.path.exists(path):
        return
    for root, dirs, files in os.walk(path, topdown=False):
        for name in files:
            os.remove(os.path.join(root, name))
        for name in dirs:
            os.rmdir(os.path.join(root, name))
    os.rmdir(path)


def keep_largest(mask):
    """PsuedoMasks.py:15-21.  Accepts the reference's numpy (H,W) mask (returned as numpy uint8) or a CUDA
    tensor (returned as a CUDA tensor).  An empty mask comes back unchanged."""
    if isinstance(mask, np.ndarray):
        if not torch.cuda.is_available():
            raise WF._native.WsdlError("wsdl_b200.keep_largest needs a CUDA device; there is no CPU fallback")
        t = torch.from_numpy(np.ascontiguousarray(mask)).cuda()
        out = WF.keep_largest(t).cpu().numpy()
        return out if out.any() else mask
    return WF.keep_largest(mask)


def _save_png(array_hwc_u8: np.ndarray, path: str) -> None:
    from PIL import Image

    Image.fromarray(array_hwc_u8).save(path, format="PNG")


def mask_to_png_array(mask_u8: np.ndarray) -> np.ndarray:
    """What torchvision.utils.save_image writes for the (1,H,W) float {0,1} mask of PsuedoMasks.py:67-69:
    *255 + 0.5, clamp, uint8, grey replicated to RGB."""
    g = (mask_u8.astype(np.uint8) != 0).astype(np.uint8) * 255
    return np.repeat(g[:, :, None], 3, axis=2)


def image_to_png_array(img: torch.Tensor) -> np.ndarray:
    """PsuedoMasks.py:71-74: per-image min-max to [0,1], then save_image's *255+0.5 -> uint8 (H,W,3)."""
    img = img.detach().float().cpu()
    img = (img - img.min()) / (img.max() - img.min())
    return img.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8).numpy()


def generate_pseudo_masks(
    loader,
    layercam_gen,
    cam_thresh=0.3,
    alpha=1.0,
    keep_largest_masks=True,
    run_id="default",
    content_root=CONTENT_ROOT,
    max_images=MAX_IMAGES,
    forward_alpha=False,
):
    """PsuedoMasks.py:23-79.  Returns (image_save_dir, save_dir); files are `{img_id}.png` as in the reference.
    The whole loader batch goes through the backbone at once (eval-mode BN keeps samples independent).

    `alpha` is accepted and IGNORED, exactly as in the reference: its call `layercam_gen.generate(img,
    class_idx=...)` (PsuedoMasks.py:58) never passes it, so the CAM is always made with the generator's default
    exponent 1.0.  `forward_alpha=True` is an extension that hands it to the generator."""
    save_dir = os.path.join(content_root, f"pseudo_masks_{run_id}")
    image_save_dir = os.path.join(content_root, f"images_{run_id}")
    for d in (save_dir, image_save_dir):
        delete_dir_recursive(d)
        os.makedirs(d)

    if not torch.cuda.is_available():
        raise WF._native.WsdlError("generate_pseudo_masks needs a CUDA device; there is no CPU fallback")
    device = torch.device("cuda")
    img_id = 0
    near_total = torch.zeros(1, dtype=torch.int64, device=device)

    for imgs, (labels, _) in loader:
        if img_id >= max_images:
            break
        take = min(imgs.size(0), max_images - img_id)
        imgs = imgs[:take].to(device)
        labels = torch.as_tensor(labels)[:take].to(device).long()

        masks, near = layercam_gen.generate_masks(imgs, cam_thresh=cam_thresh, alpha=alpha if forward_alpha else 1.0,
                                                  class_idx=labels)
        near_total += near
        if keep_largest_masks:
            masks = WF.keep_largest(masks)
        masks_host = masks.cpu().numpy()  # one D2H per batch

        for i in range(take):
            _save_png(mask_to_png_array(masks_host[i]), os.path.join(save_dir, f"{img_id}.png"))
            _save_png(image_to_png_array(imgs[i]), os.path.join(image_save_dir, f"{img_id}.png"))
            img_id += 1

    print(f"Pseudo masks saved to: {save_dir}")
    print(f"Images saved to: {image_save_dir}")
    generate_pseudo_masks.last_near_threshold_pixels = int(near_total.item())
    return image_save_dir, save_dir

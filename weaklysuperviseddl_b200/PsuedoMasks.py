"""Drop-in for the reference's TraditionalModel/PsuedoMasks.py (sic): pseudo-mask thresholding, largest-component
filter and label assembly, with LayerCAM -> threshold fused on the GPU (csrc/layercam.cu) and the connected
components on the GPU as well (csrc/ccl.cu).  One device->host copy per loader batch instead of one per image."""
from __future__ import annotations

import os
from contextlib import nullcontext as _nullcontext

import numpy as np
import torch

from . import functional as WF
from . import sharding

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


def generate_pseudo_masks_sharded(hook_source, n_images, out_size, cam_thresh=0.3, alpha=1.0, keep_largest_masks=False,
                                  chunk=128, rank=None, world=None, sink=None, mask_fn=None, device=None, streams=1,
                                  count_foreground=False):
    """Pseudo-mask generation over a whole image set, sharded by image across the GPUs of a box (BASELINE config 3;
    the per-image loop of PsuedoMasks.py:41-76 without its 500-image cap and without files).

    One process per GPU.  Image i belongs to rank i % world (`sharding.shard_indices`); a rank walks its images in
    resident chunks of `chunk`: `hook_source(indices)` returns the hooked activations and gradients of those images
    (lists of (len(indices), C_l, h_l, w_l) tensors -- in the reference's pipeline the classifier's forward/backward
    through LayerCAMGenerator's hooks, in the benchmark synthetic hooks seeded by image index), the fused LayerCAM ->
    normalise -> upsample -> threshold launch turns them into u8 masks (optionally reduced to their largest
    component), and `sink(indices, masks)` takes them (default: they are kept and returned).  No pixel ever crosses
    ranks: the only collective is the sum of three counters at the end.

    `streams=2` alternates consecutive chunks between two CUDA streams, so that the (instruction-bound, L2-resident)
    upsample / threshold kernel of one chunk runs under the (HBM-bound) channel sum of the next.
    `mask_fn(acts, grads) -> (mask u8 (B,H,W), near_threshold_count)` replaces the fused launch (the CPU tests hand in
    the oracle to check the sharding logic without a GPU).  Returns {"indices", "masks" (None with a sink),
    "counters": {"masks", "near_threshold_pixels", "foreground_pixels"} summed over ranks}; `count_foreground=True`
    adds the per-chunk pixel count (one extra read of every mask; the counter is 0 otherwise)."""
    if rank is None or world is None:
        rank, world = sharding.rank_world()
    mine = sharding.shard_indices(int(n_images), rank, world)
    own_kernel = mask_fn is None
    out_h, out_w = int(out_size[0]), int(out_size[1])
    lanes = max(1, int(streams)) if (own_kernel and torch.cuda.is_available()) else 1
    side = [torch.cuda.Stream() for _ in range(lanes)] if lanes > 1 else [None]
    near_total, fg_total = [None] * lanes, [None] * lanes
    # buffers are made once: the masks of the whole shard (or one chunk-sized buffer per stream when a sink takes them),
    # one workspace and one near-threshold counter per stream; the kernels write straight into them
    result, lane_mask, lane_ws, n_masks, pos = None, [None] * lanes, [None] * lanes, 0, 0
    cur = torch.cuda.current_stream() if lanes > 1 else None
    for s_ in side:
        if s_ is not None:
            s_.wait_stream(cur)
    kept = []
    for c, idx in enumerate(sharding.chunk(mine, int(chunk))):
        lane = c % lanes
        with torch.cuda.stream(side[lane]) if side[lane] is not None else _nullcontext():
            acts, grads = hook_source(list(idx))
            n = len(idx)
            if own_kernel:
                dev_ = acts[0].device
                if lane_ws[lane] is None:  # allocate on the caller's stream: the buffers outlive the side streams' work
                    with torch.cuda.stream(cur) if cur is not None else _nullcontext():
                        if sink is None and result is None:
                            result = torch.empty((len(mine), out_h, out_w), dtype=torch.uint8, device=dev_)
                        if sink is not None:
                            lane_mask[lane] = torch.empty((int(chunk), out_h, out_w), dtype=torch.uint8, device=dev_)
                        shapes = [tuple(a.shape[1:]) for a in acts]
                        lane_ws[lane] = torch.empty(WF.layercam_workspace_bytes(shapes, int(chunk), acts[0].dtype),
                                                    dtype=torch.uint8, device=dev_)
                        near_total[lane] = torch.zeros(1, dtype=torch.int64, device=dev_)
                    if side[lane] is not None:
                        side[lane].wait_stream(cur)  # the zero fill of the counter
                dst = result[pos:pos + n] if sink is None else lane_mask[lane][:n]
                _, masks, _ = WF.layercam_fused(acts, grads, (out_h, out_w), alpha=alpha, thresh=cam_thresh, want_cam=False,
                                                near_count=near_total[lane], mask_out=dst, workspace=lane_ws[lane])
                if keep_largest_masks:
                    masks = WF.keep_largest(masks, out=masks)
            else:
                masks, near = mask_fn(acts, grads)
                near_total[lane] = near.clone() if near_total[lane] is None else near_total[lane] + near
                if keep_largest_masks:
                    masks = WF.keep_largest(masks)
                if sink is None:
                    kept.append(masks)
            if count_foreground:
                fg = masks.sum(dtype=torch.int64)
                fg_total[lane] = fg if fg_total[lane] is None else fg_total[lane] + fg
            n_masks += n
            pos += n
            if sink is not None:
                sink(list(idx), masks)  # with several streams the buffer is reused `streams` chunks later
    for s_ in side:
        if s_ is not None:
            cur.wait_stream(s_)
    if not own_kernel and sink is None:
        result = torch.cat(kept) if kept else None
    seen = [t.device for t in near_total if t is not None]
    dev = device if device is not None else (seen[0] if seen else None)  # NCCL reduces device tensors, gloo host ones
    local = [n_masks, sum(int(t.sum().item()) for t in near_total if t is not None),
             sum(int(t.item()) for t in fg_total if t is not None)]  # the only host reads: after the last chunk
    total = sharding.all_reduce_counters(local, device=dev if (dev is not None and torch.device(dev).type == "cuda") else None)
    return {"indices": torch.tensor(list(mine), dtype=torch.int64),
            "masks": result if sink is None else None,
            "counters": {"masks": total[0], "near_threshold_pixels": total[1], "foreground_pixels": total[2]}}

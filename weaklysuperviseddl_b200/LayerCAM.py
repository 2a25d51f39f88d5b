"""Drop-in for the reference's TraditionalModel/LayerCAM.py: same class, same signatures, the post-backbone
arithmetic replaced by the fused sm_100a kernels (csrc/layercam.cu).  The backbone forward/backward stays on
cuDNN through PyTorch autograd + hooks, exactly as in the reference (LayerCAM.py:17-48)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import functional as WF


class LayerCAMGenerator:
    """reference LayerCAM.py:7-81.  `output_size` is the (224, 224) the reference hard-codes at :69."""

    alpha_mode = 0  # LayerCAM.py:74-76: mean of layers -> clamp(0) ** alpha

    def __init__(self, model, target_layer_names=["layer3", "layer4"], output_size=(224, 224)):
        self.model = model.eval()
        self.target_layer_names = target_layer_names
        self.output_size = tuple(output_size)

        self.activations = {}
        self.gradients = {}

        self._register_hooks()

    def _register_hooks(self):
        for name in self.target_layer_names:
            layer = getattr(self.model, name)
            layer.register_forward_hook(self._make_forward_hook(name))
            layer.register_full_backward_hook(self._make_backward_hook(name))

    def _make_forward_hook(self, name):
        def hook(module, input, output):
            self.activations[name] = output
        return hook

    def _make_backward_hook(self, name):
        def hook(module, grad_input, grad_output):
            self.gradients[name] = grad_output[0]
        return hook

    # -- backbone part, unchanged semantics (LayerCAM.py:35-48) --
    def _forward_backward(self, images, class_idx):
        self.activations.clear()
        self.gradients.clear()
        if images.dim() == 3:
            images = images.unsqueeze(0)  # the reference always works on one image (LayerCAM.py:38)
        images = images.detach().requires_grad_()
        with torch.enable_grad():
            logits, _ = self.model(images)
            if class_idx is None:
                class_idx = torch.argmax(logits, dim=1)
            class_scores = logits.gather(1, class_idx.view(-1, 1)).squeeze()
            class_scores.backward(torch.ones_like(class_scores), retain_graph=False)
        acts = [self.activations[n] for n in self.target_layer_names]
        grads = [self.gradients[n] for n in self.target_layer_names]
        return acts, grads

    def generate(self, images, alpha=1.0, class_idx=None):
        """images (3,H,W) [or a batch (B,3,H,W)] -> CAM (B, out_h, out_w) f32 in [0,1], detached.
        `alpha` defaults to 1.0 so that PsuedoMasks.py:58's call without alpha works."""
        acts, grads = self._forward_backward(images, class_idx)
        cam, _, _ = WF.layercam_fused(acts, grads, self.output_size, alpha=alpha, alpha_mode=self.alpha_mode)
        return cam  # (B, H, W)

    def generate_masks(self, images, cam_thresh=0.3, alpha=1.0, class_idx=None, return_cam=False):
        """Fused generate + threshold (PsuedoMasks.py:58-62): the full-resolution CAM never reaches HBM
        unless `return_cam`.  Returns (mask u8 (B,H,W), near_threshold_pixel_count[, cam])."""
        acts, grads = self._forward_backward(images, class_idx)
        cam, mask, near = WF.layercam_fused(acts, grads, self.output_size, alpha=alpha, alpha_mode=self.alpha_mode,
                                            thresh=cam_thresh, want_cam=return_cam)
        return (mask, near, cam) if return_cam else (mask, near)


MAX_EVAL_ITEMS = 11  # the reference stops after the item with index 10 (LayerCAM.py:119-120)


def _cam_truth_counts(layercam_gen, img, label, true_mask, alpha, cam_thresh):
    """One test item (LayerCAM.py:96-115) without a host synchronisation: fused CAM -> threshold on the device,
    nearest resize when the truth has another shape (it always has with the reference's loader: (1,H,W) against
    (H,W), :109), and the three counters of compute_iou_and_acc as a (1,3) device tensor."""
    true_fg = (true_mask.cuda() == 1).long()
    cls = torch.tensor([label.item() if isinstance(label, torch.Tensor) else label], device=true_fg.device)
    mask, _ = layercam_gen.generate_masks(img.cuda(), cam_thresh=cam_thresh, alpha=alpha, class_idx=cls)
    pred = mask.squeeze(0).long()
    if pred.shape != true_fg.shape:
        pred = F.interpolate(pred[None, None].float(), size=true_fg.shape[-2:], mode='nearest').squeeze().long()
    return WF.iou_acc_counts(pred, true_fg, batched=False), true_fg.numel()


def evaluate_layercam_on_test_set(layercam_gen, test_loader, alpha=1.0, cam_thresh=0.3):
    """reference LayerCAM.py:84-130: mean foreground IoU / pixel accuracy of the thresholded CAM over the first 11
    items of a batch-size-1 test loader yielding (img, (label, true_mask)); same printout, same return dict.  The
    per-item `.item()` reads of the reference (three per image) are one device->host copy at the end."""
    counts, totals = [], []
    for img, (label, true_mask) in test_loader:
        c, n = _cam_truth_counts(layercam_gen, img[0], label[0], true_mask[0], alpha, cam_thresh)
        counts.append(c)
        totals.append(n)
        if len(counts) >= MAX_EVAL_ITEMS:
            break
    host = torch.cat(counts).tolist()
    ious_fg = [inter / (union + 1e-8) for inter, union, _ in host]      # ExtraUtilities.py:17-20
    accs_fg = [correct / n for (_, _, correct), n in zip(host, totals)]
    result = {"layercam_fg_iou": sum(ious_fg) / len(ious_fg), "layercam_fg_acc": sum(accs_fg) / len(accs_fg)}
    print("\n Evaluation of CAMs on test set:")
    print(f" - LayerCam FG: Avg IoU: {result['layercam_fg_iou']:.4f} | Acc: {result['layercam_fg_acc']:.4f}")
    return result

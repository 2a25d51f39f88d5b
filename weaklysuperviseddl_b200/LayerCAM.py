"""Drop-in for the reference's TraditionalModel/LayerCAM.py: same class, same signatures, the post-backbone
arithmetic replaced by the fused sm_100a kernels (csrc/layercam.cu).  The backbone forward/backward stays on
cuDNN through PyTorch autograd + hooks, exactly as in the reference (LayerCAM.py:17-48)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import functional as WF
from .ExtraUtilities import compute_iou_and_acc


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


def evaluate_layercam_on_test_set(layercam_gen, test_loader, alpha=1.0, cam_thresh=0.3):
    """reference LayerCAM.py:84-130 (same loop, same 11-image cut-off, same return dict)."""
    ious_fg, accs_fg = [], []

    for i, (img, (label, true_mask)) in enumerate(test_loader):
        img = img[0].cuda()
        true_mask = true_mask[0].cuda()

        true_mask = (true_mask == 1).long()
        label = label[0].item() if isinstance(label[0], torch.Tensor) else label[0]
        class_tensor = torch.tensor([label]).to(img.device)

        mask, _ = layercam_gen.generate_masks(img, cam_thresh=cam_thresh, alpha=alpha, class_idx=class_tensor)
        pred_fg_mask = mask.squeeze(0).long()

        if pred_fg_mask.shape != true_mask.shape:
            pred_fg_mask = F.interpolate(pred_fg_mask.unsqueeze(0).unsqueeze(0).float(), size=true_mask.shape[-2:],
                                         mode='nearest').squeeze().long()

        iou_fg, acc_fg = compute_iou_and_acc(pred_fg_mask, true_mask)
        ious_fg.append(iou_fg)
        accs_fg.append(acc_fg)

        if i >= 10:
            break

    print("\n Evaluation of CAMs on test set:")
    print(f" - LayerCam FG: Avg IoU: {sum(ious_fg)/len(ious_fg):.4f} | Acc: {sum(accs_fg)/len(accs_fg):.4f}")

    return {
        "layercam_fg_iou": sum(ious_fg) / len(ious_fg),
        "layercam_fg_acc": sum(accs_fg) / len(accs_fg)
    }

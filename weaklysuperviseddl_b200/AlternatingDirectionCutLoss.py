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


def _network_probs(model, images):
    """S = softmax(model(images)['out']) with the model in eval mode (AlternatingDirectionCutLoss.py:716-720)."""
    model.eval()
    with torch.no_grad():
        return F.softmax(model(images)['out'], dim=1).float().contiguous()


def refine_pseudo_mask(model, image, mask, lambda_boundary=0.1, threshold=0.5, lr=1e-2, num_steps=20,
                       sigma_color=0.1, window_size=5):
    """reference AlternatingDirectionCutLoss.py:709-767, same signature and result: image (3,H,W), mask (H,W) holding
    {0,255} -> refined mask (H,W) float {0,1}.  A batch of one through `refine_pseudo_masks_batched`: Adam on X, KL to the
    network's softmax plus the dynamically weighted cut loss (:748), threshold on softmax(X)[0,1] -- two launches per step
    and no host synchronisation, where the reference reads two scalars back per step."""
    device = next(model.parameters()).device
    out = refine_pseudo_masks_batched(model, image.to(device).unsqueeze(0), mask.to(device).unsqueeze(0), lambda_boundary,
                                      threshold, lr, num_steps, sigma_color, window_size)
    return out[0]


_REFINERS = {}  # (device, B, H, W, steps, hyper-parameters) -> _Refiner: static buffers + the captured step loop


class _Refiner:
    """The step loop of a refinement as ONE CUDA graph over static buffers: first wsdl_refine_step (Xn, KL of the start),
    then per step the cut launch (loss and gradient of every image) and wsdl_refine_step (finish the step, begin the next)."""

    def __init__(self, device, B, H, W, num_steps, lam, threshold, lr, sigma_color, window_size, use_graph=True):
        self.args = (B, H, W, int(num_steps), float(lam), float(threshold), float(lr), float(sigma_color), int(window_size))
        f32 = dict(dtype=torch.float32, device=device)
        self.S, self.X = torch.empty(B, 2, H, W, **f32), torch.empty(B, 2, H, W, **f32)
        self.images = torch.empty(B, 3, H, W, **f32)
        self.m, self.v, self.Xn, self.g = (torch.empty(B, 2, H, W, **f32) for _ in range(4))
        self.kl, self.cut = torch.empty(2, B, **f32), torch.empty(B, **f32)
        self.mask = torch.empty(B, H, W, **f32)
        lib = WF._native.lib()
        with torch.cuda.device(device):
            self.n_ref = lib.wsdl_refine_workspace_bytes(B, H, W)
            self.n_pw = lib.wsdl_pairwise_workspace_bytes(B, H, W)
            self.ws_ref = torch.zeros(self.n_ref, dtype=torch.uint8, device=device)  # zeroed once: the kernels keep it so
            self.ws_pw = torch.empty(max(self.n_pw, 512), dtype=torch.uint8, device=device)
            WF._native.check(lib.wsdl_pairwise_workspace_init(self.ws_pw.data_ptr(), self.ws_pw.numel(),
                                                              torch.cuda.current_stream(device).cuda_stream),
                             "wsdl_pairwise_workspace_init")  # prepared once: the kernels leave it prepared
        self.graph = None
        if use_graph:
            side = torch.cuda.Stream(device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):  # one eager pass (kernel attributes, lazy loads) before the capture
                self.X.zero_(), self.S.fill_(0.5), self.images.zero_()
                self._enqueue()
            torch.cuda.current_stream(device).wait_stream(side)
            torch.cuda.synchronize(device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue()
            self.graph = g

    def _enqueue(self):
        B, H, W, steps, lam, thr, lr, sigma, window = self.args
        lib, check = WF._native.lib(), WF._native.check
        dev = self.X.device
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            self.m.zero_(), self.v.zero_()
            for t in range(steps + 1):  # t = 0: only Xn and KL of the start; t >= 1: finish step t (Adam step count t)
                last = t == steps
                if t > 0:
                    check(lib.wsdl_pairwise_fwd_bwd_prepared(
                        self.Xn.data_ptr(), self.images.data_ptr(), B, 2, H, W, window, sigma, 0.0, 1, 1, 1, None,
                        self.cut.data_ptr(), self.g.data_ptr(), self.ws_pw.data_ptr(), self.ws_pw.numel(), st),
                        "wsdl_pairwise_fwd_bwd_prepared")
                check(lib.wsdl_refine_step(
                    self.X.data_ptr(), self.S.data_ptr(), self.g.data_ptr() if t > 0 else None, self.cut.data_ptr(),
                    self.kl[(t + 1) % 2].data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.Xn.data_ptr(),
                    self.kl[t % 2].data_ptr(), self.mask.data_ptr() if last else None, B, H, W, lam, lr, 0.9, 0.999, 1e-8, t,
                    thr, self.ws_ref.data_ptr(), self.n_ref, 1, st), "wsdl_refine_step")

    def run(self, S, images, X0):
        self.S.copy_(S), self.images.copy_(images), self.X.copy_(X0)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._enqueue()
        return self.mask, self.X, self.Xn


def refine_pseudo_masks_batched(model, images, masks, lambda_boundary=0.1, threshold=0.5, lr=1e-2, num_steps=20,
                                sigma_color=0.1, window_size=5, return_state=False, use_graph=True):
    """`refine_pseudo_mask` (reference AlternatingDirectionCutLoss.py:709-767) for a whole batch at once, with
    nothing on the host inside the step loop.

    images (B,3,H,W), masks (B,H,W) holding {0,255}.  Every image keeps its own problem exactly as in the
    reference -- its own KL term (batchmean over a batch of one), its own cut loss (the inner softmax of
    LocalNormalizedCutLoss included, :745 -> :78), its own dynamic weight lambda * KL / (cut + 1e-6) (:748) and Adam
    with torch's defaults (element-wise, so identical for every image).  A step is two launches: the fused cut-loss
    forward + backward of all images (csrc/pairwise_sym.cu) and wsdl_refine_step (csrc/refine.cu: dynamic weight, KL
    gradient, softmax backward, Adam update, next softmax and next KL in one pass); the `num_steps` loop is captured in
    one CUDA graph per (shape, hyper-parameter) set and replayed.  Returns (B,H,W) float {0,1}
    (+ X and softmax(X) with return_state)."""
    device = next(model.parameters()).device
    images = images.to(device).float().contiguous()
    S = _network_probs(model, images)
    if S.shape[1] != 2:
        raise ValueError("refinement is the reference's two-class (pet / background) problem")
    B, _, H, W = S.shape
    X0 = F.one_hot((masks.to(device) == 255).long(), num_classes=2).permute(0, 3, 1, 2).float()
    key = (device.index if device.index is not None else torch.cuda.current_device(), B, H, W, int(num_steps),
           float(lambda_boundary), float(threshold), float(lr), float(sigma_color), int(window_size), bool(use_graph))
    ref = _REFINERS.get(key)
    if ref is None:
        if len(_REFINERS) >= 8:
            _REFINERS.pop(next(iter(_REFINERS)))
        ref = _REFINERS[key] = _Refiner(device, B, H, W, num_steps, lambda_boundary, threshold, lr, sigma_color, window_size,
                                        use_graph)
    mask, X, Xf = ref.run(S, images, X0)
    if return_state:
        return mask.clone(), X.clone(), Xf.clone()
    return mask.clone()

"""Drop-in for ConstrainToBoundaryLossSingle of the reference's TraditionalModel/
AlternatingDirectionBoundaryLoss.py:12-70, on the fused stencil kernel (csrc/pairwise.cu)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as WF


class ConstrainToBoundaryLossSingle(nn.Module):
    def __init__(self, sigma_color=0.1, sigma_space=5, window_size=5, eps=1e-8):
        super().__init__()
        self.sigma_color = sigma_color
        self.sigma_space = sigma_space
        self.window_size = window_size
        self.eps = eps  # unused by the reference too (BoundaryLoss.py:18)

    def forward(self, preds, image):
        """preds (C,H,W): already-softmaxed probabilities; image (3,H,W) in [0,1] -> 0-dim loss.
        Accepts a batch ((B,C,H,W), (B,3,H,W)) as an extension and then returns (B,) per-image losses."""
        batched = preds.dim() == 4
        v = preds if batched else preds.unsqueeze(0)
        im = image if batched else image.unsqueeze(0)
        loss = WF.pairwise_loss(v, im, window_size=self.window_size, sigma_color=self.sigma_color,
                                sigma_space=self.sigma_space, inner_softmax=False, divide_by_c=False, per_image=True)
        return loss if batched else loss.reshape(())

    @staticmethod
    def compute_affinities_single(image, sigma_color=0.1, sigma_space=5, window_size=5):
        """(3,H,W) -> list of K (1,H,W) affinity maps (BoundaryLoss.py:46-70; the reference forgot the
        @staticmethod but calls it as one at :29)."""
        out = WF.affinities(image.unsqueeze(0), sigma_color, sigma_space, window_size)  # (K,1,1,H,W)
        return [out[k, 0] for k in range(out.shape[0])]

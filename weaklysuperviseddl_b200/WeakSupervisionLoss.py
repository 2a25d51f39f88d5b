"""The loss of the weakly-supervised training step (BASELINE config 4): cross-entropy against the CAM pseudo-labels
plus the two pairwise regularisers, composed the way the reference's scripts weight them
(lambda_cut = 0.1: AlternatingDirectionCutLoss.py:709; lambda_boundary = 0.5: AlternatingDirectionBoundaryLoss.py:159).

The reference trains with CE (or Lovasz) only in SegmentationModel.py:96-113 and applies the pairwise terms in its
alternating-direction scripts; this module is the composition BASELINE.json's config 4 names.  Each pairwise term is
one fused forward+backward launch for the whole batch; under DistributedDataParallel nothing changes (the module has
no parameters: the only collective of the step is DDP's own gradient all-reduce over NCCL)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as WF


class WeakSupervisionLoss(nn.Module):
    def __init__(self, lambda_cut=0.1, lambda_boundary=0.5, sigma_cut=0.05, sigma_boundary=0.1, sigma_space=5,
                 window_size=5, ignore_index=-100):
        super().__init__()
        self.lambda_cut = lambda_cut
        self.lambda_boundary = lambda_boundary
        self.sigma_cut = sigma_cut
        self.sigma_boundary = sigma_boundary
        self.sigma_space = sigma_space
        self.window_size = window_size
        self.ignore_index = ignore_index

    def forward(self, logits, images, labels):
        """logits (B,C,H,W) any float dtype (bf16 under autocast), images (B,3,H,W), labels (B,H,W) int64.
        Returns (total, dict of the three detached terms).  The pairwise terms run in fp32."""
        ce = F.cross_entropy(logits.float(), labels, ignore_index=self.ignore_index)
        x = logits.float()
        img = images.float()
        cut = WF.pairwise_loss(x, img, self.window_size, self.sigma_cut, None, True, True, False).reshape(())
        probs = torch.softmax(x, dim=1)
        bnd = WF.pairwise_loss(probs, img, self.window_size, self.sigma_boundary, self.sigma_space, False, False,
                               True).mean()
        total = ce + self.lambda_cut * cut + self.lambda_boundary * bnd
        return total, {"ce": ce.detach(), "cut": cut.detach(), "boundary": bnd.detach()}

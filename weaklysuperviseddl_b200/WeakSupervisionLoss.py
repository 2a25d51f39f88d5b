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


_SIDE_STREAMS = {}  # device index -> (stream for the cut term, stream for the boundary term)


def _side_streams(device):
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = (torch.cuda.Stream(device), torch.cuda.Stream(device))
    return _SIDE_STREAMS[key]


class WeakSupervisionLoss(nn.Module):
    def __init__(self, lambda_cut=0.1, lambda_boundary=0.5, sigma_cut=0.05, sigma_boundary=0.1, sigma_space=5,
                 window_size=5, ignore_index=-100, concurrent=True):
        super().__init__()
        self.concurrent = concurrent  # the two pairwise launches are independent: run them on two side streams
        self.lambda_cut = lambda_cut
        self.lambda_boundary = lambda_boundary
        self.sigma_cut = sigma_cut
        self.sigma_boundary = sigma_boundary
        self.sigma_space = sigma_space
        self.window_size = window_size
        self.ignore_index = ignore_index

    def forward(self, logits, images, labels):
        """logits (B,C,H,W) any float dtype (bf16 under autocast), images (B,3,H,W) float or uint8 (8-bit pixels, read
        as value / 255), labels (B,H,W) int64 or uint8.  Returns (total, dict of the three detached terms).
        Arithmetic is fp32 on the values as given.

        Two classes, window 5 (the reference's binary pet / background problem): ONE launch computes the three loss
        values and d total / d logits straight from the tensors as they are -- no fp32 copies of a bf16 network
        output, no separate cross-entropy, no gradient rescaling pass (wsdl_weak_loss_fwd_bwd)."""
        if not logits.is_contiguous():  # a channels_last network hands over NHWC-strided logits
            logits = logits.contiguous()
        if WF.weak_loss_supported(logits, images, labels, self.window_size):
            total, ce, cut, bnd_b = WF.weak_supervision_loss(
                logits, images, labels, 1.0, self.lambda_cut, self.lambda_boundary, self.sigma_cut, self.sigma_boundary,
                self.sigma_space, self.window_size, self.ignore_index)
            return total, {"ce": ce, "cut": cut, "boundary": bnd_b.mean()}
        if images.dtype == torch.uint8:
            images = images.float() / 255.0
        labels = labels.long()
        ce = F.cross_entropy(logits.float(), labels, ignore_index=self.ignore_index)
        x = logits.float()
        img = images.float()

        def cut_term():
            return WF.pairwise_loss(x, img, self.window_size, self.sigma_cut, None, True, True, False).reshape(())

        def boundary_term():
            probs = torch.softmax(x, dim=1)
            return WF.pairwise_loss(probs, img, self.window_size, self.sigma_boundary, self.sigma_space, False, False,
                                    True).mean()

        if x.is_cuda and x.shape[1] == 2 and self.window_size == 5 and min(x.shape[2:]) >= 6:
            # two classes: both regularisers see the same probability map -> ONE fused launch (wsdl_pairwise_dual_fwd_bwd)
            pair, cut, bnd_b = WF.pairwise_dual_weighted(x, img, self.lambda_cut, self.lambda_boundary, self.sigma_cut,
                                                         self.sigma_boundary, self.sigma_space, self.window_size)
            total = ce + pair
            return total, {"ce": ce.detach(), "cut": cut.detach(), "boundary": bnd_b.mean().detach()}
        if self.concurrent and x.is_cuda:
            # CTAs of one kernel fill the SM slots the other leaves idle (ramp-up, tile loads, drain): DESIGN.md 4.2
            cur = torch.cuda.current_stream(x.device)
            s_cut, s_bnd = _side_streams(x.device)
            s_cut.wait_stream(cur)
            s_bnd.wait_stream(cur)
            with torch.cuda.stream(s_cut):
                cut = cut_term()
            with torch.cuda.stream(s_bnd):
                bnd = boundary_term()
            cur.wait_stream(s_cut)
            cur.wait_stream(s_bnd)
            for t in (cut, bnd):
                t.record_stream(cur)
            x.record_stream(s_cut), x.record_stream(s_bnd), img.record_stream(s_cut), img.record_stream(s_bnd)
        else:
            cut, bnd = cut_term(), boundary_term()
        total = ce + self.lambda_cut * cut + self.lambda_boundary * bnd
        return total, {"ce": ce.detach(), "cut": cut.detach(), "boundary": bnd.detach()}

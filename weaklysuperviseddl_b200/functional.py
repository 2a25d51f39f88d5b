"""Tensor-level entry points over the C ABI (include/wsdl_b200.h).  torch is used for device memory and
streams only; every arithmetic step runs in libwsdl_b200.so.  CUDA tensors are required -- there is no
CPU path (the CPU restatement under oracle/ is test infrastructure and is never imported here)."""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch

from . import _native

_DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
_U8, _I64 = 3, 4  # WSDL_U8 / WSDL_I64 (include/wsdl_b200.h)
NEAR_BAND = 1e-6  # north_star: pixels within 1e-6 of the threshold are counted and reported


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _native.WsdlError(f"wsdl_b200: `{name}` must be a CUDA tensor; this hot path has no CPU fallback")


def _dense(t: torch.Tensor, align: int = 16) -> torch.Tensor:
    t = t.detach()
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % align:
        t = t.clone(memory_format=torch.contiguous_format)
    return t


def layercam_fused(
    acts: Sequence[torch.Tensor],
    grads: Sequence[torch.Tensor],
    out_size: Tuple[int, int] = (224, 224),
    alpha: float = 1.0,
    alpha_mode: int = 0,
    thresh: Optional[float] = None,
    want_cam: bool = True,
    want_mask: Optional[bool] = None,
    near_band: float = NEAR_BAND,
    near_count: Optional[torch.Tensor] = None,
    mask_out: Optional[torch.Tensor] = None,
    workspace: Optional[torch.Tensor] = None,
):
    """LayerCAM.py:52-76 (+ PsuedoMasks.py:59-62 when `thresh` is given) in two kernel launches.

    acts/grads: per target layer, (B, C_l, h_l, w_l) CUDA tensors of one dtype (f32 / bf16 / f16).
    `mask_out` (contiguous (B,H,W) u8) and `workspace` (u8, at least layercam_workspace_bytes) let a caller that walks
    many chunks reuse its buffers instead of allocating per call; `near_count` (int64, 1 element) is ADDED to.
    Returns (cam (B,H,W) f32 or None, mask (B,H,W) u8 or None, near_count u64 tensor or None)."""
    if len(acts) != len(grads) or len(acts) == 0:
        raise ValueError("need one activation and one gradient per target layer")
    if want_mask is None:
        want_mask = thresh is not None
    if want_mask and thresh is None:
        raise ValueError("want_mask needs a threshold")
    dev = acts[0].device
    dtype = acts[0].dtype
    if dtype not in _DTYPE_CODE:
        raise TypeError(f"unsupported dtype {dtype}")
    A: List[torch.Tensor] = []
    G: List[torch.Tensor] = []
    B = acts[0].shape[0]
    for a, g in zip(acts, grads):
        _require_cuda(a, "activation")
        _require_cuda(g, "gradient")
        if a.dim() != 4 or a.shape != g.shape or a.shape[0] != B:
            raise ValueError(f"activation/gradient shapes must be (B,C,h,w) and equal, got {tuple(a.shape)} / {tuple(g.shape)}")
        if a.dtype != dtype or g.dtype != dtype or a.device != dev or g.device != dev:
            raise TypeError("all activations/gradients must share one dtype and device")
        A.append(_dense(a))
        G.append(_dense(g))
    n = len(A)
    lib = _native.lib()
    IntArr = ctypes.c_int * n
    PtrArr = ctypes.c_void_p * n
    Cs = IntArr(*[t.shape[1] for t in A])
    hs = IntArr(*[t.shape[2] for t in A])
    ws = IntArr(*[t.shape[3] for t in A])
    code = _DTYPE_CODE[dtype]
    out_h, out_w = int(out_size[0]), int(out_size[1])
    with torch.cuda.device(dev):
        nbytes = lib.wsdl_layercam_workspace_bytes(Cs, hs, ws, n, B, code)
        if nbytes == 0:
            raise _native.WsdlError("wsdl_layercam_workspace_bytes rejected the shapes")
        if workspace is None:
            workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        elif workspace.dtype != torch.uint8 or workspace.numel() < nbytes or workspace.device != dev or not workspace.is_contiguous():
            raise ValueError(f"workspace must be a contiguous uint8 CUDA tensor of at least {nbytes} bytes")
        cam = torch.empty((B, out_h, out_w), dtype=torch.float32, device=dev) if want_cam else None
        if want_mask and mask_out is not None:
            if (mask_out.dtype != torch.uint8 or tuple(mask_out.shape) != (B, out_h, out_w) or not mask_out.is_contiguous()
                    or mask_out.device != dev):
                raise ValueError(f"mask_out must be a contiguous uint8 CUDA tensor of shape {(B, out_h, out_w)}")
            mask = mask_out
        else:
            mask = torch.empty((B, out_h, out_w), dtype=torch.uint8, device=dev) if want_mask else None
        if near_count is None and thresh is not None:
            near_count = torch.zeros(1, dtype=torch.int64, device=dev)
        rc = lib.wsdl_layercam_fused(
            PtrArr(*[t.data_ptr() for t in A]),
            PtrArr(*[t.data_ptr() for t in G]),
            Cs, hs, ws, n, B, code, out_h, out_w,
            float(alpha), int(alpha_mode),
            float(thresh) if thresh is not None else 0.0,
            float(near_band),
            cam.data_ptr() if cam is not None else None,
            mask.data_ptr() if mask is not None else None,
            near_count.data_ptr() if (near_count is not None and thresh is not None) else None,
            workspace.data_ptr(), workspace.numel(), _stream_ptr(dev),
        )
    _native.check(rc, "wsdl_layercam_fused")
    return cam, mask, near_count


def threshold_mask(cam: torch.Tensor, thresh: float, near_band: float = NEAR_BAND):
    """`cam[cam < t] = 0; mask = cam > 0` (PsuedoMasks.py:59-62) without modifying `cam`.
    Returns (mask u8 like cam, near_count)."""
    _require_cuda(cam, "cam")
    c = _dense(cam.float(), 4)
    mask = torch.empty(c.shape, dtype=torch.uint8, device=c.device)
    near = torch.zeros(1, dtype=torch.int64, device=c.device)
    with torch.cuda.device(c.device):
        rc = _native.lib().wsdl_threshold_mask(c.data_ptr(), c.numel(), float(thresh), float(near_band), mask.data_ptr(),
                                               near.data_ptr(), _stream_ptr(c.device))
    _native.check(rc, "wsdl_threshold_mask")
    return mask, near


def _upstream(t: Optional[torch.Tensor], n: int, dev, name: str) -> Optional[int]:
    """Device pointer of an upstream-gradient vector handed to the C ABI (n floats), validated first."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.device != dev:
        raise _native.WsdlError(f"wsdl_b200: `{name}` must be a CUDA tensor on {dev}")
    if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != n:
        raise ValueError(f"`{name}` must be a contiguous float32 tensor of {n} element(s), got {t.dtype} {tuple(t.shape)}")
    return t.data_ptr()


_PAIR_WS = {}  # (device index, stream, bytes, shape) -> prepared workspace (wsdl_pairwise_workspace_init, then self-cleaning)


def _pairwise_workspace(lib, dev, nbytes: int, shape=None) -> torch.Tensor:
    """One prepared workspace per device, stream, size and problem shape: calls on one stream are ordered, so they can
    share it; it is initialised once and every kernel leaves it initialised (no per-call memset, no per-call
    allocation).  The shape is part of the key because the regions inside a workspace are laid out per (B, H, W)."""
    stream = _stream_ptr(dev)
    if torch.cuda.is_current_stream_capturing():
        # memory allocated while a CUDA graph is being captured belongs to the graph's pool: do not cache it
        ws = torch.empty(max(nbytes, 512), dtype=torch.uint8, device=dev)
        _native.check(lib.wsdl_pairwise_workspace_init(ws.data_ptr(), ws.numel(), stream), "wsdl_pairwise_workspace_init")
        return ws
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), stream, nbytes, shape)
    ws = _PAIR_WS.get(key)
    if ws is None:
        if len(_PAIR_WS) > 64:
            _PAIR_WS.clear()
        ws = torch.empty(max(nbytes, 512), dtype=torch.uint8, device=dev)
        _native.check(lib.wsdl_pairwise_workspace_init(ws.data_ptr(), ws.numel(), stream), "wsdl_pairwise_workspace_init")
        _PAIR_WS[key] = ws
    return ws


def _pairwise_raw(values, images, window, sigma_color, sigma_space, inner_softmax, divide_by_c, per_image, want_grad,
                  grad_out=None):
    B, C, H, W = values.shape
    dev = values.device
    lib = _native.lib()
    with torch.cuda.device(dev):
        nbytes = lib.wsdl_pairwise_workspace_bytes(B, H, W)
        workspace = _pairwise_workspace(lib, dev, nbytes, ("single", B, H, W))
        loss = torch.empty(B if per_image else 1, dtype=torch.float32, device=dev)
        grad = torch.empty_like(values) if want_grad else None
        rc = lib.wsdl_pairwise_fwd_bwd_prepared(
            values.data_ptr(), images.data_ptr(), B, C, H, W, int(window), float(sigma_color),
            float(sigma_space) if sigma_space is not None else 0.0,
            int(inner_softmax), int(divide_by_c), int(per_image),
            _upstream(grad_out, B if per_image else 1, dev, "grad_out"),
            loss.data_ptr(), grad.data_ptr() if grad is not None else None,
            workspace.data_ptr(), nbytes, _stream_ptr(dev),
        )
    _native.check(rc, "wsdl_pairwise_fwd_bwd_prepared")
    return loss, grad


class _PairwiseLoss(torch.autograd.Function):
    """One fused launch computes the loss and d loss / d values (for an upstream gradient of 1);
    backward only rescales the saved gradient."""

    @staticmethod
    def forward(ctx, values, images, window, sigma_color, sigma_space, inner_softmax, divide_by_c, per_image):
        need = ctx.needs_input_grad[0]
        loss, grad = _pairwise_raw(values, images, window, sigma_color, sigma_space, inner_softmax, divide_by_c,
                                   per_image, need)
        if need:
            ctx.save_for_backward(grad)
        ctx.per_image = per_image
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("wsdl_b200: gradient w.r.t. the image is not on the reference's path")
        (grad,) = ctx.saved_tensors
        go = grad_loss.detach().to(torch.float32).contiguous()
        out = torch.empty_like(grad)
        per = grad[0].numel() if ctx.per_image else 0
        with torch.cuda.device(grad.device):
            rc = _native.lib().wsdl_scale(grad.data_ptr(), out.data_ptr(), grad.numel(), go.data_ptr(), per,
                                          _stream_ptr(grad.device))
        _native.check(rc, "wsdl_scale")
        return out, None, None, None, None, None, None, None


def _prep_pair(values: torch.Tensor, images: torch.Tensor):
    _require_cuda(values, "preds")
    _require_cuda(images, "images")
    if values.dim() != 4 or images.dim() != 4 or images.shape[1] != 3:
        raise ValueError(f"expected preds (B,C,H,W) and images (B,3,H,W), got {tuple(values.shape)} / {tuple(images.shape)}")
    if values.shape[0] != images.shape[0] or values.shape[2:] != images.shape[2:]:
        raise ValueError("preds and images must agree in batch and spatial size")
    v = values if (values.dtype == torch.float32 and values.is_contiguous()) else values.float().contiguous()
    im = images.detach()
    im = im if (im.dtype == torch.float32 and im.is_contiguous()) else im.float().contiguous()
    return v, im


def pairwise_loss(values, images, window_size=5, sigma_color=0.05, sigma_space=None, inner_softmax=True,
                  divide_by_c=True, per_image=False):
    """Differentiable (w.r.t. `values`) pairwise regulariser.  Returns a (1,) tensor, or (B,) when per_image."""
    v, im = _prep_pair(values, images)
    return _PairwiseLoss.apply(v, im, int(window_size), float(sigma_color),
                               float(sigma_space) if sigma_space else 0.0, bool(inner_softmax), bool(divide_by_c),
                               bool(per_image))


def pairwise_loss_and_grad(values, images, window_size=5, sigma_color=0.05, sigma_space=None, inner_softmax=True,
                           divide_by_c=True, per_image=False, grad_out: Optional[torch.Tensor] = None):
    """The fused launch itself, outside autograd: returns (loss, d(sum loss*grad_out)/d values)."""
    v, im = _prep_pair(values.detach(), images)
    return _pairwise_raw(v, im, int(window_size), float(sigma_color), float(sigma_space) if sigma_space else 0.0,
                         bool(inner_softmax), bool(divide_by_c), bool(per_image), True, grad_out)


def affinities(images: torch.Tensor, sigma_color=0.1, sigma_space=5, window_size=5) -> torch.Tensor:
    """(B,3,H,W) -> (K,B,1,H,W); entry k is the reference's k-th list element."""
    _require_cuda(images, "image")
    im = _dense(images.float(), 4)
    B, _, H, W = im.shape
    K = window_size * window_size - 1
    out = torch.empty((K, B, 1, H, W), dtype=torch.float32, device=im.device)
    with torch.cuda.device(im.device):
        rc = _native.lib().wsdl_affinities(im.data_ptr(), B, H, W, int(window_size), float(sigma_color),
                                           float(sigma_space) if sigma_space else 0.0, out.data_ptr(),
                                           _stream_ptr(im.device))
    _native.check(rc, "wsdl_affinities")
    return out


def layercam_workspace_bytes(layer_shapes: Sequence[Tuple[int, int, int]], B: int, dtype=torch.float32) -> int:
    """Workspace of layercam_fused for per-layer (C, h, w) hooks of a batch of B."""
    n = len(layer_shapes)
    IntArr = ctypes.c_int * n
    return int(_native.lib().wsdl_layercam_workspace_bytes(IntArr(*[s[0] for s in layer_shapes]), IntArr(*[s[1] for s in layer_shapes]),
                                                           IntArr(*[s[2] for s in layer_shapes]), n, int(B), _DTYPE_CODE[dtype]))


def keep_largest(mask: torch.Tensor, return_area: bool = False, out: Optional[torch.Tensor] = None):
    """PsuedoMasks.py:15-21 on the GPU: (B,H,W) or (H,W) u8/bool CUDA mask -> same shape u8 {0,1} holding only
    the largest 8-connected component (ties: first in raster order).  `out` may be the (u8, contiguous) input itself."""
    _require_cuda(mask, "mask")
    single = mask.dim() == 2
    m = mask.unsqueeze(0) if single else mask
    if m.dim() != 3:
        raise ValueError("mask must be (H,W) or (B,H,W)")
    m = (m != 0).to(torch.uint8) if m.dtype != torch.uint8 else m
    m = m.contiguous()
    B, H, W = m.shape
    dev = m.device
    lib = _native.lib()
    if out is None:
        out = torch.empty_like(m)
    else:
        out = out.unsqueeze(0) if single else out
        if out.dtype != torch.uint8 or out.shape != m.shape or not out.is_contiguous() or out.device != dev:
            raise ValueError("out must be a contiguous uint8 CUDA tensor of the mask's shape")
    area = torch.empty(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nbytes = lib.wsdl_keep_largest_workspace_bytes(B, H, W)
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        rc = lib.wsdl_keep_largest(m.data_ptr(), B, H, W, out.data_ptr(), area.data_ptr(), workspace.data_ptr(), nbytes,
                                   _stream_ptr(dev))
    _native.check(rc, "wsdl_keep_largest")
    out = out[0] if single else out
    return (out, area) if return_area else out


_ELEM_CODE = {torch.uint8: 0, torch.bool: 0, torch.int32: 1, torch.int64: 2, torch.float32: 3}


def iou_acc_counts(pred: torch.Tensor, true: torch.Tensor, batched: Optional[bool] = None) -> torch.Tensor:
    """ExtraUtilities.py:4-21 as one pass: masks -> (B,3) int64 device tensor {intersection, union, equal} per image.
    No host synchronisation.  Shapes broadcast exactly as the reference's `&`, `|` and `==` do (its own loader hands
    a (1,H,W) truth to an (H,W) prediction, LayerCAM.py:96-112); `batched` says whether dim 0 of the broadcast shape
    indexes images (default: only when both inputs are at least 3-D and agree in dim 0)."""
    _require_cuda(pred, "pred_mask")
    _require_cuda(true, "true_mask")
    if batched is None:
        batched = pred.dim() >= 3 and true.dim() >= 3 and pred.shape[0] == true.shape[0]
    if pred.shape != true.shape:
        pred, true = torch.broadcast_tensors(pred, true)  # raises like the reference when the shapes cannot broadcast
    if pred.dtype != true.dtype:
        common = torch.promote_types(pred.dtype, true.dtype)
        pred, true = pred.to(common), true.to(common)
    if pred.dtype not in _ELEM_CODE:
        pred, true = pred.float(), true.float()
    p, t = pred.contiguous(), true.contiguous()
    B = p.shape[0] if (batched and p.dim() >= 2) else 1
    n = p.numel() // B
    counts = torch.zeros((B, 3), dtype=torch.int64, device=p.device)
    if p.numel() == 0:
        return counts
    with torch.cuda.device(p.device):
        rc = _native.lib().wsdl_iou_acc_counts(p.data_ptr(), t.data_ptr(), B, n, _ELEM_CODE[p.dtype], counts.data_ptr(),
                                               _stream_ptr(p.device))
    _native.check(rc, "wsdl_iou_acc_counts")
    return counts


def labels_from_masks(masks: torch.Tensor, size: int = 256) -> torch.Tensor:
    """Pseudo-masks (B,H,W) u8 {0,1} -> training labels (B,size,size) int64 {0,1}: the reference's PNG round trip
    (PsuedoMasks.py:67-69 -> SegmentationDataset.py:21,26,35 -> SegmentationModel.py:100) without the files."""
    _require_cuda(masks, "masks")
    m = masks.unsqueeze(0) if masks.dim() == 2 else masks
    if m.dim() != 3:
        raise ValueError("masks must be (H,W) or (B,H,W)")
    m = (m != 0).to(torch.uint8) if m.dtype != torch.uint8 else m
    m = m.contiguous()
    B, H, W = m.shape
    out = torch.empty((B, size, size), dtype=torch.int64, device=m.device)
    with torch.cuda.device(m.device):
        rc = _native.lib().wsdl_labels_from_masks(m.data_ptr(), B, H, W, int(size), out.data_ptr(), _stream_ptr(m.device))
    _native.check(rc, "wsdl_labels_from_masks")
    return out[0] if masks.dim() == 2 else out


def pairwise_dual_loss_and_grad(logits, images, sigma_cut=0.05, sigma_boundary=0.1, sigma_space=5.0, window_size=5,
                                grad_out_cut: Optional[torch.Tensor] = None, grad_out_bnd: Optional[torch.Tensor] = None,
                                want_grad: bool = True):
    """One fused launch for LocalNormalizedCutLoss(logits, images) and ConstrainToBoundaryLossSingle(softmax(logits)[b],
    images[b]) for every b (two classes, window 5).  Returns (loss_cut (1,), loss_bnd (B,), d(go_cut*cut +
    sum_b go_bnd[b]*bnd[b])/d logits or None)."""
    v, im = _prep_pair(logits.detach(), images)
    B, C, H, W = v.shape
    if C != 2:
        raise ValueError("the fused cut + boundary launch is for two classes")
    dev = v.device
    lib = _native.lib()
    with torch.cuda.device(dev):
        nbytes = lib.wsdl_pairwise_dual_workspace_bytes(B, H, W)
        workspace = _pairwise_workspace(lib, dev, nbytes, ("dual", B, H, W))
        loss_cut = torch.empty(1, dtype=torch.float32, device=dev)
        loss_bnd = torch.empty(B, dtype=torch.float32, device=dev)
        grad = torch.empty_like(v) if want_grad else None
        rc = lib.wsdl_pairwise_dual_fwd_bwd(
            v.data_ptr(), im.data_ptr(), B, H, W, int(window_size), float(sigma_cut), float(sigma_boundary),
            float(sigma_space) if sigma_space else 0.0,
            _upstream(grad_out_cut, 1, dev, "grad_out_cut"), _upstream(grad_out_bnd, B, dev, "grad_out_bnd"),
            loss_cut.data_ptr(), loss_bnd.data_ptr(), grad.data_ptr() if grad is not None else None,
            workspace.data_ptr(), nbytes, 1, _stream_ptr(dev))
    _native.check(rc, "wsdl_pairwise_dual_fwd_bwd")
    return loss_cut, loss_bnd, grad


class _PairwiseDual(torch.autograd.Function):
    """lam_cut * cut(logits) + lam_bnd * mean_b boundary(softmax(logits)[b]) from one fused launch; the weights go
    into the launch as upstream gradients, backward only scales the saved gradient."""

    @staticmethod
    def forward(ctx, logits, images, sigma_cut, sigma_bnd, sigma_space, window, lam_cut, lam_bnd):
        B = logits.shape[0]
        go_c = torch.full((1,), float(lam_cut), dtype=torch.float32, device=logits.device)
        go_b = torch.full((B,), float(lam_bnd) / B, dtype=torch.float32, device=logits.device)
        need = ctx.needs_input_grad[0]
        lc, lb, g = pairwise_dual_loss_and_grad(logits, images, sigma_cut, sigma_bnd, sigma_space, window, go_c, go_b,
                                                want_grad=need)
        ctx.save_for_backward(g if need else torch.empty(0, device=logits.device))
        ctx.in_dtype = logits.dtype
        total = (float(lam_cut) * lc + float(lam_bnd) * lb.mean()).reshape(())
        lc, lb = lc.reshape(()), lb
        ctx.mark_non_differentiable(lc, lb)
        return total, lc, lb

    @staticmethod
    def backward(ctx, g_total, _g_lc, _g_lb):
        (g,) = ctx.saved_tensors
        if g.numel() == 0:
            return (None,) * 8
        return (g * g_total).to(ctx.in_dtype), None, None, None, None, None, None, None


def pairwise_dual_weighted(logits, images, lam_cut, lam_bnd, sigma_cut=0.05, sigma_boundary=0.1, sigma_space=5.0,
                           window_size=5):
    """Differentiable lam_cut * cut + lam_bnd * mean(boundary) for a two-class batch (one launch).
    Returns (weighted sum, cut loss, per-image boundary losses); the last two are detached."""
    v, im = _prep_pair(logits, images)
    return _PairwiseDual.apply(v, im, float(sigma_cut), float(sigma_boundary), float(sigma_space) if sigma_space else 0.0,
                               int(window_size), float(lam_cut), float(lam_bnd))


# ------------------------------------------------------------------------------------------------ one-launch loss
def _tma_ok(t: torch.Tensor) -> bool:
    """Row stride and base 16-byte aligned: what the tile-streaming kernel's TMA loads need."""
    return t.is_contiguous() and t.data_ptr() % 16 == 0 and (t.shape[-1] * t.element_size()) % 16 == 0


def weak_loss_supported(logits: torch.Tensor, images: torch.Tensor, labels: Optional[torch.Tensor] = None,
                        window_size: int = 5) -> bool:
    """Whether `weak_loss_and_grad` can take these tensors as they are (two classes, window 5, H, W >= 6, CUDA,
    f32 / bf16 logits, f32 / u8 images, u8 / int64 labels, TMA-addressable rows)."""
    if not (logits.is_cuda and images.is_cuda and logits.dim() == 4 and images.dim() == 4):
        return False
    B, C, H, W = logits.shape
    if C != 2 or window_size != 5 or H < 6 or W < 6 or images.shape != (B, 3, H, W):
        return False
    if logits.dtype not in (torch.float32, torch.bfloat16) or images.dtype not in (torch.float32, torch.uint8):
        return False
    if labels is not None and (labels.dtype not in (torch.uint8, torch.int64) or labels.shape != (B, H, W) or W % 4):
        return False
    return _tma_ok(logits) and _tma_ok(images) and W % 4 == 0


def weak_loss_and_grad(logits, images, labels=None, lam_ce=1.0, grad_out_cut=None, grad_out_bnd=None, sigma_cut=0.05,
                       sigma_boundary=0.1, sigma_space=5.0, window_size=5, ignore_index=-100, want_grad=True,
                       ce_inv_count: Optional[torch.Tensor] = None, packed_out: Optional[torch.Tensor] = None):
    """ONE launch (wsdl_weak_loss_fwd_bwd) for the loss of a weakly-supervised training step on a two-class batch:
    total = lam_ce * CE(logits, labels) + go_cut * cut(logits, images) + sum_b go_bnd[b] * boundary(softmax(logits)[b],
    images[b]).  logits (B,2,H,W) f32 | bf16, images (B,3,H,W) f32 | u8 (8-bit pixels, read as value / 255), labels
    (B,H,W) u8 | int64 or None.  Returns (total (1,), ce (1,) or None, cut (1,), bnd (B,), d total / d logits in the
    logits' dtype or None); the three loss values are unweighted.  The four loss outputs are views of one (3 + B,) float
    tensor [total, ce, cut, bnd...], which the caller may provide (`packed_out`) to read everything back in one copy."""
    _require_cuda(logits, "logits")
    _require_cuda(images, "images")
    logits, images = logits.detach(), images.detach()
    if not weak_loss_supported(logits, images, labels, window_size):
        raise _native.WsdlError("wsdl_weak_loss_fwd_bwd does not take these tensors (see weak_loss_supported); "
                                "compose pairwise_loss calls instead")
    B, _, H, W = logits.shape
    dev = logits.device
    lib = _native.lib()
    with torch.cuda.device(dev):
        nbytes = lib.wsdl_weak_loss_workspace_bytes(B, H, W)
        workspace = _pairwise_workspace(lib, dev, nbytes, ("weak", B, H, W))
        if packed_out is None:
            out = torch.empty(3 + B, dtype=torch.float32, device=dev)  # total, ce, cut, bnd[B]
        else:
            out = packed_out
            if out.dtype != torch.float32 or out.numel() != 3 + B or not out.is_contiguous() or out.device != dev:
                raise ValueError(f"packed_out must be a contiguous float32 CUDA tensor of {3 + B} elements")
        grad = torch.empty_like(logits) if want_grad else None
        lab_ptr, lab_code, inv_ptr = None, 0, None
        if labels is not None:
            _require_cuda(labels, "labels")
            labels = labels.contiguous()
            lab_ptr, lab_code = labels.data_ptr(), (_U8 if labels.dtype == torch.uint8 else _I64)
            can_ignore = labels.dtype != torch.uint8 or 0 <= int(ignore_index) <= 255
            if ce_inv_count is None and can_ignore:  # F.cross_entropy's mean is over the labels that are not ignored
                scratch = torch.empty(3, dtype=torch.int64, device=dev)
                ce_inv_count = scratch[2:].view(torch.float32)[:1]
                _native.check(lib.wsdl_count_valid_labels(lab_ptr, lab_code, labels.numel(), int(ignore_index),
                                                          scratch.data_ptr(), ce_inv_count.data_ptr(), _stream_ptr(dev)),
                              "wsdl_count_valid_labels")
            inv_ptr = _upstream(ce_inv_count, 1, dev, "ce_inv_count") if ce_inv_count is not None else None
        rc = lib.wsdl_weak_loss_fwd_bwd(
            logits.data_ptr(), _DTYPE_CODE[logits.dtype], images.data_ptr(), _U8 if images.dtype == torch.uint8 else 0,
            lab_ptr, lab_code, int(ignore_index), B, H, W, int(window_size), float(sigma_cut), float(sigma_boundary),
            float(sigma_space) if sigma_space else 0.0, float(lam_ce), inv_ptr,
            _upstream(grad_out_cut, 1, dev, "grad_out_cut"), _upstream(grad_out_bnd, B, dev, "grad_out_bnd"),
            out[1:].data_ptr() if labels is not None else None, out[2:].data_ptr(), out[3:].data_ptr(), out.data_ptr(),
            grad.data_ptr() if grad is not None else None, _DTYPE_CODE[logits.dtype], workspace.data_ptr(), nbytes, 1,
            _stream_ptr(dev))
    _native.check(rc, "wsdl_weak_loss_fwd_bwd")
    return out[0:1], (out[1:2] if labels is not None else None), out[2:3], out[3:], grad


_WEIGHTS = {}  # (device index, lam_cut, lam_bnd, B) -> (go_cut (1,), go_bnd (B,)): constant device vectors, made once


def _loss_weights(dev, lam_cut: float, lam_bnd: float, B: int):
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), float(lam_cut), float(lam_bnd), B)
    w = _WEIGHTS.get(key)
    if w is None or torch.cuda.is_current_stream_capturing():
        if len(_WEIGHTS) > 64:
            _WEIGHTS.clear()
        w = (torch.full((1,), float(lam_cut), dtype=torch.float32, device=dev),
             torch.full((B,), float(lam_bnd) / B, dtype=torch.float32, device=dev))
        if not torch.cuda.is_current_stream_capturing():
            _WEIGHTS[key] = w
    return w


class _WeakLoss(torch.autograd.Function):
    """total = lam_ce CE + lam_cut cut + lam_bnd mean_b bnd from one launch; backward scales the saved gradient."""

    @staticmethod
    def forward(ctx, logits, images, labels, lam_ce, lam_cut, lam_bnd, sigma_cut, sigma_bnd, sigma_space, window,
                ignore_index):
        go_c, go_b = _loss_weights(logits.device, lam_cut, lam_bnd, logits.shape[0])
        need = ctx.needs_input_grad[0]
        total, ce, cut, bnd, g = weak_loss_and_grad(logits, images, labels, lam_ce, go_c, go_b, sigma_cut, sigma_bnd,
                                                    sigma_space, window, ignore_index, want_grad=need)
        ctx.save_for_backward(g if need else torch.empty(0, device=logits.device))
        total = total.reshape(())
        outs = (total, ce.reshape(()) if ce is not None else total.new_zeros(()), cut.reshape(()), bnd)
        ctx.mark_non_differentiable(*outs[1:])
        return outs

    @staticmethod
    def backward(ctx, g_total, _g_ce, _g_cut, _g_bnd):
        (g,) = ctx.saved_tensors
        if g.numel() == 0:
            return (None,) * 11
        return (g * g_total.to(g.dtype),) + (None,) * 10


def weak_supervision_loss(logits, images, labels=None, lam_ce=1.0, lam_cut=0.1, lam_bnd=0.5, sigma_cut=0.05,
                          sigma_boundary=0.1, sigma_space=5.0, window_size=5, ignore_index=-100):
    """Differentiable (w.r.t. logits) lam_ce * CE + lam_cut * cut + lam_bnd * mean_b boundary, one launch.
    Returns (total, ce, cut, per-image boundary losses); all but `total` are detached values."""
    return _WeakLoss.apply(logits, images, labels, float(lam_ce), float(lam_cut), float(lam_bnd), float(sigma_cut),
                           float(sigma_boundary), float(sigma_space) if sigma_space else 0.0, int(window_size),
                           int(ignore_index))

"""Drop-in for the arithmetic of the reference's TraditionalModel/ExtraUtilities.py (dataset download needs
the network and is out of scope)."""
from . import functional as WF


def compute_iou_and_acc(pred_mask, true_mask):
    """reference ExtraUtilities.py:4-21: binary IoU and pixel accuracy of two (H, W) masks.  The three reductions
    (three `.item()` syncs in the reference) are one kernel and one read-back."""
    inter, union, correct = (int(v) for v in WF.iou_acc_counts(pred_mask, true_mask, batched=False)[0].tolist())
    total = true_mask.numel()  # the reference divides by the truth's own element count, also when shapes broadcast
    iou = inter / (union + 1e-8)
    acc = correct / total
    return iou, acc


def compute_iou_and_acc_batched(pred_masks, true_masks):
    """(B,H,W) masks -> (iou[B], acc[B]) float64 device tensors, no host synchronisation."""
    c = WF.iou_acc_counts(pred_masks, true_masks, batched=True).double()
    n = pred_masks.numel() // pred_masks.shape[0]
    return c[:, 0] / (c[:, 1] + 1e-8), c[:, 2] / n

"""Drop-in for the arithmetic of the reference's TraditionalModel/ExtraUtilities.py (dataset download needs
the network and is out of scope)."""


def compute_iou_and_acc(pred_mask, true_mask):
    """reference ExtraUtilities.py:4-21: binary IoU and pixel accuracy of two (H, W) masks."""
    pred_fg = (pred_mask > 0)
    true_fg = (true_mask > 0)

    intersection = (pred_fg & true_fg).sum().item()
    union = (pred_fg | true_fg).sum().item()
    correct = (pred_mask == true_mask).sum().item()
    total = true_mask.numel()

    iou = intersection / (union + 1e-8)
    acc = correct / total
    return iou, acc

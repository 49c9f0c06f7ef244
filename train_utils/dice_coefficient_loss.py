"""API mirror of the reference's train_utils/dice_coefficient_loss.py.

On the hot path the five loss terms are computed together (forward AND backward) by the fused CUDA kernel behind
`criterion`; the per-term functions here are forward-only views of that kernel, kept for API compatibility."""
import torch

from egm_unet_b200.loss import loss_terms


def build_target(target: torch.Tensor, num_classes: int = 2, ignore_index: int = -100):
    """dice_coefficient_loss.py:7-19.  Kept for callers; the fused kernels read `target` directly (no one-hot tensor)."""
    dice_target = target.clone()
    if ignore_index >= 0:
        ignore_mask = torch.eq(target, ignore_index)
        dice_target[ignore_mask] = 0
        dice_target = torch.nn.functional.one_hot(dice_target, num_classes).float()
        dice_target[ignore_mask] = ignore_index
    else:
        dice_target = torch.nn.functional.one_hot(dice_target, num_classes).float()
    return dice_target.permute(0, 3, 1, 2)


def dice_loss(x, target_index, multiclass: bool = True, ignore_index: int = -100):
    """1 - mean Dice (dice_coefficient_loss.py:53-57); `target_index` is the int64 label map."""
    return loss_terms(x, target_index, None, ignore_index)["dice"]


def laplace_loss(x, target_index=None):
    t = target_index if target_index is not None else torch.zeros(x.shape[0], x.shape[2], x.shape[3], dtype=torch.int64, device=x.device)
    return loss_terms(x, t)["laplace"]


def lap_loss(x, target):
    return loss_terms(x, target)["lap"]


def sobel_loss(y_true, y_pred):
    """Argument names as in the reference (:94): y_true = logits, y_pred = target."""
    return loss_terms(y_true, y_pred)["sobel"]

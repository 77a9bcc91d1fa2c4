"""Losses and metric of the hot path with the reference's API (midasmednet/unet/loss.py), computed by
fused CUDA kernels: no one-hot tensor, no transposed copies (loss.py:10-21,58-88 never materialise).

Only what the reference's drivers use is built: DiceLoss, dice_metric and
the helpers they are made of.  CELoss / WeightedCrossEntropyLoss / BCELossWrapper /
PixelWiseCrossEntropyLoss / LandmarkLoss are dead code in the reference (no caller) and are not ported.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


def _weight_on(weight, device):
    if weight is None:
        return None
    return weight.detach().to(device=device, dtype=torch.float32).contiguous()


class DiceLoss(nn.Module):
    """loss.py:91-130: mean over classes of 1 - 2 w_c I_c / clamp(sum p_c + sum t_c, eps); the class weight
    multiplies the intersection only (loss.py:44-45)."""

    def __init__(self, epsilon=1e-5, weight=None, ignore_index=None, sigmoid_normalization=False,
                 skip_last_target=False):
        super().__init__()
        self.epsilon = epsilon
        self.register_buffer('weight', weight)
        self.ignore_index = ignore_index
        self.sigmoid_normalization = sigmoid_normalization
        self.skip_last_target = skip_last_target
        if ignore_index is not None or skip_last_target:
            raise NotImplementedError("ignore_index / skip_last_target are not used by any reference driver")

    def forward(self, input, target):
        assert target.dim() == 4, "target must be (N, D, H, W) class indices (loss.py:66)"
        assert input.dim() == 5 and input.shape[0] == target.shape[0] and input.shape[2:] == target.shape[1:], \
            "'input' and 'target' must have the same shape"
        loss, _dice = ops.DiceLossFn.apply(input, target, _weight_on(self.weight, input.device), self.epsilon,
                                           self.sigmoid_normalization)
        return loss


def dice_metric(logits, labels):
    """loss.py:51-55: unweighted per-class soft Dice, returns (C,)."""
    with torch.no_grad():
        _loss, dice, _sums, _, _ = ops.k_dice_fwd(logits, labels, None, 1e-5, False)
    return dice


class CrossEntropyLoss(nn.Module):
    """Drop-in for ``torch.nn.CrossEntropyLoss(weight=w)`` as wired at segmentation.py:49 / landmarks.py:49."""

    def __init__(self, weight=None):
        super().__init__()
        self.register_buffer('weight', weight)

    def forward(self, input, target):
        return ops.CrossEntropyFn.apply(input, target, _weight_on(self.weight, input.device))

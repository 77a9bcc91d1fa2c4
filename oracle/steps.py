"""Restatement of the reference training steps (test oracle; see oracle/__init__.py).

segmentation.py / landmarks.py are not importable here (configargparse, matplotlib, pytorch_lightning
missing -- SURVEY.md section 0 item 5), so their <=25-line hot-path logic is restated:

  SegmentationNet.training_step ....... segmentation.py:58-65
  LandmarkNet.training_step / loss .... landmarks.py:66-83, 125-134
  configure_optimizers (Adam defaults)  segmentation.py:119-120, landmarks.py:176-177
"""
from __future__ import annotations

import time

import torch

from . import loss as oloss
from . import unet as ounet


def _forward(kind, sd, x, **kw):
    if kind == "unet3d":
        return ounet.unet3d_forward(sd, x, **kw)
    if kind == "residual":
        return ounet.residual_unet3d_forward(sd, x, **kw)
    raise ValueError(kind)


def segmentation_step(kind, sd, batch, loss="DICE", loss_weight=None, **model_kw):
    """segmentation.py:58-65: inputs=batch['data'].float(); labels=batch['label'][:, -1].long()."""
    inputs = batch["data"].float()
    labels = batch["label"][:, -1].long()
    logits = _forward(kind, sd, inputs, **model_kw)
    w = None if loss_weight is None else torch.as_tensor(loss_weight, dtype=torch.float32)
    value = oloss.dice_loss(logits, labels, weight=w) if loss == "DICE" else \
        oloss.weighted_cross_entropy(logits, labels, w)
    return value, logits


def landmark_step(kind, sd, batch, loss_class="DICE", loss_class_weight=(0.05, 1.0),
                  loss_regression="L2", loss_regression_weight=(), **model_kw):
    """landmarks.py:66-83: heatmaps = label[:, :-1].float(); labels = label[:, -1].long();
    channel split [heatmaps | classes] at L = number of heatmap channels."""
    inputs = batch["data"].float()
    heatmaps = batch["label"][:, :-1].float()
    num_heatmaps = heatmaps.shape[1]
    labels = batch["label"][:, -1].long()
    outputs = _forward(kind, sd, inputs, **model_kw)
    total, class_loss, regression = oloss.landmark_loss(
        outputs[:, num_heatmaps:], outputs[:, :num_heatmaps], labels, heatmaps,
        torch.as_tensor(loss_class_weight, dtype=torch.float32), list(loss_regression_weight),
        loss_class, loss_regression)
    return total, class_loss, regression, outputs


def grads_of(value, sd):
    names = [k for k, v in sd.items() if v.requires_grad]
    grads = torch.autograd.grad(value, [sd[k] for k in names])
    return dict(zip(names, grads))


def leaf_state_dict(sd):
    return {k: v.detach().clone().float().requires_grad_(True) for k, v in sd.items()}


def time_training_steps(kind, sd, batch, steps=3, warmup=1, lr=1e-3, task="seg", **step_kw):
    """CPU baseline timing: forward + loss + backward + Adam.step.
    Returns (seconds per step list, last loss)."""
    leaf = leaf_state_dict(sd)
    opt = torch.optim.Adam(list(leaf.values()), lr=lr)
    times, last = [], None
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        if task == "seg":
            value, _ = segmentation_step(kind, leaf, batch, **step_kw)
        else:
            value = landmark_step(kind, leaf, batch, **step_kw)[0]
        value.backward()
        opt.step()
        last = float(value.detach())
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return times, last

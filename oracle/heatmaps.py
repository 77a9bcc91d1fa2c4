"""Builder-specified heatmap rendering and landmark extraction (test oracle; see oracle/__init__.py).

ABSENT from the reference: the reference reads heatmaps pre-rendered
as uint8 (dataset.py:261-262, 324-330; plots.py:124 vmax=255) with one Gaussian sigma per landmark
(predict.py:29).  The definitions below are therefore the specification, not a restatement:

  render:       h_l(x) = 255 * exp(-|x - p_l|^2 / (2 sigma_l^2)), truncated to uint8
  argmax:       first maximal flat index over D*H*W per (n, l) -> (d, h, w)
  soft-argmax:  sum_x x * softmax(beta * h_l)(x)

Parity for these is "unpinned by the reference"; the oracle is NumPy/PyTorch one-liners.
"""
from __future__ import annotations

import numpy as np
import torch


def render_heatmaps(points, sigmas, shape):
    """points: (N, L, 3) float voxel coordinates (d, h, w); sigmas: (L,); shape: (D, H, W).
    Returns uint8 (N, L, D, H, W).  Computed in fp32 exactly as the CUDA kernel does:
    v = 255 * expf(-(dd^2 + dh^2 + dw^2) / (2 sigma^2)), truncation toward zero."""
    points = np.asarray(points, dtype=np.float32)
    sigmas = np.asarray(sigmas, dtype=np.float32)
    d = np.arange(shape[0], dtype=np.float32)[:, None, None]
    h = np.arange(shape[1], dtype=np.float32)[None, :, None]
    w = np.arange(shape[2], dtype=np.float32)[None, None, :]
    out = np.zeros((points.shape[0], points.shape[1]) + tuple(shape), dtype=np.uint8)
    for n in range(points.shape[0]):
        for l in range(points.shape[1]):
            p = points[n, l]
            r2 = (d - p[0]) ** 2 + (h - p[1]) ** 2 + (w - p[2]) ** 2
            inv = np.float32(1.0) / (np.float32(2.0) * sigmas[l] * sigmas[l])
            out[n, l] = (np.float32(255.0) * np.exp(-(r2 * inv), dtype=np.float32)).astype(np.uint8)
    return out


def argmax_landmarks(heatmaps):
    """heatmaps: torch (N, L, D, H, W) any float/uint8 dtype -> int64 (N, L, 3) of (d, h, w);
    first maximal index (torch.argmax rule)."""
    n, l, d, h, w = heatmaps.shape
    flat = torch.argmax(heatmaps.reshape(n, l, -1).float(), dim=2)
    return torch.stack((flat // (h * w), (flat // w) % h, flat % w), dim=-1)


def soft_argmax_landmarks(heatmaps, beta=1.0):
    """Expected coordinate under softmax(beta * h): (N, L, 3) fp32 of (d, h, w)."""
    n, l, d, h, w = heatmaps.shape
    p = torch.softmax(beta * heatmaps.reshape(n, l, -1).double(), dim=2).reshape(n, l, d, h, w)
    cd = (p.sum((3, 4)) * torch.arange(d, dtype=torch.float64)).sum(-1)
    ch = (p.sum((2, 4)) * torch.arange(h, dtype=torch.float64)).sum(-1)
    cw = (p.sum((2, 3)) * torch.arange(w, dtype=torch.float64)).sum(-1)
    return torch.stack((cd, ch, cw), dim=-1).float()

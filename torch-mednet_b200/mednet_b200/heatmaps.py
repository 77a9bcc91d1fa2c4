"""Gaussian heatmap rendering and landmark extraction on the GPU (builder-specified; the reference reads
heatmaps pre-rendered as uint8 -- midasmednet/dataset.py:261-262 -- see oracle/heatmaps.py for the spec)."""
from __future__ import annotations

import torch

from . import _abi, ops
from ._abi import check, lib, make


def render_heatmaps(points, sigmas, shape):
    """points (N, L, 3) voxel coords (d,h,w), sigmas (L,), shape (D,H,W) -> uint8 (N, L, D, H, W)."""
    points = points.detach().float().contiguous()
    sigmas = torch.as_tensor(sigmas, dtype=torch.float32, device=points.device).contiguous()
    ops._need_cuda(points)
    n, l = points.shape[0], points.shape[1]
    out = torch.empty((n, l) + tuple(shape), dtype=torch.uint8, device=points.device)
    p = make("mednet_hmrender_params", points=points.data_ptr(), sigmas=sigmas.data_ptr(), out=out.data_ptr(), N=n, L=l,
             D=shape[0], H=shape[1], W=shape[2])
    check(lib().mednet_heatmap_render(_abi.C.byref(p), ops._stream()), "heatmap_render")
    ops._count()
    return out


def extract_landmarks(heatmaps, soft=False, beta=1.0):
    """heatmaps (N, L, D, H, W) float/bf16/uint8 -> argmax int64 (N, L, 3) [, soft-argmax fp32 (N, L, 3)]."""
    ops._need_cuda(heatmaps)
    heatmaps = heatmaps.contiguous()
    n, l, d, h, w = heatmaps.shape
    arg = torch.empty((n, l, 3), dtype=torch.int64, device=heatmaps.device)
    sft = torch.empty((n, l, 3), dtype=torch.float32, device=heatmaps.device) if soft else None
    p = make("mednet_landmark_params", heatmaps=heatmaps.data_ptr(), argmax=arg.data_ptr(),
             soft=None if sft is None else sft.data_ptr(), NL=n * l, D=d, H=h, W=w, dtype=ops._dt(heatmaps), beta=beta)
    ws = ops._ws(lib().mednet_landmark_workspace_bytes(_abi.C.byref(p)), heatmaps.device)
    check(lib().mednet_landmark_extract(_abi.C.byref(p), ws.data_ptr(), ws.numel(), ops._stream()), "landmark_extract")
    ops._count(4 if soft else 2)
    return (arg, sft) if soft else arg

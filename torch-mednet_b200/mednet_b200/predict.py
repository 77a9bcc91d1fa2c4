"""Sliding-window inference (examples/predict.py:82-97) with the volume, the tiles and the stitched result
resident on the GPU: gather tile -> network -> uint8 epilogue -> centre-crop scatter.  The reference moves
fp32 logits and int64 argmax to the host for every batch (predict.py:90-91); here only the final uint8
volume leaves the device.  Tiles are independent, so ranks take tiles round-robin with no data-path
collective; once per volume every rank all-gathers the centre crops of the other ranks' tiles (uint8, only the regions
each rank produced -- no reduction over the whole volume) and scatters them into its copy of the result.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _abi, ops
from ._abi import check, lib, make
from .dataset import grid_geometry


class SlidingWindowPredictor:
    def __init__(self, model, patch_size, patch_overlap, num_heatmaps, batch_size=1, rank=0, world=1):
        self.model = model
        self.patch_size = [int(v) for v in patch_size]
        self.patch_overlap = [int(v) for v in patch_overlap]
        self.num_heatmaps = int(num_heatmaps)
        self.batch_size = int(batch_size)
        self.rank, self.world = rank, world

    def tile_origins(self, img_size):
        return grid_geometry(img_size, self.patch_size, self.patch_overlap)[3]

    @torch.no_grad()
    def __call__(self, volume, combine=True):
        """volume: (C, X, Y, Z) tensor/ndarray (fp16/fp32/bf16).  Returns uint8 (L+1, X, Y, Z) on the device."""
        dev = next(self.model.parameters()).device
        vol = torch.as_tensor(np.ascontiguousarray(volume) if isinstance(volume, np.ndarray) else volume)
        vol = vol.to(dev, non_blocking=True)
        if vol.dtype not in (torch.float32, torch.bfloat16):
            vol = vol.float()
        vol = vol.contiguous()
        c, X, Y, Z = vol.shape
        origins = self.tile_origins((X, Y, Z))
        mine = origins[self.rank::self.world]
        cfg = self.model.cfg
        P, O = self.patch_size, self.patch_overlap
        out_channels = self.num_heatmaps + 1
        result = torch.zeros((out_channels, X, Y, Z), dtype=torch.uint8, device=dev)
        org_dev = torch.as_tensor(np.ascontiguousarray(mine), dtype=torch.int32, device=dev)
        self.tiles_done = 0
        share = combine and self.world > 1 and dist.is_initialized()
        crop = [P[k] - 2 * O[k] for k in range(3)]
        if share:                                   # centre crops of this rank's tiles, padded to the largest share
            n_max = -(-len(origins) // self.world)
            crops = torch.zeros((n_max, out_channels, *crop), dtype=torch.uint8, device=dev)
        for b0 in range(0, len(mine), self.batch_size):
            b = min(self.batch_size, len(mine) - b0)
            org = org_dev[b0:b0 + b].contiguous()
            tiles = torch.empty((b, P[0], P[1], P[2], c), dtype=cfg.dtype, device=dev)
            gp = make("mednet_tile_gather_params", volume=vol.data_ptr(), tiles=tiles.data_ptr(), origins=org.data_ptr(),
                      B=b, C=c, X=X, Y=Y, Z=Z, P0=P[0], P1=P[1], P2=P[2], O0=O[0], O1=O[1], O2=O[2],
                      src_dtype=ops._dt(vol), dst_dtype=ops._dt(tiles))
            check(lib().mednet_tile_gather(_abi.C.byref(gp), ops._stream()), "tile_gather")
            logits = self.model(tiles.permute(0, 4, 1, 2, 3))
            u8 = ops.k_predict_epilogue(logits, self.num_heatmaps)
            sp = make("mednet_tile_scatter_params", tiles=u8.data_ptr(), volume=result.data_ptr(), origins=org.data_ptr(),
                      B=b, Co=out_channels, X=X, Y=Y, Z=Z, P0=P[0], P1=P[1], P2=P[2], O0=O[0], O1=O[1], O2=O[2])
            check(lib().mednet_tile_scatter(_abi.C.byref(sp), ops._stream()), "tile_scatter")
            ops._count(2)
            if share:
                crops[b0:b0 + b] = u8[:, :, O[0]:P[0] - O[0], O[1]:P[1] - O[1], O[2]:P[2] - O[2]]
            self.tiles_done += b
        if share:
            gathered = torch.empty((self.world,) + tuple(crops.shape), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(gathered, crops)
            for r in range(self.world):
                theirs = origins[r::self.world]
                if r == self.rank or len(theirs) == 0:
                    continue
                org = torch.as_tensor(np.ascontiguousarray(theirs), dtype=torch.int32, device=dev)
                sp = make("mednet_tile_scatter_params", tiles=gathered[r].data_ptr(), volume=result.data_ptr(),
                          origins=org.data_ptr(), B=len(theirs), Co=out_channels, X=X, Y=Y, Z=Z, P0=crop[0], P1=crop[1],
                          P2=crop[2], O0=0, O1=0, O2=0)
                check(lib().mednet_tile_scatter(_abi.C.byref(sp), ops._stream()), "tile_scatter")
                ops._count()
        return result

#!/usr/bin/env python
"""Sliding-window inference -- re-hosted entry point of examples/predict.py.  The reference reads a Hydra config
(predict.py:20-35); Hydra is not in this image, so the same keys are flags: prediction.{patch_size, patch_overlap,
batch_size, checkpoint, model, data} and base.sigma (its length = number of heatmap channels, predict.py:29).
The inner loop (predict.py:82-97: forward, softmax->argmax->u8, clip->u8, stitch) runs on the device; ranks take tiles
round-robin (torchrun), the disjoint uint8 sub-volumes are combined once.

    python examples/predict.py --checkpoint /tmp/m/epoch=0.ckpt --model SegmentationNet --synthetic 160 160 128 \
        --patch_size 64 64 64 --patch_overlap 8 8 8 --batch_size 4 --output /tmp/pred.npy
"""
import argparse

import numpy as np
import torch

import _common  # noqa: F401  (path bootstrap)


def main(argv=None):
    from mednet_b200.landmarks import LandmarkNet, LandmarkUNet3D
    from mednet_b200.parallel import init_distributed
    from mednet_b200.predict import SlidingWindowPredictor
    from mednet_b200.segmentation import SegmentationNet, SegmentationUNet3D

    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--checkpoint", required=True)                                   # cfg.prediction.checkpoint
    ap.add_argument("--model", default="SegmentationNet",                            # cfg.prediction.model (:45-49)
                    choices=["SegmentationNet", "LandmarkNet", "SegmentationUNet3D", "LandmarkUNet3D"])
    ap.add_argument("--patch_size", type=int, nargs=3, default=[128, 128, 128])      # cfg.prediction.patch_size
    ap.add_argument("--patch_overlap", type=int, nargs=3, default=[16, 16, 16])      # cfg.prediction.patch_overlap
    ap.add_argument("--batch_size", type=int, default=4)                             # cfg.prediction.batch_size
    ap.add_argument("--sigma", type=float, nargs="*", default=[])                    # cfg.base.sigma
    ap.add_argument("--volume", default=None, help=".npy file holding a (C, X, Y, Z) volume")
    ap.add_argument("--synthetic", type=int, nargs=3, default=None, metavar=("X", "Y", "Z"))
    ap.add_argument("--output", default=None, help=".npy file for the uint8 (L+1, X, Y, Z) result (cfg.prediction.data)")
    a = ap.parse_args(argv)
    rank, local, world = init_distributed()
    dev = torch.device("cuda", local)
    cls = {"SegmentationNet": SegmentationNet, "LandmarkNet": LandmarkNet, "SegmentationUNet3D": SegmentationUNet3D,
           "LandmarkUNet3D": LandmarkUNet3D}[a.model]
    model = cls.load_from_checkpoint(a.checkpoint).to(dev)
    model.freeze()                                                                   # predict.py:50
    if a.volume:
        vol = np.load(a.volume)
    elif a.synthetic:
        vol = np.random.default_rng(0).standard_normal((model.in_channels, *a.synthetic)).astype(np.float16)
    else:
        raise SystemExit("give --volume FILE.npy or --synthetic X Y Z (HDF5/zarr readers are out of scope here)")
    pred = SlidingWindowPredictor(model, a.patch_size, a.patch_overlap, len(a.sigma), a.batch_size, rank, world)
    out = pred(vol)
    if rank == 0:
        print(f"predicted {tuple(out.shape)} uint8 from {pred.tiles_done} tiles on this rank (world {world})")
        if a.output:
            np.save(a.output, out.cpu().numpy())
    return out


if __name__ == "__main__":
    main()

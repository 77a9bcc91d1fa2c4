"""mednet_b200 -- B200-native (sm_100a) implementation of torch-mednet's UNet3D hot path.

Drop-in surface (same names as the reference package ``midasmednet``):

    from mednet_b200.unet.model import UNet3D, ResidualUNet3D
    from mednet_b200.unet.loss import DiceLoss, dice_metric
    from mednet_b200.segmentation import SegmentationNet
    from mednet_b200.landmarks import LandmarkNet
    from mednet_b200.dataset import grid_patch_generator, MedDataset, DataReaderHDF5   # reference constructor / readers
    from mednet_b200.sampler import GpuMedDataset          # MedDataset's random patches from HBM-resident volumes

``install_as_midasmednet()`` registers these modules under the reference's import paths so that unmodified
caller code (``from midasmednet.unet.model import UNet3D``) picks up the CUDA implementation.
"""
from __future__ import annotations

import sys

__version__ = "0.1.0"


def install_as_midasmednet():
    import importlib
    names = {"midasmednet": "mednet_b200", "midasmednet.unet": "mednet_b200.unet",
             "midasmednet.unet.model": "mednet_b200.unet.model",
             "midasmednet.unet.components": "mednet_b200.unet.components",
             "midasmednet.unet.loss": "mednet_b200.unet.loss", "midasmednet.segmentation": "mednet_b200.segmentation",
             "midasmednet.landmarks": "mednet_b200.landmarks", "midasmednet.dataset": "mednet_b200.dataset"}
    for alias, target in names.items():
        sys.modules[alias] = importlib.import_module(target)

#!/usr/bin/env python
"""Landmark heatmap-regression training -- re-hosted entry point of examples/train_ldmks.py (same flags, :33-59,
plus LandmarkNet.add_model_specific_args, midasmednet/landmarks.py:191-205).

    python examples/train_ldmks.py --synthetic 8 --out_channels 10 --patch_size 96 96 96 --max_epochs 1 \
        --loss_regression_weight 0.001 0.015 0.015 0.015 0.001 0.001 0.001 0.001
The number of heatmap channels is len(loss_regression_weight) (landmarks.py:55-57); the remaining output channels are
classes.
"""
import logging

import numpy as np
import torch

from _common import experiment_parser, parse_with_config, require_dataset, seed_everything, synthetic_cohort


def main(argv=None):
    from mednet_b200.dataset import SyntheticSegmentationDataset
    from mednet_b200.landmarks import LandmarkNet, LandmarkUNet3D
    from mednet_b200.trainer import Trainer

    parser = LandmarkNet.add_model_specific_args(experiment_parser("aorth_ldmks", heatmaps=True))
    parser = __import__("argparse").ArgumentParser(parents=[parser], description=__doc__)
    hparams = parse_with_config(parser, argv)
    seed_everything(hparams.seed)
    logging.getLogger().setLevel("INFO")                 # train_ldmks.py:70
    require_dataset(hparams, "train_ldmks")
    L = len(hparams.loss_regression_weight)
    K = hparams.out_channels - L
    if K < 2:
        raise SystemExit(f"out_channels ({hparams.out_channels}) must be >= len(loss_regression_weight) ({L}) + 2 classes")
    if hparams.gpu_sampler:                              # MedDataset(..., heatmap_group) of train_ldmks.py, volumes in HBM
        from mednet_b200.sampler import GpuMedDataset, IntensityAugmentation
        dev = torch.device("cuda", int(__import__("os").environ.get("LOCAL_RANK", "0")))

        def mk(n, seed, train=False):
            images, labels, hms = synthetic_cohort(n, hparams.gpu_sampler, hparams.in_channels, K, L, seed=seed)
            return GpuMedDataset(images, labels, hparams.patches_per_subject, hparams.patch_size, heatmaps=hms,
                                 class_probabilities=hparams.class_probabilities if train else None, device=dev,
                                 augmentation=IntensityAugmentation() if hparams.data_augmentation and train else None)
    else:
        def mk(n, seed, train=False):
            return SyntheticSegmentationDataset(n, hparams.patch_size, hparams.in_channels, K, num_heatmaps=L, seed=seed)
    cls = LandmarkUNet3D if hparams.arch == "unet3d" else LandmarkNet
    model = cls(hparams, training_dataset=mk(hparams.synthetic, hparams.seed, True),
                validation_dataset=mk(max(1, hparams.synthetic // 4), hparams.seed + 1))
    trainer = Trainer(gpus=hparams.gpus, max_epochs=hparams.max_epochs, default_root_dir=hparams.model_dir,
                      resume_from_checkpoint=hparams.resume, max_steps=hparams.max_steps)
    trainer.fit(model)
    return trainer


if __name__ == "__main__":
    main()

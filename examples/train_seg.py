#!/usr/bin/env python
"""Segmentation training -- re-hosted entry point of examples/train_seg.py (same flags, :34-59; same flow:
seeds :61-65, datasets :98-118, module :120-122, Trainer(gpus, max_epochs, default_root_dir, resume) :123-133).

    python examples/train_seg.py --synthetic 16 --patch_size 64 64 64 --batch_size 2 --max_epochs 1 --model_dir /tmp/m
    python examples/train_seg.py --synthetic 4 --gpu_sampler 160 160 128 --class_probabilities 0.3 0.7 \
           --patches_per_subject 8 --patch_size 64 64 64 --batch_size 2 --max_epochs 1   # patches drawn from HBM-resident volumes
    torchrun --nproc-per-node 8 examples/train_seg.py --gpus 8 --synthetic 64 ...       # data parallel
"""
import logging

import numpy as np
import torch

from _common import experiment_parser, parse_with_config, require_dataset, seed_everything, synthetic_cohort


def main(argv=None):
    from mednet_b200.dataset import SyntheticSegmentationDataset
    from mednet_b200.segmentation import SegmentationNet, SegmentationUNet3D
    from mednet_b200.trainer import Trainer

    parser = SegmentationNet.add_model_specific_args(experiment_parser("aorth"))
    parser = __import__("argparse").ArgumentParser(parents=[parser], description=__doc__)
    hparams = parse_with_config(parser, argv)
    seed_everything(hparams.seed)                        # train_seg.py:61-65 (the kernels are deterministic by construction)
    logging.getLogger().setLevel(hparams.log_level)
    require_dataset(hparams, "train_seg")
    if hparams.gpu_sampler:                              # MedDataset(...) of train_seg.py:98-118, volumes in HBM
        from mednet_b200.sampler import GpuMedDataset, IntensityAugmentation
        dev = torch.device("cuda", int(__import__("os").environ.get("LOCAL_RANK", "0")))

        def cohort(n, seed, train):
            images, labels, _ = synthetic_cohort(n, hparams.gpu_sampler, hparams.in_channels, hparams.out_channels, seed=seed)
            return GpuMedDataset(images, labels, hparams.patches_per_subject, hparams.patch_size,
                                 class_probabilities=hparams.class_probabilities if train else None, device=dev,
                                 augmentation=IntensityAugmentation() if hparams.data_augmentation and train else None)
        train_ds = cohort(hparams.synthetic, hparams.seed, True)               # train_seg.py:98-106: transform on the
        val_ds = cohort(max(1, hparams.synthetic // 4), hparams.seed + 1, False)  # training set only (:108-117)
    else:
        train_ds = SyntheticSegmentationDataset(hparams.synthetic, hparams.patch_size, hparams.in_channels,
                                                hparams.out_channels, seed=hparams.seed)
        val_ds = SyntheticSegmentationDataset(max(1, hparams.synthetic // 4), hparams.patch_size, hparams.in_channels,
                                              hparams.out_channels, seed=hparams.seed + 1)
    cls = SegmentationUNet3D if hparams.arch == "unet3d" else SegmentationNet
    model = cls(hparams, training_dataset=train_ds, validation_dataset=val_ds)
    trainer = Trainer(gpus=hparams.gpus, max_epochs=hparams.max_epochs, default_root_dir=hparams.model_dir,
                      resume_from_checkpoint=hparams.resume, max_steps=hparams.max_steps)
    trainer.fit(model)
    return trainer


if __name__ == "__main__":
    main()

"""SegmentationNet with the reference's PL-0.9 hooks (midasmednet/segmentation.py), B200-native inside.

The reference class derives from ResidualUNet3D (segmentation.py:22); `arch='unet3d'` selects the UNet3D
the north star names instead.  Hooks kept: training_step (:58-65), validation_step (:94-109),
validation_epoch_end (:111-117), configure_optimizers (:119-120), train/val_dataloader (:122-132).
``log_samples`` (matplotlib/Neptune PNG logging, :67-92) is out of scope.
"""
from __future__ import annotations

import argparse
import logging

import torch

from .optim import FusedAdam
from .parallel import sharded_loader
from .unet.loss import CrossEntropyLoss, DiceLoss, dice_metric
from .unet.model import ResidualUNet3D, UNet3D


def _make_net_base(arch):
    return {"residual": ResidualUNet3D, "unet3d": UNet3D}[arch]


class _TaskMixin:
    def configure_optimizers(self):
        return FusedAdam(self.parameters(), lr=self.learning_rate)

    def train_dataloader(self):
        """segmentation.py:122-127; under torch.distributed every rank gets its own share of the epoch."""
        return sharded_loader(self.training_dataset, self.batch_size, self.num_workers, shuffle=True,
                              epoch=getattr(self, "current_epoch", 0), seed=getattr(getattr(self, "hparams", None), "seed", 0) or 0)

    def val_dataloader(self):
        return sharded_loader(self.validation_dataset, self.batch_size, self.num_workers, shuffle=False)


def _segmentation_class(base):
    class _SegmentationNet(_TaskMixin, base):
        def __init__(self, hparams, training_dataset=None, validation_dataset=None, **kwargs):
            super().__init__(hparams.in_channels, hparams.out_channels, final_sigmoid=False, f_maps=hparams.fmaps,
                             **kwargs)
            self.hparams = hparams
            self.training_dataset = training_dataset
            self.validation_dataset = validation_dataset
            self.learning_rate = hparams.learning_rate
            self.num_workers = hparams.num_workers
            self.batch_size = hparams.batch_size
            self.out_channels = hparams.out_channels
            self.in_channels = hparams.in_channels
            if hasattr(hparams, 'loss'):                                  # segmentation.py:43-49
                assert hparams.loss in ['DICE', 'CE']
                loss_weight = torch.tensor(hparams.loss_weight)
                if hparams.loss == 'DICE':
                    self.loss = DiceLoss(weight=loss_weight)
                elif hparams.loss == 'CE':
                    self.loss = CrossEntropyLoss(weight=loss_weight)
            self.log_interval = hparams.log_interval if hasattr(hparams, 'log_interval') else 5
            self.log_vis_mip = hparams.log_vis_mip if hasattr(hparams, 'log_vis_mip') else 'mean'
            self.logger = logging.getLogger(__name__)

        def _labels(self, batch):
            lab = batch['label'][:, -1, ...]
            return lab if lab.dtype in (torch.uint8, torch.int64) else lab.long()

        def training_step(self, batch, batch_nb):
            inputs = batch['data']
            labels = self._labels(batch)                                  # uint8 class map read in place (no .long() copy)
            outputs = self(inputs)
            loss = self.loss(outputs, labels)
            # the reference calls loss.item() here (a device sync every step, segmentation.py:64); the value is
            # handed over as a 0-d tensor and materialised lazily by the trainer's logger instead
            return {'loss': loss, 'log': {"train_loss": loss.detach()}}

        def validation_step(self, batch, batch_nb):
            inputs = batch['data']
            labels = self._labels(batch)
            outputs = self(inputs)
            loss = self.loss(outputs, labels)
            per_channel_dice = dice_metric(outputs, labels)
            results = {'val_loss': loss}
            for c in range(self.out_channels):
                results[f'val_dice{c}'] = per_channel_dice[c]
            return results

        def validation_epoch_end(self, outputs):
            avg_loss = torch.stack([x['val_loss'] for x in outputs]).mean()
            logs = {"val_loss": avg_loss}
            for c in range(self.out_channels):
                logs[f"val_dice{c}"] = torch.stack([x[f"val_dice{c}"] for x in outputs]).mean()
            return {"val_loss": avg_loss, "log": logs, "progress_bar": logs}

        @staticmethod
        def add_model_specific_args(parent_parser):
            """The reference script calls this although SegmentationNet never defines it (quirk Q1); defined
            here with the flags segmentation.py reads (:30-49)."""
            parser = argparse.ArgumentParser(parents=[parent_parser], add_help=False)
            parser.add_argument("--learning_rate", type=float, default=0.001)
            parser.add_argument("--fmaps", type=int, default=32)
            parser.add_argument("--batch_size", type=int, default=4)
            parser.add_argument("--num_workers", type=int, default=4)
            parser.add_argument("--in_channels", type=int, default=1)
            parser.add_argument("--out_channels", type=int, default=2)
            parser.add_argument("--log_interval", type=int, default=5)
            parser.add_argument("--log_vis_mip", type=str, choices=['mean', 'max'], default='mean')
            parser.add_argument('--loss', choices=['DICE', 'CE'], default='DICE')
            parser.add_argument('--loss_weight', nargs='+', type=float, default=[0.05, 1.0])
            return parser

    return _SegmentationNet


SegmentationNet = _segmentation_class(ResidualUNet3D)
SegmentationNet.__name__ = SegmentationNet.__qualname__ = "SegmentationNet"
SegmentationUNet3D = _segmentation_class(UNet3D)
SegmentationUNet3D.__name__ = SegmentationUNet3D.__qualname__ = "SegmentationUNet3D"

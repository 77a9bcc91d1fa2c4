"""LandmarkNet with the reference's PL-0.9 hooks (midasmednet/landmarks.py), B200-native inside.

Output channels are [heatmaps | classes] (landmarks.py:74-75).  The whole loss -- per-channel weighted
MSE/L1 on the heatmap channels plus weighted Dice/CE on the class channels (:125-134) -- runs as fused
kernels on the un-split output tensor (ops.LandmarkLossFn).
"""
from __future__ import annotations

import argparse
import logging

import torch

from . import ops
from .segmentation import _TaskMixin
from .unet.loss import CrossEntropyLoss, DiceLoss, dice_metric
from .unet.model import ResidualUNet3D, UNet3D


def _landmark_class(base):
    class _LandmarkNet(_TaskMixin, base):
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
            if hasattr(hparams, 'loss_class'):                             # landmarks.py:43-57
                assert hparams.loss_class in ['DICE', 'CE']
                loss_class_weight = torch.tensor(hparams.loss_class_weight)
                if hparams.loss_class == 'DICE':
                    self.loss_class = DiceLoss(weight=loss_class_weight)
                elif hparams.loss_class == 'CE':
                    self.loss_class = CrossEntropyLoss(weight=loss_class_weight)
                assert hparams.loss_regression in ['L2', 'L1']
                self.loss_regression = hparams.loss_regression
                self.loss_regression_weight = list(hparams.loss_regression_weight)
                self.num_heatmaps = len(self.loss_regression_weight)
                self.register_buffer('_reg_weight', torch.tensor(self.loss_regression_weight, dtype=torch.float32),
                                     persistent=False)
            self.log_interval = hparams.log_interval if hasattr(hparams, 'log_interval') else 5
            self.log_vis_mip = hparams.log_vis_mip if hasattr(hparams, 'log_vis_mip') else 'mean'
            self.logger = logging.getLogger(__name__)

        @staticmethod
        def _split_batch(batch):
            label = batch['label']
            heatmaps = label[:, :-1, ...]
            if heatmaps.dtype not in (torch.uint8, torch.float32):
                heatmaps = heatmaps.float()
            labels = label[:, -1, ...]
            if labels.dtype not in (torch.uint8, torch.int64):
                labels = labels.long()
            return heatmaps, labels

        def loss(self, output_labels, output_heatmaps, labels, heatmaps):
            """Reference signature (landmarks.py:125): separate class / heatmap outputs."""
            class_loss = self.loss_class(output_labels, labels)
            regression_loss, _ = ops.HeatmapLossFn.apply(output_heatmaps, heatmaps.contiguous(),
                                                         self._reg_weight.to(output_heatmaps.device),
                                                         self.loss_regression == 'L1')
            return regression_loss + class_loss, class_loss, regression_loss

        def _fused_loss(self, outputs, labels, heatmaps):
            w = self.loss_class.weight
            w = None if w is None else w.detach().to(outputs.device, torch.float32).contiguous()
            return ops.LandmarkLossFn.apply(outputs, labels.contiguous(), heatmaps.contiguous(), w,
                                            self._reg_weight.to(outputs.device), isinstance(self.loss_class, CrossEntropyLoss),
                                            self.loss_regression == 'L1', getattr(self.loss_class, 'epsilon', 1e-5))

        def training_step(self, batch, batch_nb):
            inputs = batch['data']
            heatmaps, labels = self._split_batch(batch)
            outputs = self(inputs)
            assert heatmaps.shape[1] == self.num_heatmaps
            loss, class_loss, regression_loss = self._fused_loss(outputs, labels, heatmaps)
            logs = {"train_loss": loss.detach(), "class_loss": class_loss.detach(),
                    "regression_loss": regression_loss.detach()}          # reference: three .item() syncs (:80-82)
            return {'loss': loss, 'log': logs}

        def validation_step(self, batch, batch_nb):
            inputs = batch['data']
            heatmaps, labels = self._split_batch(batch)
            outputs = self(inputs)
            loss, class_loss, regression_loss = self._fused_loss(outputs, labels, heatmaps)
            per_channel_dice = dice_metric(outputs[:, self.num_heatmaps:, ...], labels)
            results = {'val_loss': loss, 'val_class_loss': class_loss, 'val_regression_loss': regression_loss}
            for c in range(self.out_channels - self.num_heatmaps):
                results[f'val_dice{c}'] = per_channel_dice[c]
            return results

        def validation_epoch_end(self, outputs):
            logs = {k: torch.stack([x[k] for x in outputs]).mean()
                    for k in ('val_loss', 'val_class_loss', 'val_regression_loss')}
            for c in range(self.out_channels - self.num_heatmaps):
                logs[f"val_dice{c}"] = torch.stack([x[f"val_dice{c}"] for x in outputs]).mean()
            return {"val_loss": logs['val_loss'], "log": logs, "progress_bar": logs}

        @staticmethod
        def add_model_specific_args(parent_parser):
            """landmarks.py:191-205 (same flags and defaults)."""
            parser = argparse.ArgumentParser(parents=[parent_parser], add_help=False)
            parser.add_argument("--learning_rate", type=float, default=0.001)
            parser.add_argument("--fmaps", type=int, default=64)
            parser.add_argument("--batch_size", type=int, default=4)
            parser.add_argument("--num_workers", type=int, default=4)
            parser.add_argument("--in_channels", type=int, default=1)
            parser.add_argument("--out_channels", type=int, default=1)
            parser.add_argument("--log_interval", type=int, default=5)
            parser.add_argument("--log_vis_mip", type=str, choices=['mean', 'max'], default='mean')
            parser.add_argument('--loss_class', choices=['DICE', 'CE'], default='DICE')
            parser.add_argument('--loss_class_weight', nargs='+', type=float, default=[0.05, 1.0])
            parser.add_argument('--loss_regression', choices=['L2', 'L1'], default='L2')
            parser.add_argument("--loss_regression_weight", type=float, nargs='+',
                                default=[0.001, 0.015, 0.015, 0.015, 0.001, 0.001])
            return parser

    return _LandmarkNet


LandmarkNet = _landmark_class(ResidualUNet3D)
LandmarkNet.__name__ = LandmarkNet.__qualname__ = "LandmarkNet"
LandmarkUNet3D = _landmark_class(UNet3D)
LandmarkUNet3D.__name__ = LandmarkUNet3D.__qualname__ = "LandmarkUNet3D"

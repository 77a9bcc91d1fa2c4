"""Minimal stand-in for the parts of pytorch-lightning 0.9 the reference modules rely on.

The reference models derive from ``pl.LightningModule`` (midasmednet/unet/model.py:2,11,113) and its
scripts drive them with ``pl.Trainer`` (examples/train_seg.py:126-132).  pytorch-lightning is not part of
this image, so this module provides the hooks the hot path needs -- ``freeze``, ``load_from_checkpoint``,
``hparams`` -- and ``mednet_b200.trainer.Trainer`` supplies the loop that calls ``training_step`` /
``validation_step`` / ``configure_optimizers`` with the PL-0.9 signatures.
"""
from __future__ import annotations

import argparse

import torch
from torch import nn


class LightningModule(nn.Module):
    current_epoch = 0
    global_step = 0

    def freeze(self):
        for p in self.parameters():
            p.requires_grad = False
        self.eval()

    def unfreeze(self):
        for p in self.parameters():
            p.requires_grad = True
        self.train()

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, **kwargs):
        """Reads a PL-0.9 style checkpoint: dict with 'state_dict' and the hparams Namespace under
        'hparams' / 'hyper_parameters' (examples/predict.py:46-50)."""
        ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
        hparams = ckpt.get("hparams", ckpt.get("hyper_parameters"))
        if isinstance(hparams, dict):
            hparams = argparse.Namespace(**hparams)
        model = cls(hparams, **kwargs) if hparams is not None else cls(**kwargs)
        model.load_state_dict(ckpt["state_dict"])
        return model

    def save_checkpoint(self, path, optimizer=None, epoch=0, global_step=0):
        ckpt = {"state_dict": self.state_dict(), "epoch": epoch, "global_step": global_step}
        hp = getattr(self, "hparams", None)
        if hp is not None:
            ckpt["hparams"] = vars(hp) if isinstance(hp, argparse.Namespace) else hp
        if optimizer is not None:
            ckpt["optimizer_states"] = [optimizer.state_dict()]
        torch.save(ckpt, path)

"""B200-native UNet3D / ResidualUNet3D: drop-in replacements for midasmednet/unet/model.py.

Same constructor and ``forward`` signatures (model.py:36-37,84 and :140-141,189), same attributes
(``encoders``, ``decoders``, ``final_conv``, ``final_activation``, ``testing``) and the same
``state_dict`` keys/shapes, so reference checkpoints load with ``load_state_dict``.  Inside, the network
runs as hand-written sm_100a kernels on NDHWC bf16 (default) or fp32 (validation mode) activations:

    model = UNet3D(1, 2, False)                      # bf16 compute, fp32 master weights
    model = UNet3D(1, 2, False, compute_dtype=torch.float32)   # fp32 validation mode (rel. err <= 1e-4)

``forward`` takes (N, Cin, D, H, W) and returns fp32 logits (N, Cout, D, H, W) (probabilities iff
``testing``), exactly like the reference.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from ..lightning_compat import LightningModule
from .components import ComputeConfig, Decoder, DoubleConv, Encoder, ExtResNetBlock, _ConvParams, to_ndhwc


def create_feature_maps(init_channel_number, number_of_fmaps):
    """model.py:7-8."""
    return [init_channel_number * 2 ** k for k in range(number_of_fmaps)]


class _FinalActivation(nn.Module):
    def __init__(self, sigmoid):
        super().__init__()
        self.sigmoid = sigmoid

    def forward(self, x):
        return ops.k_final_activation(x, self.sigmoid)

    def extra_repr(self):
        return "Sigmoid" if self.sigmoid else "Softmax(dim=1)"


class _UNetBase(LightningModule):
    def _setup(self, kwargs):
        self.testing = kwargs.get('testing', False)
        self.cfg = ComputeConfig(kwargs.get('compute_dtype', torch.bfloat16), kwargs.get('conv_impl', 'auto'))

    def set_compute_dtype(self, dtype, conv_impl=None):
        """bf16 (production) or fp32 (validation mode; convolutions on the fp32 CUDA-core path)."""
        self.cfg.dtype = dtype
        if conv_impl is not None:
            self.cfg.conv_impl = conv_impl
        return self

    def _run(self, x):
        """Encoder/decoder walk of model.py:84-103.  When every block ends in conv+activation and every consumer of
        a block output (pool, upsample/concat join, final 1x1x1 conv) can apply the activation derivative itself, the
        convolutions skip their separate activation-backward pass (`defer`, see include/mednet_b200.h)."""
        x = to_ndhwc(x, self.cfg)
        blocks = [e.basic_module for e in self.encoders] + [d.basic_module for d in self.decoders]
        grad = torch.is_grad_enabled()
        defer = grad and all(getattr(b, 'out_act', 0) for b in blocks) and \
            all(e.accepts_in_act() for e in self.encoders[1:]) and all(d.accepts_in_act() for d in self.decoders)
        feats, acts, act = [], [], 0
        for i, encoder in enumerate(self.encoders):
            if i > 0 and grad and encoder.pooling is not None and len(self.decoders) > 0:
                # x feeds both this encoder's pool and a decoder's join: one fused backward for the two gradients
                pooled, skip = ops.MaxPoolSkipFn.apply(x, act)
                feats[0] = skip
                x = encoder.basic_module.run(pooled, defer=True) if defer else encoder.basic_module.run(pooled)
            elif defer:
                x = encoder.run(x, in_act=act, defer=True)
            else:
                x = encoder.run(x)
            act = encoder.basic_module.out_act if defer else 0
            feats.insert(0, x)
            acts.insert(0, act)
        for decoder, skip, skip_act in zip(self.decoders, feats[1:], acts[1:]):
            if defer:
                x = decoder.run(skip, x, skip_act=skip_act, x_act=act, defer=True)
                act = decoder.basic_module.out_act
            else:
                x = decoder.run(skip, x)
        return ops.Conv1x1Fn.apply(x, self.final_conv.weight, self.final_conv.bias, act)


class UNet3D(_UNetBase):
    """model.py:11-110.  ``f_maps`` int -> 4 levels (model.py:44-46)."""

    def __init__(self, in_channels, out_channels, final_sigmoid, f_maps=64, layer_order='gcr', num_groups=8, **kwargs):
        super().__init__()
        self._setup(kwargs)
        if isinstance(f_maps, int):
            f_maps = create_feature_maps(f_maps, number_of_fmaps=4)
        encoders = []
        for i, out_feature_num in enumerate(f_maps):
            if i == 0:
                encoder = Encoder(in_channels, out_feature_num, apply_pooling=False, basic_module=DoubleConv,
                                  conv_layer_order=layer_order, num_groups=num_groups, cfg=self.cfg)
            else:
                encoder = Encoder(f_maps[i - 1], out_feature_num, basic_module=DoubleConv,
                                  conv_layer_order=layer_order, num_groups=num_groups, cfg=self.cfg)
            encoders.append(encoder)
        self.encoders = nn.ModuleList(encoders)
        decoders = []
        reversed_f_maps = list(reversed(f_maps))
        for i in range(len(reversed_f_maps) - 1):
            in_feature_num = reversed_f_maps[i] + reversed_f_maps[i + 1]
            out_feature_num = reversed_f_maps[i + 1]
            decoders.append(Decoder(in_feature_num, out_feature_num, basic_module=DoubleConv,
                                    conv_layer_order=layer_order, num_groups=num_groups, cfg=self.cfg))
        self.decoders = nn.ModuleList(decoders)
        self.final_conv = _ConvParams(f_maps[0], out_channels, 1, bias=True)
        self.final_activation = _FinalActivation(bool(final_sigmoid))

    def forward(self, x):
        x = self._run(x)
        if self.testing:
            x = self.final_activation(x)
        return x


class ResidualUNet3D(_UNetBase):
    """model.py:113-214.  ``f_maps`` int -> 5 levels (model.py:148-150); spatial sizes must be divisible
    by 2**(levels-1) (summation join, components.py:284)."""

    def __init__(self, in_channels, out_channels, final_sigmoid, f_maps=32, conv_layer_order='cge', num_groups=8,
                 skip_final_activation=False, **kwargs):
        super().__init__()
        self._setup(kwargs)
        if isinstance(f_maps, int):
            f_maps = create_feature_maps(f_maps, number_of_fmaps=5)
        encoders = []
        for i, out_feature_num in enumerate(f_maps):
            if i == 0:
                encoder = Encoder(in_channels, out_feature_num, apply_pooling=False, basic_module=ExtResNetBlock,
                                  conv_layer_order=conv_layer_order, num_groups=num_groups, cfg=self.cfg)
            else:
                encoder = Encoder(f_maps[i - 1], out_feature_num, basic_module=ExtResNetBlock,
                                  conv_layer_order=conv_layer_order, num_groups=num_groups, cfg=self.cfg)
            encoders.append(encoder)
        self.encoders = nn.ModuleList(encoders)
        decoders = []
        reversed_f_maps = list(reversed(f_maps))
        for i in range(len(reversed_f_maps) - 1):
            decoders.append(Decoder(reversed_f_maps[i], reversed_f_maps[i + 1], basic_module=ExtResNetBlock,
                                    conv_layer_order=conv_layer_order, num_groups=num_groups, cfg=self.cfg))
        self.decoders = nn.ModuleList(decoders)
        self.final_conv = _ConvParams(f_maps[0], out_channels, 1, bias=True)
        if not skip_final_activation:
            self.final_activation = _FinalActivation(bool(final_sigmoid))
        else:
            self.final_activation = None

    def forward(self, x):
        x = self._run(x)
        if self.testing and self.final_activation is not None:
            x = self.final_activation(x)
        return x

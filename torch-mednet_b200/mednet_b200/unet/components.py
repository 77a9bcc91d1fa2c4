"""Building blocks of the B200-native U-Nets: same constructor signatures, attribute names and
``state_dict`` keys as midasmednet/unet/components.py, executed as fused CUDA kernels on NDHWC tensors.

Reference map (file:line in /root/reference/midasmednet/unet/components.py):
  create_conv / SingleConv ... :12-67, :70-90      order-string layer factory
  DoubleConv ................. :93-133
  ExtResNetBlock ............. :136-180
  Encoder .................... :183-226
  Decoder .................... :229-287

Tensors crossing a module boundary are logically (N, C, D, H, W); internally they are channels-last-3d
(NDHWC memory) in the compute dtype.  Any NCDHW tensor is accepted and converted on entry.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from .. import ops

_NONLIN = "rle"


class ComputeConfig:
    """Shared by every block of one network: compute dtype and convolution implementation."""

    def __init__(self, dtype=torch.bfloat16, conv_impl="auto"):
        self.dtype = dtype
        self.conv_impl = conv_impl


def to_ndhwc(x, cfg):
    """(N,C,D,H,W) tensor -> NDHWC-contiguous (N,D,H,W,C) tensor in the compute dtype (view when possible)."""
    if x.dim() != 5:
        raise ValueError(f"expected a 5-D (N,C,D,H,W) tensor, got {tuple(x.shape)}")
    if not x.is_cuda:
        raise RuntimeError("mednet_b200 runs on CUDA tensors only (no CPU fallback); move the module and inputs to a GPU")
    v = x.permute(0, 2, 3, 4, 1)
    if x.dtype == cfg.dtype and v.is_contiguous():
        return v
    return ops.ToChannelsLastFn.apply(x, cfg.dtype)


def from_ndhwc(y):
    return y.permute(0, 4, 1, 2, 3)


class _ConvParams(nn.Module):
    """Parameter holder with nn.Conv3d's default initialisation (kaiming_uniform(a=sqrt(5)))."""

    def __init__(self, in_channels, out_channels, kernel_size, bias, transposed=False):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size, self.transposed = in_channels, out_channels, kernel_size, transposed
        k = kernel_size
        shape = (in_channels, out_channels, k, k, k) if transposed else (out_channels, in_channels, k, k, k)
        self.weight = nn.Parameter(torch.empty(shape))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = self.weight.shape[1] * self.kernel_size ** 3
            bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)

    def extra_repr(self):
        kind = "ConvTranspose3d" if self.transposed else "Conv3d"
        return f"{kind}({self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, bias={self.bias is not None})"


class _GroupNormParams(nn.Module):
    def __init__(self, num_groups, num_channels, eps=1e-5):
        super().__init__()
        self.num_groups, self.num_channels, self.eps = num_groups, num_channels, eps
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))

    def extra_repr(self):
        return f"GroupNorm({self.num_groups}, {self.num_channels}, eps={self.eps})"


def conv3d(in_channels, out_channels, kernel_size, bias, padding=1):
    if kernel_size != 3 or padding != 1:
        raise NotImplementedError("mednet_b200 implements the 3x3x3 / padding 1 convolution of the reference nets")
    return _ConvParams(in_channels, out_channels, kernel_size, bias)


def create_conv(in_channels, out_channels, kernel_size, order, num_groups, padding=1):
    """Same contract as components.py:12-67: list of (name, module) for one order string."""
    assert 'c' in order, "Conv layer MUST be present"
    assert order[0] not in 'rle', 'Non-linearity cannot be the first operation in the layer'
    modules = []
    for i, char in enumerate(order):
        if char == 'r':
            modules.append(('ReLU', nn.Identity()))
        elif char == 'l':
            modules.append(('LeakyReLU', nn.Identity()))
        elif char == 'e':
            modules.append(('ELU', nn.Identity()))
        elif char == 'c':
            bias = not ('g' in order or 'b' in order)
            modules.append(('conv', conv3d(in_channels, out_channels, kernel_size, bias, padding=padding)))
        elif char == 'g':
            is_before_conv = i < order.index('c')
            num_channels = in_channels if is_before_conv else out_channels
            if num_channels < num_groups:
                num_groups = 1
            assert num_channels % num_groups == 0, f'Expected number of channels in input to be divisible by num_groups. num_channels={num_channels}, num_groups={num_groups}'
            modules.append(('groupnorm', _GroupNormParams(num_groups, num_channels)))
        elif char == 'b':
            raise NotImplementedError("order letter 'b' (BatchNorm3d) is not selected by any reference model and is not built")
        else:
            raise ValueError(f"Unsupported layer type '{char}'. MUST be one of ['b', 'g', 'r', 'l', 'e', 'c']")
    return modules


class SingleConv(nn.Module):
    """One order-string layer executed as fused kernels: 'g'+activation -> one GroupNorm kernel sequence,
    'c'+activation -> conv with the activation in its epilogue."""

    def __init__(self, in_channels, out_channels, kernel_size=3, order='crg', num_groups=8, padding=1, cfg=None):
        super().__init__()
        self.order = order
        self.cfg = cfg if cfg is not None else ComputeConfig()
        for name, module in create_conv(in_channels, out_channels, kernel_size, order, num_groups, padding=padding):
            self.add_module(name, module)
        # execution plan: list of (kind, fused_activation_code)
        plan, i = [], 0
        while i < len(order):
            ch = order[i]
            nxt = order[i + 1] if i + 1 < len(order) else ''
            if ch in 'cg':
                act = ops.ACT[nxt] if nxt and nxt in _NONLIN else 0
                plan.append((ch, act))
                i += 2 if act else 1
            else:
                plan.append(('a', ops.ACT[ch]))
                i += 1
        self._plan = plan

    # -- deferred activation derivative (include/mednet_b200.h): a layer that ENDS in conv+activation can leave
    #    act'(y) to the consumers of y; a layer that STARTS with a GroupNorm can apply it for its producer.
    @property
    def out_act(self):
        kind, act = self._plan[-1]
        return act if kind == 'c' else 0

    @property
    def accepts_in_act(self):
        return self._plan[0][0] == 'g'

    @property
    def starts_with_plain_groupnorm(self):
        return self._plan[0] == ('g', 0)

    def run(self, x, residual=None, final_act=0, in_act=0, defer=False, skip_first=False):
        """x: NDHWC tensor.  ``residual``/``final_act`` fuse `out += residual; act(out)` (components.py:177-178)
        into the last kernel of the layer.  ``in_act``: activation of x's producer whose derivative this layer's
        first op applies in backward; ``defer``: leave this layer's trailing conv activation derivative to the
        consumers; ``skip_first``: the first op (a GroupNorm) was already applied by a fused producer."""
        last = len(self._plan) - 1
        if in_act and not self.accepts_in_act:
            raise RuntimeError("in_act needs a layer that starts with GroupNorm")
        if (len(self._plan) == 2 and self._plan[0] == ('g', 0) and self._plan[1][0] == 'c' and not skip_first and
                residual is None and not final_act and not in_act and torch.is_grad_enabled() and
                ops.norm_conv_input_supported(x, self.conv.weight, self.conv.bias)):
            # 'g' 'c' [act] on a tensor that needs no gradient (the image): fused first layer, no dgrad / GroupNorm backward
            gn = self.groupnorm
            return ops.NormConvInputFn.apply(x, gn.weight, gn.bias, gn.num_groups, self.conv.weight, self._plan[1][1],
                                             self.cfg.conv_impl, bool(defer and self._plan[1][1]))
        for i, (kind, act) in enumerate(self._plan):
            if i == 0 and skip_first:
                continue
            fuse = (i == last) and (residual is not None or final_act)
            if fuse and act:
                raise RuntimeError("a residual join cannot follow a layer that already ends in a non-linearity")
            if kind == 'c':
                x = ops.Conv3x3Fn.apply(x, self.conv.weight, self.conv.bias, residual if fuse else None,
                                        final_act if fuse else act, self.cfg.conv_impl,
                                        bool(defer and i == last and not fuse and act))
            elif kind == 'g':
                gn = self.groupnorm
                x = ops.GroupNormActFn.apply(x, gn.weight, gn.bias, gn.num_groups, final_act if fuse else act,
                                             residual if fuse else None, in_act if i == 0 else 0)
            else:
                x = ops.ActFn.apply(x, act)
        return x

    def forward(self, x):
        return from_ndhwc(self.run(to_ndhwc(x, self.cfg)))


class DoubleConv(nn.Module):
    """components.py:93-133."""

    def __init__(self, in_channels, out_channels, encoder, kernel_size=3, order='crg', num_groups=8, cfg=None):
        super().__init__()
        self.cfg = cfg if cfg is not None else ComputeConfig()
        if encoder:
            conv1_in_channels = in_channels
            conv1_out_channels = out_channels // 2
            if conv1_out_channels < in_channels:
                conv1_out_channels = in_channels
            conv2_in_channels, conv2_out_channels = conv1_out_channels, out_channels
        else:
            conv1_in_channels, conv1_out_channels = in_channels, out_channels
            conv2_in_channels, conv2_out_channels = out_channels, out_channels
        self.add_module('SingleConv1', SingleConv(conv1_in_channels, conv1_out_channels, kernel_size, order, num_groups,
                                                   cfg=self.cfg))
        self.add_module('SingleConv2', SingleConv(conv2_in_channels, conv2_out_channels, kernel_size, order, num_groups,
                                                   cfg=self.cfg))

    @property
    def out_act(self):
        return self.SingleConv2.out_act

    @property
    def accepts_in_act(self):
        return self.SingleConv1.accepts_in_act

    def run(self, x, in_act=0, defer=False, skip_first=False):
        inner = self.SingleConv1.out_act if self.SingleConv2.accepts_in_act else 0
        h = self.SingleConv1.run(x, in_act=in_act, defer=bool(inner), skip_first=skip_first)
        return self.SingleConv2.run(h, in_act=inner, defer=defer)

    def forward(self, x):
        return from_ndhwc(self.run(to_ndhwc(x, self.cfg)))


class ExtResNetBlock(nn.Module):
    """components.py:136-180; the residual add and the trailing non-linearity are fused into conv3's
    last kernel."""

    def __init__(self, in_channels, out_channels, kernel_size=3, order='cge', num_groups=8, cfg=None, **kwargs):
        super().__init__()
        self.cfg = cfg if cfg is not None else ComputeConfig()
        self.conv1 = SingleConv(in_channels, out_channels, kernel_size=kernel_size, order=order, num_groups=num_groups,
                                cfg=self.cfg)
        self.conv2 = SingleConv(out_channels, out_channels, kernel_size=kernel_size, order=order, num_groups=num_groups,
                                cfg=self.cfg)
        n_order = order
        for c in 'rel':
            n_order = n_order.replace(c, '')
        self.conv3 = SingleConv(out_channels, out_channels, kernel_size=kernel_size, order=n_order,
                                num_groups=num_groups, cfg=self.cfg)
        if 'l' in order:
            self._act = ops.ACT['l']
        elif 'e' in order:
            self._act = ops.ACT['e']
        else:
            self._act = ops.ACT['r']
        self.non_linearity = nn.Identity()

    out_act = 0                  # ends in GroupNorm(+residual)+activation: nothing to defer
    accepts_in_act = False

    def run(self, x, in_act=0, defer=False):
        assert not in_act and not defer
        out = self.conv1.run(x)
        residual = out
        out = self.conv2.run(out)
        return self.conv3.run(out, residual=residual, final_act=self._act)

    def forward(self, x):
        return from_ndhwc(self.run(to_ndhwc(x, self.cfg)))


class _MaxPool(nn.Module):
    def forward(self, x, in_act=0):
        return ops.MaxPoolFn.apply(x, in_act)


class Encoder(nn.Module):
    """components.py:183-226."""

    def __init__(self, in_channels, out_channels, conv_kernel_size=3, apply_pooling=True, pool_kernel_size=(2, 2, 2),
                 pool_type='max', basic_module=DoubleConv, conv_layer_order='crg', num_groups=8, cfg=None):
        super().__init__()
        assert pool_type in ['max', 'avg']
        self.cfg = cfg if cfg is not None else ComputeConfig()
        if apply_pooling:
            if pool_type != 'max' or tuple(pool_kernel_size) != (2, 2, 2):
                raise NotImplementedError("only MaxPool3d(2) (the pooling every reference model selects) is built")
            self.pooling = _MaxPool()
        else:
            self.pooling = None
        self.basic_module = basic_module(in_channels, out_channels, encoder=True, kernel_size=conv_kernel_size,
                                         order=conv_layer_order, num_groups=num_groups, cfg=self.cfg)

    def accepts_in_act(self):
        return self.pooling is not None or self.basic_module.accepts_in_act

    def run(self, x, in_act=0, defer=False):
        """in_act: activation of x's producer (derivative applied by the pool / first GroupNorm in backward)."""
        if self.pooling is not None:
            x = self.pooling(x, in_act)
            in_act = 0
        if in_act or defer:
            return self.basic_module.run(x, in_act=in_act, defer=defer)
        return self.basic_module.run(x)

    def forward(self, x):
        return from_ndhwc(self.run(to_ndhwc(x, self.cfg)))


class Decoder(nn.Module):
    """components.py:229-287: DoubleConv -> nearest upsample + concat (one kernel); otherwise learned
    ConvTranspose3d + summation join (skip add fused into the transposed-conv epilogue)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, scale_factor=(2, 2, 2), basic_module=DoubleConv,
                 conv_layer_order='crg', num_groups=8, cfg=None):
        super().__init__()
        self.cfg = cfg if cfg is not None else ComputeConfig()
        if basic_module == DoubleConv:
            self.upsample = None
        else:
            if kernel_size != 3 or tuple(scale_factor) != (2, 2, 2):
                raise NotImplementedError("only ConvTranspose3d(k3, s2, p1, op1) is built")
            self.upsample = _ConvParams(in_channels, out_channels, kernel_size, bias=True, transposed=True)
            in_channels = out_channels
        self.basic_module = basic_module(in_channels, out_channels, encoder=False, kernel_size=kernel_size,
                                         order=conv_layer_order, num_groups=num_groups, cfg=self.cfg)

    def accepts_in_act(self):
        return self.upsample is None

    def run(self, encoder_features, x, skip_act=0, x_act=0, defer=False):
        """skip_act / x_act: activations of the producers of the two inputs whose derivatives are deferred to
        this decoder's join; defer: leave the trailing conv activation derivative to the consumer."""
        if self.upsample is None:
            first = self.basic_module.SingleConv1 if isinstance(self.basic_module, DoubleConv) else None
            if first is not None and first.starts_with_plain_groupnorm and ops.upcat_gn_supported(encoder_features, x):
                gn = first.groupnorm
                dc = self.basic_module
                if (len(first._plan) == 2 and first._plan[1][0] == 'c' and first.conv.bias is None and
                        ops.upconv_supported(encoder_features, x, first.conv.weight)):
                    # upsample-aware join: GroupNorm with split outputs, conv over skip + 8-tap conv over the COARSE low part
                    sn, ln = ops.UpcatGroupNormSplitFn.apply(encoder_features, x, gn.weight, gn.bias, gn.num_groups,
                                                             skip_act, x_act)
                    inner = first.out_act if dc.SingleConv2.accepts_in_act else 0
                    h = ops.UpConvJoinFn.apply(sn, ln, first.conv.weight, first._plan[1][1], bool(inner))
                    return dc.SingleConv2.run(h, in_act=inner, defer=defer)
                # GroupNorm over the virtual concat: the (N, Cs+Cl, S) concat tensor is never written
                xn = ops.UpcatGroupNormFn.apply(encoder_features, x, gn.weight, gn.bias, gn.num_groups, skip_act, x_act)
                return self.basic_module.run(xn, defer=defer, skip_first=True)
            x = ops.UpsampleConcatFn.apply(encoder_features, x, skip_act, x_act)
        else:
            assert not skip_act and not x_act
            x = ops.ConvTranspose3x3Fn.apply(x, self.upsample.weight, self.upsample.bias, encoder_features,
                                             self.cfg.conv_impl)
        if defer:
            return self.basic_module.run(x, defer=True)
        return self.basic_module.run(x)

    def forward(self, encoder_features, x):
        return from_ndhwc(self.run(to_ndhwc(encoder_features, self.cfg), to_ndhwc(x, self.cfg)))

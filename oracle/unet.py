"""Functional fp32 restatement of the reference U-Nets (test oracle; see oracle/__init__.py).

Follows /root/reference/midasmednet/unet/model.py and components.py.  The network is
evaluated directly from a ``state_dict`` with ``torch.nn.functional`` calls so that the
oracle shares no module code with the product.

Citations (reference file:line):
  feature-map ladder ............ model.py:7-8, :44-46 (UNet3D: 4 levels), :148-150 (Residual: 5)
  order-string layer factory .... components.py:12-67
  GroupNorm channel/group rule .. components.py:45-57
  conv bias rule ................ components.py:41-44
  DoubleConv channel rule ....... components.py:114-126
  ExtResNetBlock ................ components.py:146-180
  Encoder (pool then block) ..... components.py:203-226
  Decoder (nearest+cat / convT+add) components.py:247-287
  final 1x1x1 conv + test-time activation model.py:77-82,102-108 / :179-187,207-212
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

_NONLIN = "rle"


def feature_ladder(f_maps, levels):
    """model.py:7-8 -- geometric ladder when ``f_maps`` is an int."""
    if isinstance(f_maps, int):
        return [f_maps * 2 ** k for k in range(levels)]
    return list(f_maps)


def _nonlinearity(x, ch):
    if ch == "r":                      # components.py:35-36
        return F.relu(x)
    if ch == "l":                      # components.py:37-38 (slope 0.1)
        return F.leaky_relu(x, 0.1)
    if ch == "e":                      # components.py:39-40
        return F.elu(x)
    raise ValueError(ch)


def groups_for(channels, num_groups):
    """components.py:53-56."""
    g = num_groups if channels >= num_groups else 1
    assert channels % g == 0
    return g


class Storage:
    """Where a 16-bit pipeline rounds.  The reference keeps every tensor in fp32 (``Storage()`` = identity).
    ``Storage.bf16()`` restates the SAME network with bfloat16 *storage*: every activation tensor that is
    materialised between two ops (network input, GroupNorm(+activation) output, convolution(+activation)
    output, residual join output) and every 3x3x3 convolution weight is rounded to bf16, all arithmetic stays
    fp32 -- exactly what `model.to(torch.bfloat16)` / autocast does to the reference on a GPU, and what the
    product's bf16 mode implements.  Used by the bf16 parity gate: this random-initialised network amplifies a
    single bf16 rounding of its INPUT into ~1-2 % of logit error (tests/test_oracle_golden.py pins that), so
    "bf16 result vs fp32 reference" measures the chaos of the network, not the correctness of the kernels;
    "bf16 result vs the reference evaluated with bf16 storage" measures the kernels."""

    def __init__(self, act=None, weight=None):
        self.act = act if act is not None else (lambda t: t)
        self.weight = weight if weight is not None else (lambda t: t)

    @staticmethod
    def bf16(master_weights=False):
        """``master_weights``: the convolution weights are rounded to bf16 for the arithmetic but their gradient stays
        fp32 (fp32 master weights, what the product and mixed-precision training keep); by default the gradient takes the
        cast's own backward and is rounded to bf16 as well (``model.bfloat16()``)."""
        r = lambda t: t.to(torch.bfloat16).to(torch.float32)
        rw = (lambda t: t + (r(t) - t).detach()) if master_weights else r
        return Storage(r, rw)


_FP32 = Storage()


def single_conv(x, sd, prefix, order, num_groups, st=_FP32, join=None, trace=None):
    """One order-string layer (components.py:12-67, :70-90).

    Note the reference re-binds ``num_groups`` inside its loop (components.py:53-54, quirk
    Q13); with a single 'g' per order string this has no further effect.
    ``join``: optional (residual, non-linearity letter) applied after the last op of the layer
    (components.py:177-178) before the result is stored.
    ``trace``: optional list; receives ("layer", prefix, input, output) -- the tensors a layer-wise (teacher-forced)
    parity test feeds to / expects from the product's layer.
    """
    assert "c" in order and order[0] not in _NONLIN
    n = len(order)
    x_in = x
    for i, ch in enumerate(order):
        if ch == "c":
            bias = sd.get(prefix + "conv.bias")       # present only without g/b (components.py:43)
            x = F.conv3d(x, st.weight(sd[prefix + "conv.weight"]), bias, padding=1)
        elif ch == "g":
            c = x.shape[1]
            x = F.group_norm(x, groups_for(c, num_groups), sd[prefix + "groupnorm.weight"],
                             sd[prefix + "groupnorm.bias"], eps=1e-5)
        elif ch in _NONLIN:
            x = _nonlinearity(x, ch)
        else:
            raise ValueError(f"Unsupported layer type '{ch}'")
        last = i == n - 1
        if last and join is not None:
            x = _nonlinearity(x + join[0], join[1])
        # a tensor is materialised after an op unless a non-linearity (fused into that op) follows
        if last or order[i + 1] not in _NONLIN:
            x = st.act(x)
    if trace is not None:
        trace.append(("layer", prefix, x_in, x))
    return x


def double_conv(x, sd, prefix, order, num_groups, st=_FP32, trace=None):
    """components.py:114-133 -- channel counts are implied by the weights."""
    x = single_conv(x, sd, prefix + "SingleConv1.", order, num_groups, st, trace=trace)
    return single_conv(x, sd, prefix + "SingleConv2.", order, num_groups, st, trace=trace)


def ext_resnet_block(x, sd, prefix, order, num_groups, st=_FP32):
    """components.py:168-180."""
    out = single_conv(x, sd, prefix + "conv1.", order, num_groups, st)
    residual = out
    out = single_conv(out, sd, prefix + "conv2.", order, num_groups, st)
    n_order = "".join(c for c in order if c not in _NONLIN)      # components.py:154-156
    act = "l" if "l" in order else ("e" if "e" in order else "r")   # components.py:161-166
    return single_conv(out, sd, prefix + "conv3.", n_order, num_groups, st, join=(residual, act))


def unet3d_forward(sd, x, f_maps=64, layer_order="gcr", num_groups=8, testing=False,
                   final_sigmoid=False, storage=_FP32, trace=None):
    """UNet3D.forward, model.py:84-110.  ``trace``: see single_conv; additionally receives ("pool", i, in, out),
    ("join", j, skip, low, concat) and ("final", input, logits)."""
    f_maps = feature_ladder(f_maps, 4)
    st = storage
    x = st.act(x)
    skips = []
    for i in range(len(f_maps)):
        if i > 0:
            x_in = x
            x = F.max_pool3d(x, 2)                               # components.py:210,224
            if trace is not None:
                trace.append(("pool", i, x_in, x))
        x = double_conv(x, sd, f"encoders.{i}.basic_module.", layer_order, num_groups, st, trace=trace)
        skips.insert(0, x)
    skips = skips[1:]
    for j, skip in enumerate(skips):
        low = x
        x = F.interpolate(x, size=skip.shape[2:], mode="nearest")   # components.py:277-278
        x = torch.cat((skip, x), dim=1)                              # components.py:280
        if trace is not None:
            trace.append(("join", j, skip, low, x))
        x = double_conv(x, sd, f"decoders.{j}.basic_module.", layer_order, num_groups, st, trace=trace)
    x_in = x
    x = F.conv3d(x, sd["final_conv.weight"], sd["final_conv.bias"])   # model.py:102
    if trace is not None:
        trace.append(("final", x_in, x))
    if testing:
        x = torch.sigmoid(x) if final_sigmoid else torch.softmax(x, dim=1)
    return x


def residual_unet3d_forward(sd, x, f_maps=32, conv_layer_order="cge", num_groups=8, testing=False,
                            final_sigmoid=False, skip_final_activation=False, storage=_FP32):
    """ResidualUNet3D.forward, model.py:189-214."""
    f_maps = feature_ladder(f_maps, 5)
    st = storage
    x = st.act(x)
    skips = []
    for i in range(len(f_maps)):
        if i > 0:
            x = F.max_pool3d(x, 2)
        x = ext_resnet_block(x, sd, f"encoders.{i}.basic_module.", conv_layer_order, num_groups, st)
        skips.insert(0, x)
    skips = skips[1:]
    for j, skip in enumerate(skips):
        x = F.conv_transpose3d(x, st.weight(sd[f"decoders.{j}.upsample.weight"]), sd[f"decoders.{j}.upsample.bias"],
                               stride=2, padding=1, output_padding=1)   # components.py:259-264
        x = st.act(x + skip)                                           # components.py:284
        x = ext_resnet_block(x, sd, f"decoders.{j}.basic_module.", conv_layer_order, num_groups, st)
    x = F.conv3d(x, sd["final_conv.weight"], sd["final_conv.bias"])    # model.py:207
    if testing and not skip_final_activation:
        x = torch.sigmoid(x) if final_sigmoid else torch.softmax(x, dim=1)
    return x


# ----------------------------------------------------------------------------------------------
# state_dict construction with PyTorch default initialisation (conv: kaiming_uniform(a=sqrt 5),
# GroupNorm gamma=1, beta=0 -- SURVEY.md section 8(b)).  Used when the tests need weights without a
# product module at hand (e.g. the CPU baseline in bench.py).
# ----------------------------------------------------------------------------------------------
def _conv_init(cout, cin, k, bias, gen, transposed=False):
    shape = (cin, cout, k, k, k) if transposed else (cout, cin, k, k, k)
    fan_in = shape[1] * k ** 3
    bound = 1.0 / fan_in ** 0.5
    w = (torch.rand(shape, generator=gen) * 2 - 1) * bound
    b = (torch.rand(shape[1] if transposed else cout, generator=gen) * 2 - 1) * bound if bias else None
    return w, b


def _single_conv_sd(sd, prefix, cin, cout, order, gen):
    has_norm = "g" in order or "b" in order
    w, b = _conv_init(cout, cin, 3, not has_norm, gen)
    sd[prefix + "conv.weight"] = w
    if b is not None:
        sd[prefix + "conv.bias"] = b
    if "g" in order:
        c = cin if order.index("g") < order.index("c") else cout
        sd[prefix + "groupnorm.weight"] = torch.ones(c)
        sd[prefix + "groupnorm.bias"] = torch.zeros(c)


def make_unet3d_state_dict(in_channels, out_channels, f_maps=64, layer_order="gcr", seed=0):
    gen = torch.Generator().manual_seed(seed)
    f_maps = feature_ladder(f_maps, 4)
    sd = {}
    for i, f in enumerate(f_maps):
        cin = in_channels if i == 0 else f_maps[i - 1]
        mid = max(f // 2, cin)                                          # components.py:116-122
        _single_conv_sd(sd, f"encoders.{i}.basic_module.SingleConv1.", cin, mid, layer_order, gen)
        _single_conv_sd(sd, f"encoders.{i}.basic_module.SingleConv2.", mid, f, layer_order, gen)
    rev = list(reversed(f_maps))
    for j in range(len(rev) - 1):
        cin, cout = rev[j] + rev[j + 1], rev[j + 1]                     # model.py:66-68
        _single_conv_sd(sd, f"decoders.{j}.basic_module.SingleConv1.", cin, cout, layer_order, gen)
        _single_conv_sd(sd, f"decoders.{j}.basic_module.SingleConv2.", cout, cout, layer_order, gen)
    w, b = _conv_init(out_channels, f_maps[0], 1, True, gen)
    sd["final_conv.weight"], sd["final_conv.bias"] = w, b
    return sd


def make_residual_unet3d_state_dict(in_channels, out_channels, f_maps=32, conv_layer_order="cge", seed=0):
    gen = torch.Generator().manual_seed(seed)
    f_maps = feature_ladder(f_maps, 5)
    n_order = "".join(c for c in conv_layer_order if c not in _NONLIN)
    sd = {}

    def block(prefix, cin, cout):
        _single_conv_sd(sd, prefix + "conv1.", cin, cout, conv_layer_order, gen)
        _single_conv_sd(sd, prefix + "conv2.", cout, cout, conv_layer_order, gen)
        _single_conv_sd(sd, prefix + "conv3.", cout, cout, n_order, gen)

    for i, f in enumerate(f_maps):
        block(f"encoders.{i}.basic_module.", in_channels if i == 0 else f_maps[i - 1], f)
    rev = list(reversed(f_maps))
    for j in range(len(rev) - 1):
        w, b = _conv_init(rev[j + 1], rev[j], 3, True, gen, transposed=True)
        sd[f"decoders.{j}.upsample.weight"], sd[f"decoders.{j}.upsample.bias"] = w, b
        block(f"decoders.{j}.basic_module.", rev[j + 1], rev[j + 1])
    w, b = _conv_init(out_channels, f_maps[0], 1, True, gen)
    sd["final_conv.weight"], sd["final_conv.bias"] = w, b
    return sd


def conv_flops_per_voxel_unet3d(in_channels, out_channels, f_maps=64):
    """Algorithmic conv FLOPs per input voxel: returns (fwd, train)."""
    f_maps = feature_ladder(f_maps, 4)
    fwd = 0.0
    first = None
    for i, f in enumerate(f_maps):
        cin = in_channels if i == 0 else f_maps[i - 1]
        mid = max(f // 2, cin)
        scale = 1.0 / 8 ** i
        a = 2 * 27 * cin * mid * scale
        if first is None:
            first = a
        fwd += a + 2 * 27 * mid * f * scale
    rev = list(reversed(f_maps))
    for j in range(len(rev) - 1):
        scale = 1.0 / 8 ** (len(f_maps) - 2 - j)
        fwd += (2 * 27 * (rev[j] + rev[j + 1]) * rev[j + 1] + 2 * 27 * rev[j + 1] ** 2) * scale
    fwd += 2 * f_maps[0] * out_channels
    return fwd, 3 * fwd - first

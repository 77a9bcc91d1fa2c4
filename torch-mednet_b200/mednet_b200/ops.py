"""Operator layer: thin launch wrappers over the C ABI (``k_*``) and the ``torch.autograd.Function``s
that the modules compose.  PyTorch is used for device memory, streams and the autograd tape only --
every arithmetic step is a kernel of libmednet_b200.so.  No fallback: CPU tensors raise.

Activations are NDHWC-contiguous tensors of shape (N, D, H, W, C) in the compute dtype (bf16 or fp32).
"""
from __future__ import annotations

import os

import torch

from . import _abi
from ._abi import check, lib, make

ACT = {None: 0, "none": 0, "r": 1, "relu": 1, "l": 2, "leaky": 2, "e": 3, "elu": 3}
ACT_PARAM = {0: 0.0, 1: 0.0, 2: 0.1, 3: 1.0}            # LeakyReLU slope 0.1 (components.py:38), ELU alpha 1
_DT = {torch.float32: 0, torch.bfloat16: 1, torch.uint8: 2, torch.int64: 3}
IMPL = {"auto": 0, "simt": 1, "tcgen05": 2}

launch_count = 0          # kernels-API calls issued (bench.py reads it for `gpu_launches`)
wgrad_events = []         # same, per wgrad launch (tensor-core or SIMT, whichever ran)
conv_events = None        # when a list: (executed FLOPs, start event, end event, reference-op FLOPs) per tcgen05 conv launch


def _dt(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"mednet_b200: unsupported dtype {t.dtype}")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("mednet_b200 kernels run on CUDA tensors only (no CPU fallback)")


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _count(n=1):
    global launch_count
    launch_count += n


# =============================================================================================
# tcgen05 addressing calibration (once per process)
# =============================================================================================
tcgen05_variants = {}      # row_bytes -> dict(enabled, dense_halo, base_offset_mode) as registered with the library
_tc_calibrated = False


def _probe(a, rb, shift, sbo, bo_mode):
    k = rb // 2
    out = torch.full((128, k), -1.0, dtype=torch.float32, device=a.device)
    r = lib().mednet_tcgen05_probe(a.data_ptr(), rb, a.shape[0], shift, sbo, bo_mode, out.data_ptr(), _stream())
    check(r, "tcgen05_probe")
    torch.cuda.synchronize()
    return out


def probe_report(device="cuda"):
    """Runs every (swizzle width, pitch, base-offset) variant and returns {name: bool} -- diagnostics."""
    report = {}
    for rb in (128, 64, 32):
        k = rb // 2
        g = torch.Generator(device="cpu").manual_seed(rb)
        a = torch.randint(0, 256, (256, k), generator=g).to(torch.bfloat16).to(device)
        af = a.float()
        rows = torch.arange(128, device=device)
        for name, pitch, shifts in (("aligned", 8, (0, 8)), ("padded", 16, (1, 2, 8)), ("dense", 10, (1, 2, 11, 22))):
            for bo in (1, 0, 2):
                ok = True
                for shift in shifts:
                    got = _probe(a, rb, shift, pitch * rb, bo)
                    want = af[shift + (rows // 8) * pitch + rows % 8]
                    ok = ok and bool(torch.equal(got, want))
                report[f"rb{rb}.{name}.bo{bo}"] = ok
    return report


def calibrate_tcgen05(force=False):
    """Finds, per swizzle width, a descriptor variant under which shifted windows of a TMA-written tile are
    addressed correctly and registers it with the library (mednet_tcgen05_configure).  Preference: dense
    halo (one TMA box per plane, least shared memory) over padded rows."""
    global _tc_calibrated
    if _tc_calibrated and not force:
        return tcgen05_variants
    _tc_calibrated = True
    if not torch.cuda.is_available() or not lib().mednet_device_has_tcgen05():
        return tcgen05_variants
    rep = probe_report()
    for rb in (128, 64, 32):
        choice = None
        for name, dense in (("dense", 1), ("padded", 0)):
            for bo in (1, 0, 2):
                if rep[f"rb{rb}.{name}.bo{bo}"] and rep[f"rb{rb}.aligned.bo{bo}"]:
                    choice = dict(enabled=1, dense_halo=dense, base_offset_mode=bo)
                    break
            if choice:
                break
        if choice is None:
            choice = dict(enabled=0, dense_halo=0, base_offset_mode=1)
            import warnings
            warnings.warn(f"mednet_b200: no working tcgen05 descriptor variant for {rb}-byte rows; "
                          "3x3x3 convolutions with that channel chunking run on the CUDA-core kernel")
        check(lib().mednet_tcgen05_configure(rb, choice["enabled"], choice["dense_halo"], choice["base_offset_mode"]),
              "tcgen05_configure")
        tcgen05_variants[rb] = choice
    tcgen05_variants["report"] = rep
    return tcgen05_variants


# =============================================================================================
# launch wrappers
# =============================================================================================
def k_layout(src, to_channels_last, dst_dtype):
    """NCDHW (N,C,D,H,W) <-> NDHWC (N,D,H,W,C) with dtype cast."""
    _need_cuda(src)
    src = src.contiguous()
    if to_channels_last:
        n, c = src.shape[0], src.shape[1]
        sp = tuple(src.shape[2:])
        dst = torch.empty((n,) + sp + (c,), dtype=dst_dtype, device=src.device)
    else:
        n, c = src.shape[0], src.shape[-1]
        sp = tuple(src.shape[1:-1])
        dst = torch.empty((n, c) + sp, dtype=dst_dtype, device=src.device)
    s = 1
    for v in sp:
        s *= v
    p = make("mednet_layout_params", src=_ptr(src), dst=_ptr(dst), N=n, C=c, S=s, src_dtype=_dt(src),
             dst_dtype=_DT[dst_dtype], to_channels_last=int(to_channels_last))
    check(lib().mednet_layout_convert(_abi.C.byref(p), _stream()), "layout_convert")
    _count()
    return dst


def k_channel_pad(x, c_dst):
    """(..., Cs) -> (..., c_dst): zero padding (c_dst > Cs) or truncation of the channel axis."""
    _need_cuda(x)
    x = _c(x)
    cs = x.shape[-1]
    out = torch.empty(tuple(x.shape[:-1]) + (c_dst,), dtype=x.dtype, device=x.device)
    p = make("mednet_chpad_params", src=_ptr(x), dst=_ptr(out), rows=x.numel() // cs, Cs=cs, Cd=c_dst, dtype=_dt(x))
    check(lib().mednet_channel_pad(_abi.C.byref(p), _stream()), "channel_pad")
    _count()
    return out


def k_intensity_augment(x, coef, dst_dtype):
    """Brightness / gamma / contrast chain of the training scripts on a patch batch (include/mednet_b200.h).
    x: (B, P0, P1, P2, C) fp32 NDHWC; coef: (B, 2 + 2C) fp32 device tensor of the host-drawn decisions."""
    _need_cuda(x)
    if x.dtype != torch.float32 or coef.dtype != torch.float32:
        raise TypeError("mednet_b200: intensity augmentation takes fp32 patches and coefficients")
    x, coef = _c(x), _c(coef)
    b, c = x.shape[0], x.shape[-1]
    if tuple(coef.shape) != (b, 2 + 2 * c):
        raise ValueError(f"coef must be ({b}, {2 + 2 * c}), got {tuple(coef.shape)}")
    y = torch.empty(x.shape, dtype=dst_dtype, device=x.device)
    p = make("mednet_intensity_aug_params", x=_ptr(x), y=_ptr(y), coef=_ptr(coef), V=x.numel() // (b * c), B=b, C=c,
             dst_dtype=_DT[dst_dtype])
    ws = _ws(lib().mednet_intensity_augment_workspace_bytes(_abi.C.byref(p)), x.device)
    check(lib().mednet_intensity_augment(_abi.C.byref(p), _ptr(ws), ws.numel(), _stream()), "intensity_augment")
    _count(3)
    return y


def conv_select_impl(x_shape, out_sp, K, Nout, dtype, gather, impl, x_ptr=0, w_ptr=0, y_ptr=0):
    n, di, hi, wi = x_shape[0], x_shape[1], x_shape[2], x_shape[3]
    if dtype == torch.bfloat16 and impl != "simt":
        calibrate_tcgen05()
    p = make("mednet_conv3d_params", x=x_ptr, w=w_ptr, y=y_ptr, N=n, Di=di, Hi=hi, Wi=wi, Do=out_sp[0], Ho=out_sp[1],
             Wo=out_sp[2], K=K, Nout=Nout, dtype=_DT[dtype], gather=gather, impl=IMPL[impl])
    r = lib().mednet_conv3d_select_impl(_abi.C.byref(p))
    if r < 0:
        check(r, "conv3d_select_impl")
    return r


def k_pack_weights(w, cin, cout, dtype, layout, transposed=False):
    _need_cuda(w)
    w = w.detach().contiguous().float()
    out = torch.empty(cin * cout * 27 * (8 if layout >= 4 else 1), dtype=dtype, device=w.device)
    p = make("mednet_wpack_params", w_oidhw=_ptr(w), w_packed=_ptr(out), Cin=cin, Cout=cout, dtype=_DT[dtype],
             layout=layout, transposed=int(transposed))
    check(lib().mednet_conv3d_pack_weights(_abi.C.byref(p), _stream()), "conv3d_pack_weights")
    _count()
    return out


def k_conv3(x, w_packed, nout, out_sp, gather, impl_id, bias=None, addend=None, act=0, y_f32=False):
    """y = act(gather_conv(x, w) + bias + addend); x (N,Di,Hi,Wi,K) -> y (N,Do,Ho,Wo,nout).
    y_f32: write the (unrounded) result as fp32, to be completed by a second launch that takes it as its fp32 addend."""
    _need_cuda(x, w_packed)
    n, di, hi, wi, k = x.shape
    y = torch.empty((n,) + tuple(out_sp) + (nout,), dtype=torch.float32 if y_f32 else x.dtype, device=x.device)
    addend_f32 = addend is not None and addend.dtype == torch.float32 and x.dtype != torch.float32
    p = make("mednet_conv3d_params", x=_ptr(x), w=_ptr(w_packed), bias=_ptr(bias), addend=_ptr(addend), y=_ptr(y), N=n,
             Di=di, Hi=hi, Wi=wi, Do=out_sp[0], Ho=out_sp[1], Wo=out_sp[2], K=k, Nout=nout, dtype=_dt(x), act=act,
             act_param=ACT_PARAM[act], gather=gather, impl=impl_id, y_f32=int(y_f32), addend_f32=int(addend_f32))
    ws = _ws(256, x.device)
    timing = conv_events is not None and impl_id == 2
    if timing:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib().mednet_conv3d_fprop(_abi.C.byref(p), _ptr(ws), ws.numel(), _stream()), "conv3d_fprop")
    if timing:
        e1.record()
        vo, vi = out_sp[0] * out_sp[1] * out_sp[2], di * hi * wi
        # taps per output voxel: EXECUTED by the kernel / ALGORITHMIC (what the reference op performs: 27 per fine voxel)
        taps, ref_taps = {0: (27, 27), 1: (27 / 8, 27 / 8), 2: (27, 27), 3: (8, 27), 4: (64, 216)}[gather]
        conv_events.append((2.0 * n * vo * k * nout * taps, e0, e1, 2.0 * n * vo * k * nout * ref_taps))
    _count()
    return y


def wgrad_params(a, b, gather, impl="auto", dw=None, dbias=None, ld=0, c0=0, transposed=False):
    n, da, ha, wa, ca = a.shape
    _, db, hb, wb, cb = b.shape
    return make("mednet_wgrad_params", a=_ptr(a), b=_ptr(b), dw=_ptr(dw), dbias=_ptr(dbias), N=n, Da=da, Ha=ha, Wa=wa,
                Db=db, Hb=hb, Wb=wb, Ca=ca, Cb=cb, dtype=_dt(a), gather=gather, impl=IMPL[impl], accumulate=0,
                dw_ld=ld, dw_c0=c0, dw_transposed=int(transposed))


def k_wgrad(a, b, gather, impl="auto", want_bias=False):
    """dw (Ca,Cb,3,3,3) fp32 = sum_rows a[row] (x) b[gather(row, tap)]; optional bias gradient."""
    _need_cuda(a, b)
    ca, cb = a.shape[-1], b.shape[-1]
    dw = torch.empty((ca, cb, 3, 3, 3), dtype=torch.float32, device=a.device)
    nb = ca if gather == 0 else cb
    dbias = torch.empty(nb, dtype=torch.float32, device=a.device) if want_bias else None
    p = wgrad_params(a, b, gather, impl, dw, dbias)
    ws = _ws(lib().mednet_conv3d_wgrad_workspace_bytes(_abi.C.byref(p)), a.device)
    timing = conv_events is not None
    if timing:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib().mednet_conv3d_wgrad(_abi.C.byref(p), _ptr(ws), ws.numel(), _stream()), "conv3d_wgrad")
    if timing:
        e1.record()
        n, d, h, w = a.shape[:4]
        wgrad_events.append((2.0 * n * d * h * w * ca * cb * 27, e0, e1))
    _count(2)
    return dw, dbias


# =============================================================================================
# asynchronous weight gradients
# =============================================================================================
# The weight gradient of a convolution depends only on (x, dpre) and is consumed only by the optimiser, while everything
# else in backward is a serial chain (dgrad -> GroupNorm backward -> ...).  When a parameter's .grad lives in a buffer whose
# owner promises to call sync_async_wgrad() before reading it (FusedAdam's flat gradient buffer does), its wgrad is enqueued
# on a SIDE stream and accumulated straight into that buffer: the tensor-core wgrad kernels then overlap the memory-bound
# GroupNorm / pool / join backward kernels of the following layers instead of running between them.
_side_streams = {}
_async_pending = set()          # device indices with weight gradients in flight on their side stream
async_grad_listener = None      # callable(param): told when a parameter's gradient has been ENQUEUED on the side stream
                                # (the data-parallel reducer counts its bucket down and launches the all-reduce behind it)


def side_stream(device=None):
    if isinstance(device, int):
        device = torch.device("cuda", device)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=key)
    return _side_streams[key]


def mark_async_grad(param, enabled=True):
    """Owner of param.grad promises to call sync_async_wgrad() before reading or overwriting it."""
    param._mednet_async_grad = bool(enabled)


def sync_async_wgrad(device=None):
    """Make the current stream of `device` (default: the current device) wait for every weight gradient enqueued on
    that device's side stream."""
    key = torch.cuda.current_device() if device is None else torch.device(device).index
    if key in _async_pending:
        torch.cuda.current_stream(key).wait_stream(side_stream(key))
        _async_pending.discard(key)


def _async_wgrad_ok(weight):
    g = weight.grad
    return (getattr(weight, "_mednet_async_grad", False) and g is not None and g.dtype == torch.float32 and
            g.is_contiguous() and g.is_cuda)


def k_wgrad_into(a, b, gather, impl, dw, accumulate=True, ld=0, c0=0, transposed=False):
    """Weight gradient accumulated (or written) straight into `dw` (fp32, PyTorch layout) on the CURRENT stream.
    ld / c0 / transposed: `dw` is a wider gradient tensor of `ld` channels per row and the result goes to the channel range
    starting at c0 (include/mednet_b200.h, mednet_wgrad_params)."""
    _need_cuda(a, b, dw)
    p = wgrad_params(a, b, gather, impl, dw, None, ld, c0, transposed)
    p.accumulate = int(accumulate)
    ws = _ws(lib().mednet_conv3d_wgrad_workspace_bytes(_abi.C.byref(p)), a.device)
    timing = conv_events is not None
    if timing:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib().mednet_conv3d_wgrad(_abi.C.byref(p), _ptr(ws), ws.numel(), _stream()), "conv3d_wgrad")
    if timing:
        e1.record()
        n, d, h, w = a.shape[:4]
        wgrad_events.append((2.0 * n * d * h * w * a.shape[-1] * b.shape[-1] * 27, e0, e1))
    _count(2)


def wgrad_async(a, b, gather, impl, weight):
    """Enqueue dW += wgrad(a, b) into weight.grad on the side stream (see the section comment)."""
    main, side = torch.cuda.current_stream(a.device), side_stream(a.device)
    side.wait_stream(main)                       # a / b (dpre, x) were produced on the main stream
    with torch.cuda.stream(side):
        k_wgrad_into(a, b, gather, impl, weight.grad, accumulate=True)
    a.record_stream(side)                        # the caching allocator must not recycle them before the side stream is done
    b.record_stream(side)
    _async_pending.add(a.device.index)
    if async_grad_listener is not None:
        async_grad_listener(weight)


def k_border_class_sums(x):
    """(27, C) fp32 sums of an NDHWC tensor over the voxels of each border class (include/mednet_b200.h)."""
    _need_cuda(x)
    x = _c(x)
    n, d, h, w, c = x.shape
    bins = torch.empty((27, c), dtype=torch.float32, device=x.device)
    p = make("mednet_border_sums_params", x=_ptr(x), bins=_ptr(bins), N=n, D=d, H=h, W=w, C=c, dtype=_dt(x))
    nbytes = lib().mednet_border_class_sums_workspace_bytes(_abi.C.byref(p))
    if nbytes == 0:
        raise _abi.MednetError("border_class_sums: unsupported shape (needs D, H, W >= 2)")
    ws = _ws(nbytes, x.device)
    check(lib().mednet_border_class_sums(_abi.C.byref(p), _ptr(ws), ws.numel(), _stream()), "border_class_sums")
    _count(2)
    return bins


def k_affine_input_grads(weight, ghat, bins, gamma, beta, conv_dtype):
    """dW, dgamma, dbeta of `conv(gamma * xhat + beta)` from ghat = wgrad(dpre, xhat) and the border-class sums of dpre."""
    _need_cuda(weight, ghat, bins, gamma, beta)
    cout, cin = weight.shape[0], weight.shape[1]
    w = weight.detach().contiguous().float()
    dw = torch.empty_like(w)
    dgamma = torch.empty(cin, dtype=torch.float32, device=w.device)
    dbeta = torch.empty(cin, dtype=torch.float32, device=w.device)
    p = make("mednet_affine_input_params", w=_ptr(w), ghat=_ptr(ghat), bins=_ptr(bins), gamma=_ptr(gamma), beta=_ptr(beta),
             dw=_ptr(dw), dgamma=_ptr(dgamma), dbeta=_ptr(dbeta), Cout=cout, Cin=cin, dtype=_DT[conv_dtype])
    check(lib().mednet_conv3d_affine_input_grads(_abi.C.byref(p), _stream()), "conv3d_affine_input_grads")
    _count()
    return dw, dgamma, dbeta


def k_conv1_fwd(x, w2d, bias):
    _need_cuda(x, w2d, bias)
    n, sp, cin = x.shape[0], tuple(x.shape[1:-1]), x.shape[-1]
    cout = w2d.shape[0]
    s = sp[0] * sp[1] * sp[2]
    y = torch.empty((n, cout) + sp, dtype=torch.float32, device=x.device)
    p = make("mednet_conv1_params", x=_ptr(x), w=_ptr(w2d), bias=_ptr(bias), y=_ptr(y), N=n, S=s, Cin=cin, Cout=cout,
             dtype=_dt(x))
    check(lib().mednet_conv1x1_fwd(_abi.C.byref(p), _stream()), "conv1x1_fwd")
    _count()
    return y


def k_conv1_bwd(x, w2d, dy, need_dx=True, in_act=0):
    _need_cuda(x, w2d, dy)
    n, sp, cin = x.shape[0], tuple(x.shape[1:-1]), x.shape[-1]
    cout = w2d.shape[0]
    s = sp[0] * sp[1] * sp[2]
    dx = torch.empty_like(x) if need_dx else None
    dw = torch.empty((cout, cin), dtype=torch.float32, device=x.device)
    db = torch.empty(cout, dtype=torch.float32, device=x.device)
    p = make("mednet_conv1_bwd_params", x=_ptr(x), w=_ptr(w2d), dy=_ptr(dy), dx=_ptr(dx), dw=_ptr(dw), db=_ptr(db), N=n,
             S=s, Cin=cin, Cout=cout, dtype=_dt(x), accumulate=0, in_act=in_act, in_act_param=ACT_PARAM[in_act])
    ws = _ws(lib().mednet_conv1x1_bwd_workspace_bytes(_abi.C.byref(p)), x.device)
    check(lib().mednet_conv1x1_bwd(_abi.C.byref(p), _ptr(ws), ws.numel(), _stream()), "conv1x1_bwd")
    _count(4)
    return dx, dw, db


def k_gn_fwd(x, gamma, beta, groups, act=0, residual=None, eps=1e-5):
    _need_cuda(x, gamma, beta)
    n, c = x.shape[0], x.shape[-1]
    s = x.numel() // (n * c)
    y = torch.empty_like(x)
    mean = torch.empty((n, groups), dtype=torch.float32, device=x.device)
    rstd = torch.empty((n, groups), dtype=torch.float32, device=x.device)
    p = make("mednet_gn_fwd_params", x=_ptr(x), gamma=_ptr(gamma), beta=_ptr(beta), residual=_ptr(residual), y=_ptr(y),
             mean=_ptr(mean), rstd=_ptr(rstd), N=n, S=s, C=c, G=groups, dtype=_dt(x), act=act,
             act_param=ACT_PARAM[act], eps=eps)
    ws = _ws(lib().mednet_groupnorm_fwd_workspace_bytes(_abi.C.byref(p)), x.device)
    check(lib().mednet_groupnorm_fwd(_abi.C.byref(p), _ptr(ws), ws.numel(), _stream()), "groupnorm_fwd")
    _count(3)
    return y, mean, rstd


def k_gn_bwd(x, y, dy, gamma, mean, rstd, groups, act=0, want_dresidual=False, in_act=0):
    _need_cuda(x, dy)
    n, c = x.shape[0], x.shape[-1]
    s = x.numel() // (n * c)
    dx = torch.empty_like(x)
    dres = torch.empty_like(x) if want_dresidual else None
    dgamma = torch.empty(c, dtype=torch.float32, device=x.device)
    dbeta = torch.empty(c, dtype=torch.float32, device=x.device)
    p = make("mednet_gn_bwd_params", x=_ptr(x), y=_ptr(y), dy=_ptr(dy), gamma=_ptr(gamma), mean=_ptr(mean),
             rstd=_ptr(rstd), dx=_ptr(dx), dresidual=_ptr(dres), dgamma=_ptr(dgamma), dbeta=_ptr(dbeta), N=n, S=s, C=c,
             G=groups, dtype=_dt(x), act=act, act_param=ACT_PARAM[act], accumulate=0, in_act=in_act,
             in_act_param=ACT_PARAM[in_act])
    ws = _ws(lib().mednet_groupnorm_bwd_workspace_bytes(_abi.C.byref(p)), x.device)
    check(lib().mednet_groupnorm_bwd(_abi.C.byref(p), _ptr(ws), ws.numel(), _stream()), "groupnorm_bwd")
    _count(4)
    return dx, dgamma, dbeta, dres


def k_act_fwd(x, act):
    _need_cuda(x)
    y = torch.empty_like(x)
    p = make("mednet_act_fwd_params", x=_ptr(x), y=_ptr(y), numel=x.numel(), dtype=_dt(x), act=act,
             act_param=ACT_PARAM[act])
    check(lib().mednet_act_fwd(_abi.C.byref(p), _stream()), "act_fwd")
    _count()
    return y


def k_act_bwd(y, dy, act):
    _need_cuda(y, dy)
    dx = torch.empty_like(dy)
    p = make("mednet_act_bwd_params", y=_ptr(y), dy=_ptr(dy), dx=_ptr(dx), numel=y.numel(), dtype=_dt(y), act=act,
             act_param=ACT_PARAM[act])
    check(lib().mednet_act_bwd(_abi.C.byref(p), _stream()), "act_bwd")
    _count()
    return dx


def k_pool_fwd(x):
    _need_cuda(x)
    n, d, h, w, c = x.shape
    y = torch.empty((n, d // 2, h // 2, w // 2, c), dtype=x.dtype, device=x.device)
    idx = torch.empty(y.shape, dtype=torch.uint8, device=x.device)
    p = make("mednet_pool_params", x=_ptr(x), y=_ptr(y), idx=_ptr(idx), N=n, D=d, H=h, W=w, C=c, dtype=_dt(x))
    check(lib().mednet_maxpool3d_fwd(_abi.C.byref(p), _stream()), "maxpool3d_fwd")
    _count()
    return y, idx


def k_pool_bwd(dy, idx, in_shape, y=None, in_act=0, addend=None):
    _need_cuda(dy, idx)
    n, d, h, w, c = in_shape
    dx = torch.empty(in_shape, dtype=dy.dtype, device=dy.device)
    p = make("mednet_pool_bwd_params", dy=_ptr(dy), idx=_ptr(idx), dx=_ptr(dx), N=n, D=d, H=h, W=w, C=c, dtype=_dt(dy),
             y=_ptr(y), in_act=in_act, in_act_param=ACT_PARAM[in_act], addend=_ptr(addend))
    check(lib().mednet_maxpool3d_bwd(_abi.C.byref(p), _stream()), "maxpool3d_bwd")
    _count()
    return dx


def k_pool_indices_i64(idx, in_shape):
    n, d, h, w, c = in_shape
    out = torch.empty((n, c, d // 2, h // 2, w // 2), dtype=torch.int64, device=idx.device)
    check(lib().mednet_maxpool3d_indices_i64(_ptr(idx), _ptr(out), n, d, h, w, c, _stream()), "maxpool3d_indices_i64")
    _count()
    return out


def k_upcat_fwd(skip, low):
    _need_cuda(skip, low)
    n, D, H, W, cs = skip.shape
    _, d, h, w, cl = low.shape
    out = torch.empty((n, D, H, W, cs + cl), dtype=low.dtype, device=low.device)
    p = make("mednet_upcat_params", skip=_ptr(skip), low=_ptr(low), out=_ptr(out), N=n, D=D, H=H, W=W, d=d, h=h, w=w,
             Cs=cs, Cl=cl, dtype=_dt(low))
    check(lib().mednet_upsample_concat_fwd(_abi.C.byref(p), _stream()), "upsample_concat_fwd")
    _count()
    return out


def k_upcat_bwd(dout, skip_shape, low_shape, skip=None, low=None, skip_act=0, low_act=0):
    _need_cuda(dout)
    n, D, H, W, cs = skip_shape
    _, d, h, w, cl = low_shape
    dskip = torch.empty(skip_shape, dtype=dout.dtype, device=dout.device)
    dlow = torch.empty(low_shape, dtype=dout.dtype, device=dout.device)
    p = make("mednet_upcat_bwd_params", dout=_ptr(dout), dskip=_ptr(dskip), dlow=_ptr(dlow), N=n, D=D, H=H, W=W, d=d,
             h=h, w=w, Cs=cs, Cl=cl, dtype=_dt(dout), skip=_ptr(skip), low=_ptr(low), skip_act=skip_act, low_act=low_act,
             skip_act_param=ACT_PARAM[skip_act], low_act_param=ACT_PARAM[low_act])
    check(lib().mednet_upsample_concat_bwd(_abi.C.byref(p), _stream()), "upsample_concat_bwd")
    _count(2)
    return dskip, dlow


def upcat_gn_supported(skip, low):
    """Exact 2x nearest upsampling (the only case the fused virtual-concat GroupNorm handles)."""
    return tuple(skip.shape[1:4]) == tuple(2 * v for v in low.shape[1:4]) and skip.numel() // skip.shape[0] < 2 ** 31


def k_upcat_gn_fwd(skip, low, gamma, beta, groups, eps=1e-5):
    _need_cuda(skip, low, gamma, beta)
    n, D, H, W, cs = skip.shape
    _, d, h, w, cl = low.shape
    y = torch.empty((n, D, H, W, cs + cl), dtype=low.dtype, device=low.device)
    mean = torch.empty((n, groups), dtype=torch.float32, device=low.device)
    rstd = torch.empty((n, groups), dtype=torch.float32, device=low.device)
    p = make("mednet_upcat_gn_fwd_params", skip=_ptr(skip), low=_ptr(low), gamma=_ptr(gamma), beta=_ptr(beta), y=_ptr(y),
             mean=_ptr(mean), rstd=_ptr(rstd), N=n, D=D, H=H, W=W, d=d, h=h, w=w, Cs=cs, Cl=cl, G=groups, dtype=_dt(low),
             eps=eps)
    ws = _ws(lib().mednet_upcat_groupnorm_fwd_workspace_bytes(_abi.C.byref(p)), low.device)
    check(lib().mednet_upcat_groupnorm_fwd(_abi.C.byref(p), _ptr(ws), ws.numel(), _stream()), "upcat_groupnorm_fwd")
    _count(4)
    return y, mean, rstd


def k_upcat_gn_bwd(skip, low, dy, gamma, mean, rstd, groups, skip_act=0, low_act=0):
    _need_cuda(skip, low, dy)
    n, D, H, W, cs = skip.shape
    _, d, h, w, cl = low.shape
    dskip, dlow = torch.empty_like(skip), torch.empty_like(low)
    dgamma = torch.empty(cs + cl, dtype=torch.float32, device=low.device)
    dbeta = torch.empty(cs + cl, dtype=torch.float32, device=low.device)
    p = make("mednet_upcat_gn_bwd_params", skip=_ptr(skip), low=_ptr(low), dy=_ptr(dy), gamma=_ptr(gamma), mean=_ptr(mean),
             rstd=_ptr(rstd), dskip=_ptr(dskip), dlow=_ptr(dlow), dgamma=_ptr(dgamma), dbeta=_ptr(dbeta), N=n, D=D, H=H,
             W=W, d=d, h=h, w=w, Cs=cs, Cl=cl, G=groups, dtype=_dt(low), accumulate=0, skip_act=skip_act, low_act=low_act,
             skip_act_param=ACT_PARAM[skip_act], low_act_param=ACT_PARAM[low_act])
    ws = _ws(lib().mednet_upcat_groupnorm_bwd_workspace_bytes(_abi.C.byref(p)), low.device)
    check(lib().mednet_upcat_groupnorm_bwd(_abi.C.byref(p), _ptr(ws), ws.numel(), _stream()), "upcat_groupnorm_bwd")
    _count(5)
    return dskip, dlow, dgamma, dbeta


def k_upcat_gn_split_fwd(skip, low, gamma, beta, groups, eps=1e-5):
    """GroupNorm over the virtual concat with SPLIT outputs: normalised skip (full resolution) and normalised low (coarse)."""
    _need_cuda(skip, low, gamma, beta)
    n, D, H, W, cs = skip.shape
    _, d, h, w, cl = low.shape
    ys, yl = torch.empty_like(skip), torch.empty_like(low)
    mean = torch.empty((n, groups), dtype=torch.float32, device=low.device)
    rstd = torch.empty((n, groups), dtype=torch.float32, device=low.device)
    p = make("mednet_upcat_gn_split_fwd_params", skip=_ptr(skip), low=_ptr(low), gamma=_ptr(gamma), beta=_ptr(beta),
             y_skip=_ptr(ys), y_low=_ptr(yl), mean=_ptr(mean), rstd=_ptr(rstd), N=n, D=D, H=H, W=W, d=d, h=h, w=w, Cs=cs, Cl=cl,
             G=groups, dtype=_dt(low), eps=eps)
    ws = _ws(lib().mednet_upcat_groupnorm_split_fwd_workspace_bytes(_abi.C.byref(p)), low.device)
    check(lib().mednet_upcat_groupnorm_split_fwd(_abi.C.byref(p), _ptr(ws), ws.numel(), _stream()), "upcat_groupnorm_split_fwd")
    _count(6)
    return ys, yl, mean, rstd


def k_upcat_gn_split_bwd(skip, low, dys, dyl, gamma, mean, rstd, groups, skip_act=0, low_act=0):
    _need_cuda(skip, low, dys, dyl)
    n, D, H, W, cs = skip.shape
    _, d, h, w, cl = low.shape
    dskip, dlow = torch.empty_like(skip), torch.empty_like(low)
    dgamma = torch.empty(cs + cl, dtype=torch.float32, device=low.device)
    dbeta = torch.empty(cs + cl, dtype=torch.float32, device=low.device)
    p = make("mednet_upcat_gn_split_bwd_params", skip=_ptr(skip), low=_ptr(low), dy_skip=_ptr(dys), dy_low=_ptr(dyl),
             gamma=_ptr(gamma), mean=_ptr(mean), rstd=_ptr(rstd), dskip=_ptr(dskip), dlow=_ptr(dlow), dgamma=_ptr(dgamma),
             dbeta=_ptr(dbeta), N=n, D=D, H=H, W=W, d=d, h=h, w=w, Cs=cs, Cl=cl, G=groups, dtype=_dt(low), accumulate=0,
             skip_act=skip_act, low_act=low_act, skip_act_param=ACT_PARAM[skip_act], low_act_param=ACT_PARAM[low_act])
    ws = _ws(lib().mednet_upcat_groupnorm_split_bwd_workspace_bytes(_abi.C.byref(p)), low.device)
    check(lib().mednet_upcat_groupnorm_split_bwd(_abi.C.byref(p), _ptr(ws), ws.numel(), _stream()), "upcat_groupnorm_split_bwd")
    _count(6)
    return dskip, dlow, dgamma, dbeta


def _logit_view(t):
    """(N, C, *spatial) tensor (possibly a channel slice) -> (tensor, N, C, S, batch_stride)."""
    n, c = t.shape[0], t.shape[1]
    s = t.numel() // (n * c)
    ok = t.dim() >= 3 and t[0, 0].is_contiguous() and (c == 1 or t.stride(1) == s)
    if not ok:
        t = t.contiguous()
    return t, n, c, s, (t.stride(0) if n > 1 else c * s)


def _check_class_weight(weight, c, what):
    """The kernels index weight[class]: a weight vector of another length is a caller error (the reference raises a
    shape mismatch in DiceLoss / a size check in nn.CrossEntropyLoss; e.g. --loss_weight 0.05 1.0 with 3 classes)."""
    if weight is not None and weight.numel() != c:
        raise RuntimeError(f"{what}: weight tensor should be defined for all {c} classes, got {weight.numel()} values")


def k_dice_fwd(logits, labels, weight, eps, sigmoid):
    _need_cuda(logits, labels)
    _check_class_weight(weight, logits.shape[1], "DiceLoss")
    logits, n, c, s, bs = _logit_view(logits)
    labels = labels.contiguous()
    dev = logits.device
    sums = torch.empty(3 * c, dtype=torch.float32, device=dev)
    dice = torch.empty(c, dtype=torch.float32, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    p = make("mednet_dice_params", logits=_ptr(logits), labels=_ptr(labels), weight=_ptr(weight), sums=_ptr(sums),
             dice=_ptr(dice), loss=_ptr(loss), N=n, S=s, batch_stride=bs, C=c, logits_dtype=_dt(logits),
             label_dtype=_dt(labels), sigmoid=int(sigmoid), epsilon=eps)
    ws = _ws(lib().mednet_dice_workspace_bytes(_abi.C.byref(p)), dev)
    check(lib().mednet_dice_fwd(_abi.C.byref(p), _ptr(ws), ws.numel(), _stream()), "dice_fwd")
    _count(2)
    return loss, dice, sums, logits, labels


def k_dice_bwd(logits, labels, weight, sums, grad_out, eps, sigmoid, out=None, out_batch_stride=None):
    logits, n, c, s, bs = _logit_view(logits)
    if out is None:
        out = torch.empty((n, c) + tuple(logits.shape[2:]), dtype=torch.float32, device=logits.device)
        out_batch_stride = c * s
    p = make("mednet_dice_bwd_params", logits=_ptr(logits), labels=_ptr(labels), weight=_ptr(weight), sums=_ptr(sums),
             grad_out=_ptr(grad_out), dlogits=_ptr(out), N=n, S=s, batch_stride=bs, batch_stride_out=out_batch_stride,
             C=c, logits_dtype=_dt(logits), label_dtype=_dt(labels), dlogits_dtype=_dt(out), sigmoid=int(sigmoid),
             epsilon=eps)
    check(lib().mednet_dice_bwd(_abi.C.byref(p), _stream()), "dice_bwd")
    _count()
    return out


def k_ce_fwd(logits, labels, weight):
    _need_cuda(logits, labels)
    _check_class_weight(weight, logits.shape[1], "CrossEntropyLoss")
    logits, n, c, s, bs = _logit_view(logits)
    labels = labels.contiguous()
    dev = logits.device
    sums = torch.empty(2, dtype=torch.float32, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    p = make("mednet_ce_params", logits=_ptr(logits), labels=_ptr(labels), weight=_ptr(weight), sums=_ptr(sums),
             loss=_ptr(loss), N=n, S=s, batch_stride=bs, C=c, logits_dtype=_dt(logits), label_dtype=_dt(labels))
    ws = _ws(lib().mednet_ce_workspace_bytes(_abi.C.byref(p)), dev)
    check(lib().mednet_ce_fwd(_abi.C.byref(p), _ptr(ws), ws.numel(), _stream()), "ce_fwd")
    _count(2)
    return loss, sums, logits, labels


def k_ce_bwd(logits, labels, weight, sums, grad_out, out=None, out_batch_stride=None):
    logits, n, c, s, bs = _logit_view(logits)
    if out is None:
        out = torch.empty((n, c) + tuple(logits.shape[2:]), dtype=torch.float32, device=logits.device)
        out_batch_stride = c * s
    p = make("mednet_ce_bwd_params", logits=_ptr(logits), labels=_ptr(labels), weight=_ptr(weight), sums=_ptr(sums),
             grad_out=_ptr(grad_out), dlogits=_ptr(out), N=n, S=s, batch_stride=bs, batch_stride_out=out_batch_stride,
             C=c, logits_dtype=_dt(logits), label_dtype=_dt(labels), dlogits_dtype=_dt(out))
    check(lib().mednet_ce_bwd(_abi.C.byref(p), _stream()), "ce_bwd")
    _count()
    return out


def k_hm_fwd(pred, target, weight, l1):
    _need_cuda(pred, target, weight)
    _check_class_weight(weight, pred.shape[1], "heatmap regression loss")
    pred, n, L, s, bs = _logit_view(pred)
    target = target.contiguous()
    dev = pred.device
    per_channel = torch.empty(L, dtype=torch.float32, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    p = make("mednet_hmloss_params", pred=_ptr(pred), target=_ptr(target), weight=_ptr(weight),
             per_channel=_ptr(per_channel), loss=_ptr(loss), N=n, S=s, batch_stride=bs, L=L, pred_dtype=_dt(pred),
             target_dtype=_dt(target), l1=int(l1))
    ws = _ws(lib().mednet_heatmap_loss_workspace_bytes(_abi.C.byref(p)), dev)
    check(lib().mednet_heatmap_loss_fwd(_abi.C.byref(p), _ptr(ws), ws.numel(), _stream()), "heatmap_loss_fwd")
    _count(2)
    return loss, per_channel, pred, target


def k_hm_bwd(pred, target, weight, grad_out, l1, out=None, out_batch_stride=None):
    pred, n, L, s, bs = _logit_view(pred)
    if out is None:
        out = torch.empty((n, L) + tuple(pred.shape[2:]), dtype=torch.float32, device=pred.device)
        out_batch_stride = L * s
    p = make("mednet_hmloss_bwd_params", pred=_ptr(pred), target=_ptr(target), weight=_ptr(weight),
             grad_out=_ptr(grad_out), dpred=_ptr(out), N=n, S=s, batch_stride=bs, batch_stride_out=out_batch_stride, L=L,
             pred_dtype=_dt(pred), target_dtype=_dt(target), dpred_dtype=_dt(out), l1=int(l1))
    check(lib().mednet_heatmap_loss_bwd(_abi.C.byref(p), _stream()), "heatmap_loss_bwd")
    _count()
    return out


def k_predict_epilogue(logits, num_heatmaps):
    """(N, L+K, *sp) float logits -> (N, L+1, *sp) uint8 (examples/predict.py:88-94)."""
    _need_cuda(logits)
    logits = logits.contiguous()
    n, c = logits.shape[0], logits.shape[1]
    sp = tuple(logits.shape[2:])
    s = logits.numel() // (n * c)
    out = torch.empty((n, num_heatmaps + 1) + sp, dtype=torch.uint8, device=logits.device)
    p = make("mednet_predict_params", logits=_ptr(logits), out=_ptr(out), N=n, S=s, L=num_heatmaps, K=c - num_heatmaps,
             logits_dtype=_dt(logits))
    check(lib().mednet_predict_epilogue(_abi.C.byref(p), _stream()), "predict_epilogue")
    _count()
    return out


def k_final_activation(logits, sigmoid):
    _need_cuda(logits)
    logits = logits.contiguous().float()
    n, c = logits.shape[0], logits.shape[1]
    s = logits.numel() // (n * c)
    out = torch.empty_like(logits)
    check(lib().mednet_final_activation(_ptr(logits), _ptr(out), n, s, c, int(sigmoid), _stream()), "final_activation")
    _count()
    return out


def k_adam(param, grad, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    _need_cuda(param, grad, exp_avg, exp_avg_sq)
    p = make("mednet_adam_params", param=_ptr(param), grad=_ptr(grad), exp_avg=_ptr(exp_avg),
             exp_avg_sq=_ptr(exp_avg_sq), numel=param.numel(), lr=lr, beta1=beta1, beta2=beta2, eps=eps,
             grad_scale=grad_scale, step=step)
    check(lib().mednet_adam_step(_abi.C.byref(p), _stream()), "adam_step")
    _count()


# =============================================================================================
# autograd Functions (NDHWC tensors in, NDHWC tensors out)
# =============================================================================================
def _c(t):
    return t if t.is_contiguous() else t.contiguous()


class Conv3x3Fn(torch.autograd.Function):
    """3x3x3 conv, stride 1, pad 1 with fused bias / addend / activation epilogue.
    ref: midasmednet/unet/components.py:8-9 (+ :35-40 for the fused non-linearity)."""

    @staticmethod
    def forward(ctx, x, weight, bias, addend, act, impl, defer_act=False):
        """defer_act: the activation derivative is applied by the CONSUMERS of y (their `in_act`, see
        include/mednet_b200.h "Deferred activation derivative"); backward then receives d(pre-activation)."""
        x = _c(x)
        cout, cin = weight.shape[0], weight.shape[1]
        sp = tuple(x.shape[1:4])
        impl_id = conv_select_impl(x.shape, sp, cin, cout, x.dtype, 0, impl, x.data_ptr(), 0, 0)
        wp = k_pack_weights(weight, cin, cout, x.dtype, 2 if impl_id == 2 else 0)
        add = _c(addend) if addend is not None else None
        y = k_conv3(x, wp, cout, sp, 0, impl_id, bias=bias.detach().float() if bias is not None else None,
                    addend=add, act=act)
        bwd_act = 0 if defer_act else act
        ctx.weight_ref = weight                      # the Parameter itself (its .grad buffer may be written asynchronously)
        ctx.save_for_backward(x, weight, y if bwd_act else None)
        ctx.act, ctx.impl, ctx.has_bias, ctx.has_addend = bwd_act, impl, bias is not None, addend is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        dy = _c(dy)
        cout, cin = weight.shape[0], weight.shape[1]
        dpre = k_act_bwd(y, dy, ctx.act) if ctx.act else dy
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            sp = tuple(x.shape[1:4])
            impl_id = conv_select_impl(dpre.shape, sp, cout, cin, dpre.dtype, 0, ctx.impl, dpre.data_ptr(), 0, 0)
            wp = k_pack_weights(weight, cin, cout, dpre.dtype, 3 if impl_id == 2 else 1)
            dx = k_conv3(dpre, wp, cin, sp, 0, impl_id)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            wimpl = "simt" if ctx.impl == "simt" else "auto"
            if not ctx.has_bias and _async_wgrad_ok(ctx.weight_ref):
                wgrad_async(dpre, x, 0, wimpl, ctx.weight_ref)       # overlaps the rest of backward; dw stays None
            else:
                dw, db = k_wgrad(dpre, x, 0, wimpl, want_bias=ctx.has_bias)
        return dx, dw, db, (dpre if ctx.has_addend else None), None, None, None


def norm_conv_input_supported(x, weight, bias):
    """GroupNorm -> conv on a tensor that needs no gradient (the image), few input channels, no conv bias."""
    return (not x.requires_grad and bias is None and weight.shape[1] <= 4 and min(x.shape[1:4]) >= 2 and
            weight.shape[0] % 8 == 0)


class NormConvInputFn(torch.autograd.Function):
    """GroupNorm -> 3x3x3 conv (+ activation) applied to a tensor that needs NO gradient: the first layer of the 'gcr'
    networks (components.py:45-57 then :8-9 on the image).  Backward launches neither the convolution's data gradient nor
    the GroupNorm backward: dgamma / dbeta / dW follow from ONE weight gradient against the normalised input and the
    border-class sums of the output gradient (adjoint identity, csrc/input_affine.cu)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, groups, weight, act, impl, defer_act=False):
        x = _c(x)
        g, b = gamma.detach().float(), beta.detach().float()
        xn, _mean, _rstd = k_gn_fwd(x, g, b, groups, 0, None)
        cout, cin = weight.shape[0], weight.shape[1]
        sp = tuple(x.shape[1:4])
        impl_id = conv_select_impl(xn.shape, sp, cin, cout, xn.dtype, 0, impl, xn.data_ptr(), 0, 0)
        wp = k_pack_weights(weight, cin, cout, xn.dtype, 2 if impl_id == 2 else 0)
        y = k_conv3(xn, wp, cout, sp, 0, impl_id, act=act)
        bwd_act = 0 if defer_act else act
        ctx.save_for_backward(x, weight, g, b, y if bwd_act else None)
        ctx.cfg = (groups, bwd_act, impl)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, g, b, y = ctx.saved_tensors
        groups, act, impl = ctx.cfg
        dy = _c(dy)
        dpre = k_act_bwd(y, dy, act) if act else dy
        xhat, _m, _r = k_gn_fwd(x, torch.ones_like(g), torch.zeros_like(b), groups, 0, None)   # recomputed: 1-4 channels
        ghat, _ = k_wgrad(dpre, xhat, 0, "simt" if impl == "simt" else "auto")
        dw, dgamma, dbeta = k_affine_input_grads(weight, ghat, k_border_class_sums(dpre), g, b, dpre.dtype)
        return None, dgamma, dbeta, None, dw, None, None, None


class ConvTranspose3x3Fn(torch.autograd.Function):
    """ConvTranspose3d(k3, s2, p1, op1) + bias, fused with the summation join `x += encoder_features`.
    ref: midasmednet/unet/components.py:259-264, 283-284."""

    @staticmethod
    def forward(ctx, x, weight, bias, skip, impl):
        x = _c(x)
        cin, cout = weight.shape[0], weight.shape[1]
        out_sp = tuple(2 * v for v in x.shape[1:4])
        if skip is not None and tuple(skip.shape[1:4]) != out_sp:
            raise RuntimeError(f"The size of tensor a {out_sp} must match the size of tensor b "
                               f"{tuple(skip.shape[1:4])} (summation join, components.py:284)")
        # tensor cores: 8 output-parity sub-convolutions of 1/2/4/8 taps over the input grid (one launch)
        impl_id = conv_select_impl(x.shape, out_sp, cin, cout, x.dtype, 1, impl, x.data_ptr(), 0, 0)
        wp = k_pack_weights(weight, cin, cout, x.dtype, 4 if impl_id == 2 else 0, transposed=True)
        y = k_conv3(x, wp, cout, out_sp, 1, impl_id, bias=bias.detach().float() if bias is not None else None,
                    addend=_c(skip) if skip is not None else None)
        ctx.save_for_backward(x, weight)
        ctx.has_bias, ctx.has_skip, ctx.impl = bias is not None, skip is not None, impl
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = _c(dy)
        cin, cout = weight.shape[0], weight.shape[1]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            in_sp = tuple(x.shape[1:4])
            impl_id = conv_select_impl(dy.shape, in_sp, cout, cin, dy.dtype, 2, ctx.impl, dy.data_ptr(), 0, 0)
            wp = k_pack_weights(weight, cin, cout, dy.dtype, 5 if impl_id == 2 else 1, transposed=True)
            dx = k_conv3(dy, wp, cin, in_sp, 2, impl_id)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw, db = k_wgrad(x, dy, 2, "simt" if ctx.impl == "simt" else "auto", want_bias=ctx.has_bias)
        return dx, dw, db, (dy if ctx.has_skip else None), None


class GroupNormActFn(torch.autograd.Function):
    """GroupNorm (+ residual add) (+ activation).  ref: components.py:57, :36-40, :177-178."""

    @staticmethod
    def forward(ctx, x, gamma, beta, groups, act, residual, in_act=0):
        x = _c(x)
        res = _c(residual) if residual is not None else None
        g, b = gamma.detach().float(), beta.detach().float()
        y, mean, rstd = k_gn_fwd(x, g, b, groups, act, res)
        ctx.save_for_backward(x, y if act else None, g, mean, rstd)
        ctx.groups, ctx.act, ctx.has_res, ctx.in_act = groups, act, residual is not None, in_act
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, g, mean, rstd = ctx.saved_tensors
        dx, dgamma, dbeta, dres = k_gn_bwd(x, y, _c(dy), g, mean, rstd, ctx.groups, ctx.act, ctx.has_res, ctx.in_act)
        return dx, dgamma, dbeta, None, None, dres, None


class UpcatGroupNormFn(torch.autograd.Function):
    """GroupNorm(cat((skip, nearest_up2(low)), 1)) without materialising the concat.
    ref: components.py:277-280 (interpolate + cat) followed by :57 (the 'g' of the decoder's first 'gcr' layer)."""

    @staticmethod
    def forward(ctx, skip, low, gamma, beta, groups, skip_act, low_act):
        skip, low = _c(skip), _c(low)
        g, b = gamma.detach().float(), beta.detach().float()
        y, mean, rstd = k_upcat_gn_fwd(skip, low, g, b, groups)
        ctx.save_for_backward(skip, low, g, mean, rstd)
        ctx.cfg = (groups, skip_act, low_act)
        return y

    @staticmethod
    def backward(ctx, dy):
        skip, low, g, mean, rstd = ctx.saved_tensors
        groups, skip_act, low_act = ctx.cfg
        dskip, dlow, dgamma, dbeta = k_upcat_gn_bwd(skip, low, _c(dy), g, mean, rstd, groups, skip_act, low_act)
        return dskip, dlow, dgamma, dbeta, None, None, None


# Upsample-aware decoder join (MEDNET_GATHER_UPCONV_*): conv3(cat(skip, nearest_up2(low))) evaluated as
# conv3(skip, W[:, :Cs]) + upconv(low, W[:, Cs:]) where the second term runs on the COARSE grid with 8 summed taps per output
# parity class instead of 27 fine taps -- 2/3 of the decoder's input channels cost 3.4x fewer MMAs in fprop, dgrad and wgrad,
# and the (Cs + Cl)-channel full-resolution tensor and its gradient are never written.
UPCONV = os.environ.get("MEDNET_UPCONV", "1") != "0"          # A/B switch: "0" = materialised-concat path of round 1


def upconv_supported(skip, low, weight):
    if not UPCONV or skip.dtype != torch.bfloat16 or not upcat_gn_supported(skip, low):
        return False
    cs, cl, cout = skip.shape[-1], low.shape[-1], weight.shape[0]
    if weight.shape[1] != cs + cl or cs % 16 or cl % 16 or cout % 16:
        return False
    calibrate_tcgen05()
    lib_ = lib()
    sp, csp = tuple(skip.shape[1:4]), tuple(low.shape[1:4])

    def sel(x_shape, out_sp, k, nout, gather):
        n, di, hi, wi = x_shape[0], x_shape[1], x_shape[2], x_shape[3]
        p = make("mednet_conv3d_params", x=skip.data_ptr(), w=skip.data_ptr(), y=skip.data_ptr(), N=n, Di=di, Hi=hi, Wi=wi,
                 Do=out_sp[0], Ho=out_sp[1], Wo=out_sp[2], K=k, Nout=nout, dtype=_DT[skip.dtype], gather=gather, impl=0)
        return lib_.mednet_conv3d_select_impl(_abi.C.byref(p))
    return (sel(skip.shape, sp, cs, cout, 0) == 2 and sel(low.shape, sp, cl, cout, 3) == 2 and
            sel(skip.shape, sp, cout, cs, 0) == 2 and sel(skip.shape, csp, cout, cl, 4) == 2)


class UpcatGroupNormSplitFn(torch.autograd.Function):
    """GroupNorm(cat((skip, nearest_up2(low)), 1)) returned as (normalised skip, normalised low on the coarse grid).
    ref: components.py:277-280 followed by :57."""

    @staticmethod
    def forward(ctx, skip, low, gamma, beta, groups, skip_act, low_act):
        skip, low = _c(skip), _c(low)
        g, b = gamma.detach().float(), beta.detach().float()
        ys, yl, mean, rstd = k_upcat_gn_split_fwd(skip, low, g, b, groups)
        ctx.save_for_backward(skip, low, g, mean, rstd)
        ctx.cfg = (groups, skip_act, low_act)
        return ys, yl

    @staticmethod
    def backward(ctx, dys, dyl):
        skip, low, g, mean, rstd = ctx.saved_tensors
        groups, skip_act, low_act = ctx.cfg
        dskip, dlow, dgamma, dbeta = k_upcat_gn_split_bwd(skip, low, _c(dys), _c(dyl), g, mean, rstd, groups, skip_act, low_act)
        return dskip, dlow, dgamma, dbeta, None, None, None


class UpConvJoinFn(torch.autograd.Function):
    """act(conv3(xs, W[:, :Cs]) + conv3(nearest_up2(xl), W[:, Cs:])) with xl kept on the coarse grid.
    ref: components.py:277-280 (interpolate + cat) then :8-9 (Conv3d) and :35-40 (fused non-linearity)."""

    @staticmethod
    def forward(ctx, xs, xl, weight, act, defer_act=False):
        xs, xl = _c(xs), _c(xl)
        cs, cl, cout = xs.shape[-1], xl.shape[-1], weight.shape[0]
        sp = tuple(xs.shape[1:4])
        w = weight.detach()
        ws_, wl_ = w[:, :cs].contiguous(), w[:, cs:].contiguous()
        # coarse-grid part first, as an UNROUNDED fp32 partial sum (a bf16 one would move ~1e-3 of the pre-activations across
        # zero: percent-level ReLU' noise); the full-resolution skip part adds it in its epilogue and applies the activation
        y2 = k_conv3(xl, k_pack_weights(wl_, cl, cout, xl.dtype, 6), cout, sp, 3, 2, y_f32=True)
        y = k_conv3(xs, k_pack_weights(ws_, cs, cout, xs.dtype, 2), cout, sp, 0, 2, addend=y2, act=act)
        bwd_act = 0 if defer_act else act
        ctx.weight_ref = weight                      # the Parameter itself (its .grad buffer may be written asynchronously)
        ctx.save_for_backward(xs, xl, weight, y if bwd_act else None)
        ctx.act = bwd_act
        return y

    @staticmethod
    def backward(ctx, dy):
        xs, xl, weight, y = ctx.saved_tensors
        dy = _c(dy)
        dpre = k_act_bwd(y, dy, ctx.act) if ctx.act else dy
        cs, cl, cout = xs.shape[-1], xl.shape[-1], weight.shape[0]
        w = weight.detach()
        ws_, wl_ = w[:, :cs].contiguous(), w[:, cs:].contiguous()
        dxs = dxl = dw = None
        if ctx.needs_input_grad[0]:
            dxs = k_conv3(dpre, k_pack_weights(ws_, cs, cout, dpre.dtype, 3), cs, tuple(xs.shape[1:4]), 0, 2)
        if ctx.needs_input_grad[1]:
            dxl = k_conv3(dpre, k_pack_weights(wl_, cl, cout, dpre.dtype, 7), cl, tuple(xl.shape[1:4]), 4, 2)
        if ctx.needs_input_grad[2]:
            # both halves go straight into their channel range of the ONE (Cout, Cs + Cl, 3,3,3) gradient: the skip channels
            # as (Cout, Cs), the upsampled channels from the per-parity-class passes, computed as (Cl, Cout), transposed
            def both(dst, accumulate):
                k_wgrad_into(dpre, xs, 0, "tcgen05", dst, accumulate, ld=cs + cl, c0=0)
                k_wgrad_into(xl, dpre, 4, "tcgen05", dst, accumulate, ld=cs + cl, c0=cs, transposed=True)
            if _async_wgrad_ok(ctx.weight_ref):
                main, side = torch.cuda.current_stream(dpre.device), side_stream(dpre.device)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    both(ctx.weight_ref.grad, True)
                for t in (dpre, xs, xl):
                    t.record_stream(side)
                _async_pending.add(dpre.device.index)
                if async_grad_listener is not None:
                    async_grad_listener(ctx.weight_ref)
            else:
                dw = torch.empty(weight.shape, dtype=torch.float32, device=weight.device)
                both(dw, False)
        return dxs, dxl, dw, None, None


class ActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act):
        y = k_act_fwd(_c(x), act)
        ctx.save_for_backward(y)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return k_act_bwd(y, _c(dy), ctx.act), None


class MaxPoolFn(torch.autograd.Function):
    """MaxPool3d(2).  ref: components.py:210,224."""

    @staticmethod
    def forward(ctx, x, in_act=0):
        x = _c(x)
        y, idx = k_pool_fwd(x)
        ctx.save_for_backward(idx, y if in_act else None)    # y is kept alive by its consumer anyway
        ctx.in_shape, ctx.in_act = tuple(x.shape), in_act
        return y

    @staticmethod
    def backward(ctx, dy):
        idx, y = ctx.saved_tensors
        return k_pool_bwd(_c(dy), idx, ctx.in_shape, y, ctx.in_act), None


class MaxPoolSkipFn(torch.autograd.Function):
    """MaxPool3d(2) of an encoder output that ALSO feeds a skip connection: returns (pooled, x) and, in backward,
    adds the skip path's gradient to the pool gradient inside the pool-backward kernel (one pass instead of autograd's
    pool-backward + accumulate).  ref: components.py:210,224 and the encoder features kept at model.py:90-95."""

    @staticmethod
    def forward(ctx, x, in_act=0):
        x = _c(x)
        y, idx = k_pool_fwd(x)
        ctx.save_for_backward(idx, y if in_act else None)
        ctx.in_shape, ctx.in_act = tuple(x.shape), in_act
        ctx.set_materialize_grads(False)                      # an unused branch arrives as None, not as a zeros tensor
        return y, x.view_as(x)

    @staticmethod
    def backward(ctx, dy, dskip):
        idx, y = ctx.saved_tensors
        if dy is None:
            return dskip, None
        return k_pool_bwd(_c(dy), idx, ctx.in_shape, y, ctx.in_act, _c(dskip) if dskip is not None else None), None


class UpsampleConcatFn(torch.autograd.Function):
    """F.interpolate(x, size=skip.shape[2:], 'nearest') + cat((skip, x), 1).  ref: components.py:277-280."""

    @staticmethod
    def forward(ctx, skip, low, skip_act=0, low_act=0):
        skip, low = _c(skip), _c(low)
        ctx.shapes = (tuple(skip.shape), tuple(low.shape))
        ctx.acts = (skip_act, low_act)
        ctx.save_for_backward(skip if skip_act else None, low if low_act else None)
        return k_upcat_fwd(skip, low)

    @staticmethod
    def backward(ctx, dout):
        skip, low = ctx.saved_tensors
        dskip, dlow = k_upcat_bwd(_c(dout), *ctx.shapes, skip, low, *ctx.acts)
        return dskip, dlow, None, None


class Conv1x1Fn(torch.autograd.Function):
    """Final 1x1x1 conv with bias: NDHWC activations -> NCDHW fp32 logits.  ref: model.py:77,102."""

    @staticmethod
    def forward(ctx, x, weight, bias, in_act=0):
        x = _c(x)
        w2 = weight.detach().float().reshape(weight.shape[0], -1).contiguous()
        y = k_conv1_fwd(x, w2, bias.detach().float().contiguous())
        ctx.save_for_backward(x, w2)
        ctx.wshape, ctx.in_act = tuple(weight.shape), in_act
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w2 = ctx.saved_tensors
        dx, dw, db = k_conv1_bwd(x, w2, dy.contiguous().float(), need_dx=ctx.needs_input_grad[0], in_act=ctx.in_act)
        return dx, dw.reshape(ctx.wshape), db, None


class ToChannelsLastFn(torch.autograd.Function):
    """(N,C,D,H,W) any float dtype -> (N,D,H,W,C) compute dtype."""

    @staticmethod
    def forward(ctx, x, dtype):
        ctx.src_dtype = x.dtype
        return k_layout(x, True, dtype)

    @staticmethod
    def backward(ctx, dy):
        return k_layout(_c(dy), False, ctx.src_dtype), None


class DiceLossFn(torch.autograd.Function):
    """Fused softmax|sigmoid + weighted soft Dice.  ref: midasmednet/unet/loss.py:114-130."""

    @staticmethod
    def forward(ctx, logits, labels, weight, eps, sigmoid):
        loss, dice, sums, lg, lb = k_dice_fwd(logits, labels, weight, eps, sigmoid)
        ctx.save_for_backward(lg, lb, weight, sums)
        ctx.eps, ctx.sigmoid = eps, sigmoid
        ctx.mark_non_differentiable(dice)
        return loss, dice

    @staticmethod
    def backward(ctx, grad_loss, _grad_dice):
        lg, lb, weight, sums = ctx.saved_tensors
        g = grad_loss.detach().float().contiguous()
        return k_dice_bwd(lg, lb, weight, sums, g, ctx.eps, ctx.sigmoid).to(lg.dtype), None, None, None, None


class CrossEntropyFn(torch.autograd.Function):
    """Weighted-mean cross entropy.  ref: midasmednet/segmentation.py:49, landmarks.py:49."""

    @staticmethod
    def forward(ctx, logits, labels, weight):
        loss, sums, lg, lb = k_ce_fwd(logits, labels, weight)
        ctx.save_for_backward(lg, lb, weight, sums)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        lg, lb, weight, sums = ctx.saved_tensors
        g = grad_loss.detach().float().contiguous()
        return k_ce_bwd(lg, lb, weight, sums, g).to(lg.dtype), None, None


class HeatmapLossFn(torch.autograd.Function):
    """sum_c w_c * mean((o_c - h_c)^2 | |o_c - h_c|).  ref: midasmednet/landmarks.py:125-134."""

    @staticmethod
    def forward(ctx, pred, target, weight, l1):
        loss, per_channel, pr, tg = k_hm_fwd(pred, target, weight, l1)
        ctx.save_for_backward(pr, tg, weight)
        ctx.l1 = l1
        ctx.mark_non_differentiable(per_channel)
        return loss, per_channel

    @staticmethod
    def backward(ctx, grad_loss, _g):
        pr, tg, weight = ctx.saved_tensors
        g = grad_loss.detach().float().contiguous()
        return k_hm_bwd(pr, tg, weight, g, ctx.l1).to(pr.dtype), None, None, None


class LandmarkLossFn(torch.autograd.Function):
    """The whole LandmarkNet.loss on the un-split network output: heatmap regression on channels [0,L),
    Dice|CE on channels [L, L+K); one gradient tensor, no slice/pad copies.
    ref: midasmednet/landmarks.py:74-75 (split), :125-134 (loss)."""

    @staticmethod
    def forward(ctx, outputs, labels, heatmaps, class_weight, reg_weight, use_ce, l1, eps):
        L = heatmaps.shape[1]
        outputs = outputs.contiguous()
        hm_loss, per_channel, _, tg = k_hm_fwd(outputs[:, :L], heatmaps, reg_weight, l1)
        if use_ce:
            cls_loss, sums, _, lb = k_ce_fwd(outputs[:, L:], labels, class_weight)
        else:
            cls_loss, _dice, sums, _, lb = k_dice_fwd(outputs[:, L:], labels, class_weight, eps, False)
        ctx.save_for_backward(outputs, lb, tg, class_weight, reg_weight, sums)
        ctx.cfg = (L, use_ce, l1, eps)
        total = hm_loss + cls_loss       # scalar add on device (torch op on two 0-d tensors: plumbing, not hot path)
        return total, cls_loss, hm_loss

    @staticmethod
    def backward(ctx, g_total, g_cls, g_hm):
        outputs, lb, tg, class_weight, reg_weight, sums = ctx.saved_tensors
        L, use_ce, l1, eps = ctx.cfg
        n, c = outputs.shape[0], outputs.shape[1]
        s = outputs.numel() // (n * c)
        zero = torch.zeros((), dtype=torch.float32, device=outputs.device)
        gc = ((g_total if g_total is not None else zero) + (g_cls if g_cls is not None else zero)).float().contiguous()
        gh = ((g_total if g_total is not None else zero) + (g_hm if g_hm is not None else zero)).float().contiguous()
        d = torch.empty(outputs.shape, dtype=torch.float32, device=outputs.device)
        k_hm_bwd(outputs[:, :L], tg, reg_weight, gh, l1, out=d[:, :L], out_batch_stride=c * s)
        if use_ce:
            k_ce_bwd(outputs[:, L:], lb, class_weight, sums, gc, out=d[:, L:], out_batch_stride=c * s)
        else:
            k_dice_bwd(outputs[:, L:], lb, class_weight, sums, gc, eps, False, out=d[:, L:], out_batch_stride=c * s)
        return d.to(outputs.dtype), None, None, None, None, None, None, None

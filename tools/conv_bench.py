#!/usr/bin/env python
"""Per-layer micro-benchmark of the 3x3x3 convolution kernels (fprop / dgrad / wgrad) on the UNet3D layer shapes.
GPU only.  Prints one JSON line per (layer, pass): algorithmic TFLOP/s from CUDA events, median of `reps`."""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-mednet_b200"))
from mednet_b200 import ops  # noqa: E402

# (Cin, Cout, spatial divisor) of UNet3D f=64
LAYERS = [(32, 64, 1), (64, 64, 2), (64, 128, 2), (128, 128, 4), (128, 256, 4), (256, 256, 8), (256, 512, 8),
          (768, 256, 4), (256, 256, 4), (384, 128, 2), (128, 128, 2), (192, 64, 1), (64, 64, 1)]


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--edge", type=int, default=128)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--passes", default="fprop,wgrad")
    ap.add_argument("--wgrad-impl", default="auto")
    ap.add_argument("--dual-issue", type=int, default=1)
    ap.add_argument("--wt-fastest", type=int, default=1)
    ap.add_argument("--pair-planes", type=int, default=1)
    ap.add_argument("--d-fastest", type=int, default=1)
    ap.add_argument("--kd-merge", type=int, default=1)
    ap.add_argument("--ntile-max", type=int, default=128)
    ap.add_argument("--wgrad-dual", type=int, default=1)
    ap.add_argument("--layers", default="", help="comma-separated indices into LAYERS (default: all)")
    a = ap.parse_args()
    dev = "cuda"
    ops.calibrate_tcgen05()
    from mednet_b200._abi import check, lib
    check(lib().mednet_tcgen05_set_option(b"dual_issue", a.dual_issue), "set_option")
    check(lib().mednet_tcgen05_set_option(b"wgrad_wt_fastest", a.wt_fastest), "set_option")
    check(lib().mednet_tcgen05_set_option(b"wgrad_pair_planes", a.pair_planes), "set_option")
    check(lib().mednet_tcgen05_set_option(b"wgrad_d_fastest", a.d_fastest), "set_option")
    check(lib().mednet_tcgen05_set_option(b"kd_merge", a.kd_merge), "set_option")
    check(lib().mednet_tcgen05_set_option(b"ntile_max", a.ntile_max), "set_option")
    check(lib().mednet_tcgen05_set_option(b"wgrad_dual_issue", a.wgrad_dual), "set_option")
    layers = [LAYERS[int(i)] for i in a.layers.split(",")] if a.layers else LAYERS
    for cin, cout, div in layers:
        s = a.edge // div
        x = torch.randn(a.batch, s, s, s, cin, device=dev).to(torch.bfloat16)
        dy = torch.randn(a.batch, s, s, s, cout, device=dev).to(torch.bfloat16)
        w = torch.randn(cout, cin, 3, 3, 3, device=dev) * 0.05
        flops = 2.0 * a.batch * s ** 3 * cin * cout * 27
        if "fprop" in a.passes:
            impl = ops.conv_select_impl(x.shape, (s, s, s), cin, cout, x.dtype, 0, "auto", x.data_ptr())
            wp = ops.k_pack_weights(w, cin, cout, x.dtype, 2 if impl == 2 else 0)
            ms = timed(lambda: ops.k_conv3(x, wp, cout, (s, s, s), 0, impl), a.reps)
            print(json.dumps(dict(layer=f"{cin}->{cout}@{s}^3", op="fprop", impl=impl, ms=ms, tflops=flops / ms / 1e9,
                                  dual_issue=a.dual_issue, kd_merge=a.kd_merge)))
        if "dgrad" in a.passes:
            impl = ops.conv_select_impl(dy.shape, (s, s, s), cout, cin, dy.dtype, 0, "auto", dy.data_ptr())
            wp = ops.k_pack_weights(w, cin, cout, dy.dtype, 3 if impl == 2 else 1)
            ms = timed(lambda: ops.k_conv3(dy, wp, cin, (s, s, s), 0, impl), a.reps)
            print(json.dumps(dict(layer=f"{cin}->{cout}@{s}^3", op="dgrad", impl=impl, ms=ms, tflops=flops / ms / 1e9,
                                  dual_issue=a.dual_issue, kd_merge=a.kd_merge)))
        if "wgrad" in a.passes:
            ms = timed(lambda: ops.k_wgrad(dy, x, 0, a.wgrad_impl), a.reps)
            print(json.dumps(dict(layer=f"{cin}->{cout}@{s}^3", op="wgrad", impl=a.wgrad_impl, ms=ms, tflops=flops / ms / 1e9,
                                  wt_fastest=a.wt_fastest, pair_planes=a.pair_planes)))
        del x, dy


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Where the cycles of conv3_tc_kernel go: runs the profiling instantiation (mednet_tcgen05_set_option("conv_profile", 1))
on UNet3D layer shapes and prints, per layer, the mean per-CTA share of cycles each warp role spends WAITING:
MMA issuer (for halo planes / weight stages / a free accumulator), epilogue (for a finished accumulator), the two TMA
producers (for free slots).  GPU only."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-mednet_b200"))
from mednet_b200 import _abi, ops  # noqa: E402
from mednet_b200._abi import check, lib, make  # noqa: E402

LAYERS = [(32, 64, 1), (64, 64, 2), (64, 128, 2), (128, 128, 4), (128, 256, 4), (768, 256, 4), (256, 256, 4), (384, 128, 2),
          (128, 128, 2), (192, 64, 1), (64, 64, 1), (64, 192, 1), (64, 32, 1)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--edge", type=int, default=128)
    ap.add_argument("--kd-merge", type=int, default=1)
    ap.add_argument("--act", type=int, default=1)
    a = ap.parse_args()
    ops.calibrate_tcgen05()
    check(lib().mednet_tcgen05_set_option(b"kd_merge", a.kd_merge), "set_option")
    nsm = lib().mednet_sm_count()
    for cin, cout, div in LAYERS:
        s = a.edge // div
        x = torch.randn(a.batch, s, s, s, cin, device="cuda").to(torch.bfloat16)
        w = torch.randn(cout, cin, 3, 3, 3, device="cuda") * 0.05
        wp = ops.k_pack_weights(w, cin, cout, x.dtype, 2)
        y = torch.empty((a.batch, s, s, s, cout), dtype=x.dtype, device="cuda")
        ws = torch.zeros(nsm * 8, dtype=torch.int64, device="cuda")
        p = make("mednet_conv3d_params", x=x.data_ptr(), w=wp.data_ptr(), y=y.data_ptr(), N=a.batch, Di=s, Hi=s, Wi=s, Do=s,
                 Ho=s, Wo=s, K=cin, Nout=cout, dtype=1, act=a.act, act_param=0.0, gather=0, impl=2)
        check(lib().mednet_tcgen05_set_option(b"conv_profile", 0), "set_option")
        for _ in range(2):
            check(lib().mednet_conv3d_fprop(_abi.C.byref(p), ws.data_ptr(), ws.numel() * 8, ops._stream()), "fprop")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib().mednet_conv3d_fprop(_abi.C.byref(p), ws.data_ptr(), ws.numel() * 8, ops._stream()), "fprop")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        check(lib().mednet_tcgen05_set_option(b"conv_profile", 1), "set_option")
        check(lib().mednet_conv3d_fprop(_abi.C.byref(p), ws.data_ptr(), ws.numel() * 8, ops._stream()), "fprop")
        torch.cuda.synchronize()
        c = ws.view(nsm, 8).double()
        c = c[c[:, 0] > 0]
        tot, epi = c[:, 0].mean().item(), c[:, 4].mean().item()
        f = lambda i, d: round((c[:, i].mean().item() / d), 3)
        flops = 2.0 * a.batch * s ** 3 * cin * cout * 27
        print(json.dumps(dict(layer=f"{cin}->{cout}@{s}^3", kd_merge=a.kd_merge, ms=round(ms, 3), tflops=round(flops / ms / 1e9, 1),
                              mma_cycles=int(tot), mma_wait_halo=f(1, tot), mma_wait_weights=f(2, tot), mma_wait_acc=f(3, tot),
                              epi_wait_acc=f(5, epi), halo_prod_wait=f(6, tot), w_prod_wait=f(7, tot),
                              clk_ghz=round(tot / (ms * 1e6), 3))))
    check(lib().mednet_tcgen05_set_option(b"conv_profile", 0), "set_option")
    # ---- weight gradient: share of the MMA issuer's cycles spent waiting for a TMA-filled stage, per layer
    for cin, cout, div in LAYERS:
        if cin == 64 and cout in (192, 32):
            continue
        s = a.edge // div
        x = torch.randn(a.batch, s, s, s, cin, device="cuda").to(torch.bfloat16)
        dy = torch.randn(a.batch, s, s, s, cout, device="cuda").to(torch.bfloat16)
        dw = torch.empty((cout, cin, 3, 3, 3), dtype=torch.float32, device="cuda")
        p = ops.wgrad_params(dy, x, 0, "tcgen05", dw, None)
        check(lib().mednet_tcgen05_set_option(b"wgrad_profile", 1), "set_option")
        nbytes = lib().mednet_conv3d_wgrad_workspace_bytes(_abi.C.byref(p))
        ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
        check(lib().mednet_tcgen05_set_option(b"wgrad_profile", 0), "set_option")
        for _ in range(2):
            check(lib().mednet_conv3d_wgrad(_abi.C.byref(p), ws.data_ptr(), nbytes, ops._stream()), "wgrad")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib().mednet_conv3d_wgrad(_abi.C.byref(p), ws.data_ptr(), nbytes, ops._stream()), "wgrad")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        check(lib().mednet_tcgen05_set_option(b"wgrad_profile", 1), "set_option")
        ws[nbytes - 65536:].zero_()
        check(lib().mednet_conv3d_wgrad(_abi.C.byref(p), ws.data_ptr(), nbytes, ops._stream()), "wgrad")
        torch.cuda.synchronize()
        check(lib().mednet_tcgen05_set_option(b"wgrad_profile", 0), "set_option")
        c = ws[nbytes - 65536:].view(torch.int64).view(2, 1024, 4).double()
        flops = 2.0 * a.batch * s ** 3 * cin * cout * 27
        row = dict(layer=f"{cin}->{cout}@{s}^3", op="wgrad", ms=round(ms, 3), tflops=round(flops / ms / 1e9, 1))
        for seg in range(2):
            cs = c[seg][c[seg][:, 0] > 0]
            if len(cs):
                row[f"seg{seg}"] = dict(ctas=len(cs), mma_cycles=int(cs[:, 0].mean()), max_cycles=int(cs[:, 0].max()),
                                        wait_stage=round((cs[:, 1].mean() / cs[:, 0].mean()).item(), 3),
                                        producer_wait=round((cs[:, 2].mean() / cs[:, 0].mean()).item(), 3),
                                        bricks=cs[:, 3].mean().item(),
                                        clk_per_brick=int((cs[:, 0].sum() / cs[:, 3].sum()).item()))
        print(json.dumps(row))


if __name__ == "__main__":
    main()

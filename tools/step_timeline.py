#!/usr/bin/env python
"""In-step timeline of one training step: every C-ABI launch is bracketed by CUDA events on the stream it is issued on
(main stream, or the side stream of the asynchronous weight gradients), so kernel durations are measured under the
clocks and contention of the real step (ncu's launch list is serialised, cold-cache and runs at other clocks).
Prints, per entry point, launches / total ms / share of the step, the busy time and the idle gaps of each stream.
GPU only:   python tools/step_timeline.py [--workload cfg3]"""
import argparse
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "torch-mednet_b200"))
import bench  # noqa: E402
from mednet_b200 import _abi, ops  # noqa: E402


class Recorder:
    def __init__(self, handle):
        self._h, self.rows, self.on = handle, [], False

    def __getattr__(self, name):
        fn = getattr(self._h, name)
        if not name.startswith("mednet_") or "workspace_bytes" in name or "select_impl" in name or name in (
                "mednet_abi_version", "mednet_error_string", "mednet_sm_count", "mednet_device_has_tcgen05",
                "mednet_tcgen05_set_option", "mednet_tcgen05_configure"):
            return fn

        def wrapped(*a):
            if not self.on:
                return fn(*a)
            st = torch.cuda.current_stream()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            r = fn(*a)
            e1.record(st)
            self.rows.append((name, st.cuda_stream, e0, e1))
            return r
        return wrapped


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--dump", default="", help="write every (start ms, end ms, stream, entry point) row of the timed steps here")
    ap.add_argument("--sync-wgrad", action="store_true", help="weight gradients on the main stream (no side stream)")
    a = ap.parse_args()
    wl = bench.WORKLOADS[a.workload]
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    from mednet_b200.landmarks import LandmarkNet, LandmarkUNet3D
    from mednet_b200.segmentation import SegmentationNet, SegmentationUNet3D
    hp = bench.hparams_for(wl)
    cls = (LandmarkNet if wl["heatmaps"] else SegmentationNet) if wl["arch"] == "residual" else \
        (LandmarkUNet3D if wl["heatmaps"] else SegmentationUNet3D)
    model = cls(hp).to(dev)
    opt = model.configure_optimizers()
    if a.sync_wgrad:
        opt.async_wgrad = False
    opt.zero_grad()
    batch = bench.synthetic_batch(wl, 1000, dev)
    rec = Recorder(_abi.lib())
    _abi._lib = rec

    def step():
        out = model.training_step(batch, 0)
        out["loss"].backward()
        opt.step()
        opt.zero_grad()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    rec.on = True
    t0.record()
    for _ in range(a.steps):
        step()
    t1.record()
    torch.cuda.synchronize()
    rec.on = False
    total = t0.elapsed_time(t1) / a.steps
    agg = collections.defaultdict(lambda: [0, 0.0])
    streams = collections.defaultdict(list)
    for name, sid, e0, e1 in rec.rows:
        d = e0.elapsed_time(e1)
        agg[name][0] += 1
        agg[name][1] += d
        streams[sid].append((t0.elapsed_time(e0), t0.elapsed_time(e1), name))
    if a.dump:
        with open(a.dump, "w") as f:
            for name, sid, e0, e1 in rec.rows:
                f.write(f"{t0.elapsed_time(e0):.4f} {t0.elapsed_time(e1):.4f} {sid:#x} {name}\n")
    print(f"# {a.workload}: {total:.2f} ms/step with event recording ({len(rec.rows) // a.steps} ABI calls per step)")
    print(f"{'ms/step':>9} {'share':>6} {'calls':>6}  entry point")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t / a.steps:9.3f} {100 * t / a.steps / total:5.1f}% {n // a.steps:6d}  {k}")
    for sid, rows in streams.items():
        rows.sort()
        busy = sum(e - s for s, e, _ in rows) / a.steps
        gaps = [(rows[i + 1][0] - rows[i][1], rows[i][2], rows[i + 1][2]) for i in range(len(rows) - 1)]
        idle = sum(g for g, _, _ in gaps if g > 0) / a.steps
        print(f"# stream {sid:#x}: busy {busy:.2f} ms/step, idle between its launches {idle:.2f} ms/step")
        big = collections.defaultdict(float)
        for g, a_, b_ in gaps:
            if g > 0.02:
                big[(a_, b_)] += g / a.steps
        for (a_, b_), g in sorted(big.items(), key=lambda x: -x[1])[:12]:
            print(f"#     gap {g:7.3f} ms/step  after {a_} before {b_}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the UNet3D training hot path (BASELINE.json metric: UNet3D train voxels/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|cfg5|cfg1] [--impl reference]

One step = one pass of the hot path over one synthetic batch: forward -> loss -> backward ->
(bucketed all-reduce when N > 1) -> fused Adam.  Prints ONE JSON line (rank 0).

  value      whole-job voxels/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e        same metric through the public training_step API with HOST (pinned) batches: the H2D copy of
             every step's inputs and the D2H read of the loss are inside the timed region
  roofline   dominant kernel (tcgen05 implicit-GEMM conv): the FLOPs its launches execute divided by their
             summed CUDA-event durations, against the measured sustained dense bf16 peak
             (`reference_equivalent`: the same with the FLOPs of the reference's ops those launches replace)
  cpu_baseline  the reference's own modules (oracle/_ref, kind "reference"; the oracle port when absent)
             timed on the host cores on a bounded sample: batch 1 on the workload's own patch edge
  clocks     SM clock / throttle reasons sampled through NVML during the timed region
`--impl reference` times that CPU path alone with all host threads (the reference ships no GPU kernels of
its own: every FLOP of it runs inside stock PyTorch).
The host is kept at most one step ahead of the device and the timed region is rehearsed once, untimed
(`config.host_run_ahead`; DESIGN.md section 6 "Measurement hygiene").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "torch-mednet_b200"))

import torch  # noqa: E402

WORKLOADS = {
    # name: (arch, f_maps, classes, heatmaps, batch per GPU, patch edge, loss)
    "cfg1": dict(arch="unet3d", f_maps=64, classes=2, heatmaps=0, batch=2, edge=64, loss="DICE",
                 desc="UNet3D(1,2) f=64 4-level, batch 2, 64^3, Dice (examples/train_seg.py shape)"),
    "cfg2": dict(arch="unet3d", f_maps=64, classes=2, heatmaps=8, batch=4, edge=96, loss="DICE",
                 desc="UNet3D landmark heatmap regression, 8 heatmaps + 2 classes, batch 4, 96^3"),
    "cfg3": dict(arch="unet3d", f_maps=64, classes=4, heatmaps=0, batch=8, edge=128, loss="DICE",
                 desc="UNet3D(1,4) f=64 4-level multi-class segmentation, batch 8 per GPU, 128^3, Dice"),
    "cfg5": dict(arch="unet3d", f_maps=[32, 64, 128, 256, 512], classes=2, heatmaps=0, batch=2, edge=160, loss="DICE",
                 desc="wide UNet3D f=[32..512] 5-level, batch 2 per GPU, 160^3, Dice"),
    "tiny": dict(arch="unet3d", f_maps=[16, 32, 64], classes=2, heatmaps=0, batch=2, edge=32, loss="DICE",
                 desc="smoke-sized UNet3D"),
    # the network the reference's task modules actually derive from (segmentation.py:22): not the headline metric
    "res32": dict(arch="residual", f_maps=32, classes=2, heatmaps=0, batch=2, edge=128, loss="DICE",
                  desc="ResidualUNet3D(1,2) f=32 5-level (SegmentationNet as shipped), batch 2, 128^3, Dice"),
    # sliding-window inference (examples/predict.py): not the headline metric, measured with --workload cfg4
    "cfg4": dict(arch="unet3d", f_maps=64, classes=2, heatmaps=0, batch=4, edge=128, loss="DICE", predict=True,
                 volume=(512, 512, 400), overlap=16,
                 desc="sliding-window inference, 512x512x400 volume, 128^3 tiles, overlap 16 (180 tiles), tile batch 4"),
    # further variants SURVEY.md section 8(d) lists for the same configurations (same code paths, other parameters)
    "cfg3_ce": dict(arch="unet3d", f_maps=64, classes=4, heatmaps=0, batch=8, edge=128, loss="CE",
                    desc="cfg-3 with the weighted cross-entropy loss instead of Dice"),
    "cfg2_res": dict(arch="residual", f_maps=64, classes=2, heatmaps=8, batch=4, edge=96, loss="DICE",
                     desc="cfg-2 on LandmarkNet as shipped: ResidualUNet3D f=64 5-level, 8 heatmaps + 2 classes, batch 4, 96^3"),
    "cfg5_res": dict(arch="residual", f_maps=32, classes=2, heatmaps=0, batch=2, edge=160, loss="DICE",
                     desc="cfg-5 on ResidualUNet3D(1,2) f=32 5-level, batch 2 per GPU, 160^3, Dice"),
    "cfg4_o32": dict(arch="unet3d", f_maps=64, classes=2, heatmaps=0, batch=4, edge=128, loss="DICE", predict=True,
                     volume=(512, 512, 400), overlap=32,
                     desc="sliding-window inference, 512x512x400 volume, 128^3 tiles, overlap 32 (448 tiles), tile batch 4"),
    # the caller side of the training path (MedDataset.__getitem__ + augmentation): not the headline metric
    "sampler": dict(sampler=True, subjects=4, volume=(320, 320, 256), batch=8, edge=128, probs=[0.3, 0.7],
                    desc="random patch sampling from 4 HBM-resident subjects 320x320x256, batch 8 x 128^3 x 1 ch, "
                         "class_probabilities [0.3, 0.7]"),
}


def synthetic_batch(wl, seed, device, pin=False):
    """MedDataset contract (dataset.py:332-346): data (B,1,S,S,S) f32; label (B,L+1,S,S,S) u8, class map last."""
    g = torch.Generator().manual_seed(seed)
    b, s = wl["batch"], wl["edge"]
    data = torch.randn((b, 1, s, s, s), generator=g)
    cls = torch.randint(0, wl["classes"], (b, 1, s, s, s), generator=g, dtype=torch.uint8)
    if wl["heatmaps"]:
        hm = torch.zeros((b, wl["heatmaps"], s, s, s), dtype=torch.uint8)
        ax = torch.arange(s, dtype=torch.float32)
        for n in range(b):
            pts = torch.rand((wl["heatmaps"], 3), generator=g) * s
            for l in range(wl["heatmaps"]):
                r2 = ((ax[:, None, None] - pts[l, 0]) ** 2 + (ax[None, :, None] - pts[l, 1]) ** 2 +
                      (ax[None, None, :] - pts[l, 2]) ** 2)
                hm[n, l] = (255.0 * torch.exp(-r2 / 18.0)).to(torch.uint8)       # sigma = 3 voxels
        label = torch.cat([hm, cls], dim=1)
    else:
        label = cls
    batch = {"data": data, "label": label}
    if pin:
        return {k: v.pin_memory() for k, v in batch.items()}
    return {k: v.to(device) for k, v in batch.items()}


def hparams_for(wl):
    ns = argparse.Namespace(in_channels=1, fmaps=wl["f_maps"], learning_rate=1e-3, num_workers=0, batch_size=wl["batch"])
    if wl["heatmaps"]:
        ns.out_channels = wl["heatmaps"] + wl["classes"]
        ns.loss_class, ns.loss_class_weight = wl["loss"], [0.05] + [1.0] * (wl["classes"] - 1)
        ns.loss_regression = "L2"
        ns.loss_regression_weight = ([0.001, 0.015, 0.015, 0.015] + [0.001] * 8)[:wl["heatmaps"]]
    else:
        ns.out_channels = wl["classes"]
        ns.loss, ns.loss_weight = wl["loss"], [0.05] + [1.0] * (wl["classes"] - 1)
    return ns


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_oracle_voxels_per_s(wl, steps, warmup, edge=None, batch=None, forward_only=False):
    """Times the reference's CPU path (forward + loss + backward + Adam, or forward only for the predict workloads) on the
    host cores; returns (voxels/s, info).  kind "reference": the reference's own modules (oracle/_ref, vendored unmodified
    by oracle/build_ref.py) driven by the step of segmentation.py:58-65 / landmarks.py:66-83; kind "port": the oracle
    restatement when oracle/_ref is absent."""
    import time
    from oracle import build_ref
    from oracle import loss as oloss
    from oracle import steps as osteps
    from oracle import unet as ounet
    torch.set_num_threads(os.cpu_count() or 1)
    edge = edge or wl["edge"]
    batch = batch or wl["batch"]
    sub = dict(wl, edge=edge, batch=batch)
    out_ch = wl["classes"] + wl["heatmaps"]
    kind = "residual" if wl["arch"] == "residual" else "unet3d"
    b = synthetic_batch(sub, 0, "cpu")
    hp = hparams_for(wl)
    ref = build_ref.load()
    if ref is not None:
        rmodel, rloss = ref
        torch.manual_seed(0)
        net = (rmodel.ResidualUNet3D if kind == "residual" else rmodel.UNet3D)(1, out_ch, False, f_maps=wl["f_maps"])
        opt = torch.optim.Adam(net.parameters(), lr=1e-3)                     # segmentation.py:119-120
        if wl["heatmaps"]:
            cw = torch.tensor(hp.loss_class_weight)
            rw = list(hp.loss_regression_weight)
            L = wl["heatmaps"]
            class_loss = rloss.DiceLoss(weight=cw) if hp.loss_class == "DICE" else torch.nn.CrossEntropyLoss(weight=cw)
        else:
            w = torch.tensor(hp.loss_weight)
            crit = rloss.DiceLoss(weight=w) if hp.loss == "DICE" else torch.nn.CrossEntropyLoss(weight=w)   # segmentation.py:43-49
        times = []
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            inputs = b["data"].float()
            if forward_only:
                with torch.no_grad():
                    net(inputs)
            else:
                opt.zero_grad()
                out = net(inputs)
                labels = b["label"][:, -1].long()
                if wl["heatmaps"]:                                             # landmarks.py:66-83, 125-134
                    hm = b["label"][:, :-1].float()
                    value = class_loss(out[:, L:], labels)
                    for c in range(L):
                        value = value + rw[c] * torch.nn.functional.mse_loss(out[:, c], hm[:, c])
                else:
                    value = crit(out, labels)
                value.backward()
                opt.step()
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        impl = "the reference's own midasmednet.unet modules (oracle/_ref)"
        k = "reference"
    else:
        sd = (ounet.make_residual_unet3d_state_dict if kind == "residual" else ounet.make_unet3d_state_dict)(1, out_ch, wl["f_maps"])
        kw = dict(f_maps=wl["f_maps"])
        if forward_only:
            times = []
            fwd = ounet.residual_unet3d_forward if kind == "residual" else ounet.unet3d_forward
            for it in range(warmup + steps):
                t0 = time.perf_counter()
                with torch.no_grad():
                    fwd(sd, b["data"].float(), **kw)
                if it >= warmup:
                    times.append(time.perf_counter() - t0)
        elif wl["heatmaps"]:
            times, _ = osteps.time_training_steps(kind, sd, b, steps=steps, warmup=warmup, task="ldmk",
                                                  loss_class=hp.loss_class, loss_class_weight=hp.loss_class_weight,
                                                  loss_regression_weight=hp.loss_regression_weight, **kw)
        else:
            times, _ = osteps.time_training_steps(kind, sd, b, steps=steps, warmup=warmup, task="seg", loss=hp.loss,
                                                  loss_weight=hp.loss_weight, **kw)
        impl = "oracle port of the reference"
        k = "port"
    t = sum(times) / len(times)
    vox = batch * edge ** 3
    what = "forward" if forward_only else "fwd+loss+bwd+Adam"
    return vox / t, dict(seconds_per_step=t, kind=k, edge=edge, batch=batch,
                         sample=f"batch {batch} x {edge}^3 patch, {len(times)} timed step(s) after {warmup} warm-up, {what}, "
                                f"{impl} (plain PyTorch fp32, {torch.get_num_threads()} threads)")


def run_reference_arm(args, wl):
    """The reference's CPU implementation of the path on this box's host cores, on the workload's own patch edge with a
    batch of one (reduced N, voxels/s normalisation) so that K steps end within minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    predict = bool(wl.get("predict"))
    v, info = cpu_oracle_voxels_per_s(wl, steps=max(1, args.steps), warmup=max(0, min(args.warmup, 1)), edge=wl["edge"],
                                      batch=1, forward_only=predict)
    metric = "UNet3D train voxels/s"
    if predict:            # useful voxels of a tile = its centre crop (dataset.py:452-474)
        v *= ((wl["edge"] - 2 * wl["overlap"]) / wl["edge"]) ** 3
        metric = "UNet3D predict voxels/s"
    line = {"impl": "reference", "metric": metric, "value": v, "unit": "voxels/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": info["seconds_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "strong" if predict else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": args.workload + ": " + wl["desc"], "per_step_sample": f"batch 1 x {wl['edge']}^3 (same patch "
                       "edge and network as the GPU arm, batch reduced to 1; voxels/s normalised)",
                       "l2": "inputs larger than L2"},
            "cpu_baseline": {"value": v, "unit": "voxels/s", "cores": torch.get_num_threads(), "kind": info["kind"],
                             "sample": info["sample"]},
            "e2e": {"value": v, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (every 100 ms).  NVML is queried from a thread of this
    process: `prepare()` (NVML init, device handle) runs before the warm-up, so nothing is started inside the timed region
    -- spawning `nvidia-smi -lms` there was measured to stretch a 0.4 s region by up to 27 % while the tool initialised.
    Falls back to an `nvidia-smi` loop started in prepare() (i.e. before the warm-up) when pynvml is unavailable."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.nvml, self.handle, self.proc = index, None, None, None
        self.samples, self.thread, self.running = [], None, False
        self.smi_skip = 0

    def prepare(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: map through CUDA_VISIBLE_DEVICES when it lists plain indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip().isdigit()]
            phys = int(ids[self.index]) if len(ids) > self.index else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            # first use of the two queries of _loop here, outside any timed region (driver-side lazy setup)
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        except Exception:
            self.nvml = None
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                              "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                             stderr=subprocess.DEVNULL, text=True)
            except OSError:
                self.proc = None

    def _loop(self):
        n = self.nvml
        while self.running:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                bits = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((mhz, bits))
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        if self.nvml is not None:
            import threading
            self.running = True
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        else:
            self.t_start = time.time()

    def stop(self):
        if self.nvml is not None:
            self.running = False
            self.thread.join(timeout=2)
            n = self.nvml
            names = {"hw_slowdown": n.nvmlClocksThrottleReasonHwSlowdown,
                     "hw_thermal_slowdown": n.nvmlClocksThrottleReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": n.nvmlClocksThrottleReasonSwThermalSlowdown,
                     "sw_power_cap": n.nvmlClocksThrottleReasonSwPowerCap}
            sm = [m for m, _ in self.samples]
            reasons = sorted(k for k, bit in names.items() if any(b & bit for _, b in self.samples))
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "source": "NVML, 100 ms, timed region only"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        dur = time.time() - self.t_start
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out = ""
        lines = out.strip().splitlines()
        lines = lines[-max(1, int(dur / 0.2) + 1):]              # the samples of the timed region (200 ms apart)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, val in zip(names, f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 200 started before the warm-up"}


# ------------------------------------------------------------------------------------------------ predict arm
def run_predict(args, wl):
    """cfg-4: tiles sharded round-robin over the ranks, no data-path collective; one "step" = the whole volume.
    value = USEFUL output voxels / s (the BASELINE.json 'predict voxels/s'); computed voxels/s reported beside it."""
    import numpy as np
    import torch.distributed as dist
    from mednet_b200.parallel import init_distributed
    from mednet_b200.predict import SlidingWindowPredictor
    from mednet_b200.segmentation import SegmentationUNet3D
    rank, local, world = init_distributed()
    dev = torch.device("cuda", local)
    torch.manual_seed(0)
    model = SegmentationUNet3D(hparams_for(wl)).to(dev)
    model.freeze()
    X, Y, Z = wl["volume"]
    vol_host = torch.randn((1, X, Y, Z), generator=torch.Generator().manual_seed(7)).to(torch.bfloat16).pin_memory()
    vol_dev = vol_host.to(dev)
    pred = SlidingWindowPredictor(model, [wl["edge"]] * 3, [wl["overlap"]] * 3, wl["heatmaps"], wl["batch"], rank, world)
    tiles = len(pred.tile_origins((X, Y, Z)))

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    from mednet_b200 import ops
    steps, warm = max(1, min(args.steps, 3)), 1
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.prepare()
    for _ in range(warm):
        pred(vol_dev)
    if rank == 0:
        sampler.start()
    ops.conv_events = []
    launches0 = ops.launch_count
    ms = timed(lambda: pred(vol_dev), steps)
    launches = ops.launch_count - launches0
    conv_events, ops.conv_events = ops.conv_events, None
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(lambda: pred(vol_host).cpu(), steps)        # host volume in, stitched uint8 volume back on the host
    if rank == 0:
        useful = X * Y * Z
        line = {"metric": "UNet3D predict voxels/s", "value": useful * steps / (ms * 1e-3), "unit": "voxels/s",
                "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": args.workload + ": " + wl["desc"], "tiles": tiles,
                           "computed_voxels_per_s": tiles * wl["edge"] ** 3 * steps / (ms * 1e-3),
                           "sharding": f"tiles round-robin over {world} rank(s), no data-path collective; disjoint uint8 "
                                       "regions combined once per volume (all_gather of per-rank tile outputs)",
                           "l2": "inputs larger than L2"},
                "clocks": clocks,
                "e2e": {"value": useful * steps / (ms_e2e * 1e-3), "unit": "voxels/s",
                        "h2d_bytes_per_step": vol_host.numel() * 2, "d2h_bytes_per_step": (wl["heatmaps"] + 1) * useful,
                        "ms_per_step": ms_e2e / steps},
                "gpu_launches": launches, "roofline": conv_roofline(conv_events, ms, None, None)}
        if not args.no_cpu_baseline and world == 1:
            v, info = cpu_oracle_voxels_per_s(wl, steps=2, warmup=1, edge=wl["edge"], batch=1, forward_only=True)
            v *= ((wl["edge"] - 2 * wl["overlap"]) / wl["edge"]) ** 3      # useful voxels of a tile = its centre crop
            line["cpu_baseline"] = {"value": v, "unit": "voxels/s", "cores": torch.get_num_threads(), "kind": info["kind"],
                                    "sample": info["sample"] + "; useful (centre-crop) voxels of the tile counted"}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def conv_roofline(conv_events, ms_total, traffic, traffic_src):
    """roofline object for the dominant kernel (tcgen05 implicit-GEMM conv, fprop [+ dgrad]) from its per-launch CUDA
    events inside the timed region."""
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)" if peaks else \
        "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"
    tc_flops = sum(e[0] for e in conv_events)
    ref_flops = sum(e[3] for e in conv_events)
    tc_ms = sum(e[1].elapsed_time(e[2]) for e in conv_events)
    achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    return {"bound": "tensor", "kernel": "conv3_tc_kernel (tcgen05 implicit-GEMM 3x3x3 fprop+dgrad)",
            # achieved = FLOPs the launches EXECUTE / their time (tensor-pipe utilisation, <= 1 of peak).  The decoder
            # join convolutions run 8 summed taps per voxel on the coarse grid instead of the reference op's 27
            # (MEDNET_GATHER_UPCONV_*): reference_equivalent counts the reference op's FLOPs over the same time.
            "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf if peak_tf else None,
            "reference_equivalent": ref_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0,
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "launches_timed": len(conv_events),
            "share_of_step": tc_ms / ms_total if ms_total else None}


def run_sampler(args, wl):
    """--workload sampler: one "step" = one collated patch batch (positions drawn on the host in the reference's order,
    crops [+ augmentation] on the device) ready for training_step.  value: plain sampling; config reports the augmented
    variant; cpu_baseline: the reference's host procedure (oracle restatement, one DataLoader worker = one core) on the
    same cohort, crop + casts + collate, without the pinned H2D copy a training step would add."""
    import time
    import numpy as np
    from mednet_b200 import ops
    from mednet_b200.sampler import GpuMedDataset, IntensityAugmentation
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    images, labels = [], []
    for _ in range(wl["subjects"]):
        img = torch.randn((1, *wl["volume"]), device=dev, generator=g)
        lab = ((img[0] > 1.5) & (torch.rand(wl["volume"], device=dev, generator=g) > 0.5)).to(torch.uint8)
        images.append(img)
        labels.append(lab[None])
    P, B = [wl["edge"]] * 3, wl["batch"]
    vox = B * wl["edge"] ** 3
    steps, warm = max(args.steps, 8), max(args.warmup, 3)
    res = {}
    for name, aug in (("plain", None), ("augmented", IntensityAugmentation())):
        ds = GpuMedDataset(images, labels, 16, P, class_probabilities=wl["probs"], device=dev,
                           rng=np.random.RandomState(1), augmentation=aug)
        idx = np.arange(B)
        for _ in range(warm):
            ds.batch(idx)
        torch.cuda.synchronize()
        n0 = ops.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            ds.batch(idx)
        e1.record()
        torch.cuda.synchronize()
        res[name] = dict(wall_ms=(time.perf_counter() - t0) * 1e3 / steps, dev_ms=e0.elapsed_time(e1) / steps,
                         launches=(ops.launch_count - n0) // steps)
    line = {"metric": "MedDataset patch sampling voxels/s", "value": vox / (res["plain"]["wall_ms"] * 1e-3), "unit": "voxels/s",
            "n_gpus": 1, "steps": steps, "warmup": warm, "ms_per_step": res["plain"]["wall_ms"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "sampler: " + wl["desc"], "timing": "host wall clock around batch() incl. the NumPy draws, "
                       "synchronised at both ends", "device_ms_per_step": res["plain"]["dev_ms"],
                       "augmented_ms_per_step": res["augmented"]["wall_ms"],
                       "augmented_voxels_per_s": vox / (res["augmented"]["wall_ms"] * 1e-3),
                       "augmented_launches": res["augmented"]["launches"], "l2": "source volumes (420 MB) larger than L2"},
            "gpu_launches": res["plain"]["launches"] * steps}
    if not args.no_cpu_baseline:
        from oracle import augment as oaug                    # CPU baseline leg only (see module docstring)
        from oracle import sampling as osamp
        imgs_h, labs_h = [i.cpu().numpy() for i in images], [l.cpu().numpy() for l in labels]
        maps = [osamp.label_any_maps(l[0], len(wl["probs"])) for l in labs_h]
        cpu = {}
        for name, aug in (("plain", False), ("augmented", True)):
            np.random.seed(1)
            t0 = time.perf_counter()
            for _ in range(2):
                data, label = [], []
                for i in range(B):
                    s = i % wl["subjects"]
                    ini, _ = osamp.sample_patch_position(labs_h[s][0], P, wl["probs"], maps[s])
                    d, l = osamp.crop_patch(imgs_h[s], labs_h[s], ini, P)
                    data.append(oaug.augment_patch(d) if aug else d)
                    label.append(l)
                np.stack(data), np.stack(label)
            cpu[name] = vox * 2 / (time.perf_counter() - t0)
        line["cpu_baseline"] = {"value": cpu["plain"], "unit": "voxels/s", "cores": 1, "kind": "port",
                                "sample": "2 batches of 8 x 128^3 (position draws, crop, casts, collate); augmented: "
                                          f"{cpu['augmented']:.4g} voxels/s"}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--conv-impl", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--dual-issue", type=int, default=1, help="tcgen05 conv: second MMA-issuing thread (A/B switch)")
    ap.add_argument("--kd-merge", type=int, default=1, help="tcgen05 conv: kd-merged wide-N MMAs (A/B switch)")
    ap.add_argument("--wgrad-dual", type=int, default=1, help="tcgen05 wgrad: second MMA-issuing thread (A/B switch)")
    ap.add_argument("--class-merge", type=int, default=1,
                    help="tcgen05 fprop/dgrad of the upsample-conv parity classes: kd-merged MMAs, TD + 1 planes (A/B switch)")
    ap.add_argument("--wgrad-reduce-s-fastest", type=int, default=0,
                    help="tcgen05 wgrad split reduction: coalesced thread order (A/B switch)")
    ap.add_argument("--wgrad-class-merge", type=int, default=1,
                    help="tcgen05 wgrad parity-class passes: all tap groups in one role, needed kw windows only (A/B switch)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if wl.get("sampler"):
        if args.impl == "reference":
            raise SystemExit("the sampler workload has no separate reference arm: its line carries the host procedure as cpu_baseline")
        run_sampler(args, wl)
        return
    if args.impl == "reference":
        run_reference_arm(args, wl)
        return
    if wl.get("predict"):
        run_predict(args, wl)
        return
    if args.warmup < 3:
        args.warmup = 3                      # timing rule: >= 3 warm-up steps

    import torch.distributed as dist
    from mednet_b200 import ops
    from mednet_b200.landmarks import LandmarkNet, LandmarkUNet3D
    from mednet_b200.parallel import BucketedAllReduce, init_distributed
    from mednet_b200.segmentation import SegmentationNet, SegmentationUNet3D

    rank, local, world = init_distributed()
    dev = torch.device("cuda", local)
    torch.manual_seed(0)
    from mednet_b200._abi import check, lib
    check(lib().mednet_tcgen05_set_option(b"dual_issue", args.dual_issue), "tcgen05_set_option")
    check(lib().mednet_tcgen05_set_option(b"kd_merge", args.kd_merge), "tcgen05_set_option")
    check(lib().mednet_tcgen05_set_option(b"wgrad_dual_issue", args.wgrad_dual), "tcgen05_set_option")
    check(lib().mednet_tcgen05_set_option(b"wgrad_class_merge", args.wgrad_class_merge), "tcgen05_set_option")
    check(lib().mednet_tcgen05_set_option(b"class_merge", args.class_merge), "tcgen05_set_option")
    check(lib().mednet_tcgen05_set_option(b"wgrad_reduce_s_fastest", args.wgrad_reduce_s_fastest), "tcgen05_set_option")
    hp = hparams_for(wl)
    if wl["arch"] == "residual":
        cls = LandmarkNet if wl["heatmaps"] else SegmentationNet
    else:
        cls = LandmarkUNet3D if wl["heatmaps"] else SegmentationUNet3D
    model = cls(hp, conv_impl=args.conv_impl).to(dev)
    opt = model.configure_optimizers()
    reducer = BucketedAllReduce(opt.grad_slices(), opt.flat_grad) if world > 1 else None
    opt.zero_grad()
    batch_dev = synthetic_batch(wl, 1000 + rank, dev)
    batch_host = synthetic_batch(wl, 1000 + rank, dev, pin=True)
    h2d = sum(v.numel() * v.element_size() for v in batch_host.values())

    inflight = []

    def step(batch):
        # The host is kept at most ONE step ahead of the device (it waits for the end of step i-1 before it enqueues step
        # i+1; a full step of queued work never lets the GPU starve).  Unbounded, the host ran 4-5 steps ahead in a 10-step
        # region but at most 3 in the warm-up, and the first timed region was intermittently 4-6 % slower than the
        # identical region measured after it (81.99 / 79.78 ms against 77.7 / 75.2 ms end-to-end in the same process,
        # profiles/r02/r02n_*, r02o_*; the per-launch conv timings were FASTER in those regions, i.e. the device was
        # waiting, not slow).  The end-to-end region, which reads the loss every step, never showed it.
        if len(inflight) > 1:
            inflight.pop(0).synchronize()
        out = model.training_step(batch, 0)
        out["loss"].backward()
        if reducer is not None:
            opt.grad_scale = reducer.finish()
        opt.step()
        opt.zero_grad()
        ev = torch.cuda.Event()
        ev.record()
        inflight.append(ev)
        return out["loss"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.prepare()
    t_warm = time.perf_counter()
    for _ in range(args.warmup):
        step(batch_dev)
    # settle: W steps can end before the caching allocator, lazily loaded kernel variants and the power state have
    # (a 3-step warm-up was measured 7 % slower in the first timed region than in the one after it); keep stepping,
    # untimed, until the warm-up has lasted ~1.5 s, and run the host-batch path once as well
    torch.cuda.synchronize()
    recent, best = [], float("inf")
    while True:
        t0 = time.perf_counter()
        step(batch_dev)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = min(best, dt)
        recent = (recent + [dt])[-3:]
        elapsed = time.perf_counter() - t_warm
        # settled: >= 1.5 s of warm-up and the last three steps within 2 % of the fastest seen (or 8 s, whatever comes first)
        if elapsed > 8.0 or (elapsed > 1.5 and len(recent) == 3 and max(recent) < 1.02 * best):
            break
    step({k: v.to(dev, non_blocking=True) for k, v in batch_host.items()}).item()
    for _ in range(args.steps):              # dress rehearsal of the timed region (same queueing pattern, untimed)
        step(batch_dev)
    torch.cuda.synchronize()
    sampling = rank == 0 and not os.environ.get("MEDNET_BENCH_NOSAMPLER")
    if sampling:
        sampler.start()
    ops.conv_events = None if os.environ.get("MEDNET_BENCH_NOEVENTS") else []   # per-launch CUDA events (see ops.k_conv3)
    launches0 = ops.launch_count
    ms = timed(lambda: step(batch_dev), args.steps)
    launches = ops.launch_count - launches0
    conv_events, ops.conv_events = ops.conv_events or [], None
    clocks = sampler.stop() if sampling else None

    def e2e_step():
        b = {k: v.to(dev, non_blocking=True) for k, v in batch_host.items()}
        return step(b).item()                # D2H read of the step's result (4 bytes) inside the timed region

    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)

    vox_step = world * wl["batch"] * wl["edge"] ** 3
    value = vox_step * args.steps / (ms * 1e-3)
    e2e = vox_step * args.steps / (ms_e2e * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    traffic, traffic_src = None, None
    if args.workload == "cfg3":
        try:                                  # DRAM bytes per launch of the dominant kernel from the committed ncu capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic, traffic_src = tj["dram_bytes_per_launch"], "static (not measured in this run): " + tj["source"]
        except (OSError, KeyError, ValueError):
            pass
    roofline = conv_roofline(conv_events, ms, traffic, traffic_src)
    line = {"metric": "UNet3D train voxels/s", "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": args.workload + ": " + wl["desc"], "per_gpu_batch": wl["batch"], "patch": wl["edge"],
                       "parallelism": f"dp{world}", "l2": "inputs larger than L2 (activations are GBs per step)",
                       "host_run_ahead": "at most one step",
                       "conv_impl": args.conv_impl, "tcgen05_variants": {str(k): v for k, v in ops.tcgen05_variants.items()
                                                                           if k != "report"}},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "voxels/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "roofline": roofline}
    if not args.no_cpu_baseline and world == 1:
        # bounded sample of the same workload: ~15-30 s of host work (batch of one on the workload's own patch edge)
        v, info = cpu_oracle_voxels_per_s(wl, steps=2 if wl["edge"] > 96 else 4, warmup=1, edge=wl["edge"], batch=1)
        line["cpu_baseline"] = {"value": v, "unit": "voxels/s", "cores": torch.get_num_threads(), "kind": info["kind"],
                                "sample": info["sample"]}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

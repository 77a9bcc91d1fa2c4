"""Patch-batch data parallelism: one process per GPU, bucketed gradient all-reduce overlapped with the
backward pass (replaces what pytorch-lightning's DDP did for `Trainer(gpus=N)`, examples/train_seg.py:126).

GroupNorm statistics are per sample (components.py:57), so the only exchange is the gradient all-reduce.
Gradients live in ONE flat fp32 buffer (FusedAdam); buckets are contiguous slices of it laid out in
backward-completion order (final_conv -> decoders -> encoders, i.e. reverse registration order).  A
post-accumulate hook per parameter counts its bucket down; the last arrival launches an asynchronous
`all_reduce(SUM)` of the bucket view on the process group's communication stream (NCCL over NVLink), which
therefore overlaps the rest of backward.  The 1/world scaling is folded into the fused Adam kernel.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class FlatGradients:
    """Device-agnostic flat gradient buffer (used directly by the CPU/gloo tests; FusedAdam owns the GPU one)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.slices = []
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            self.slices.append((p, off, n))
            off += n

    def zero(self):
        self.flat.zero_()


class BucketedAllReduce:
    def __init__(self, grad_slices, flat_grad, bucket_bytes=25 << 20, process_group=None):
        """grad_slices: [(param, offset, numel)] in registration order; flat_grad: the flat fp32 buffer."""
        self.flat = flat_grad
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        cap = max(1, bucket_bytes // 4)
        # buckets over the flat buffer from the END (backward order); each is one contiguous range
        self.buckets = []          # [lo, hi, remaining, total]
        self.bucket_of = {}
        hi = lo = flat_grad.numel()
        members = 0
        for p, off, n in reversed(grad_slices):
            if members and (hi - off) > cap:
                self.buckets.append([lo, hi, members, members])
                hi, members = lo, 0
            lo = off
            self.bucket_of[id(p)] = len(self.buckets)
            members += 1
        if members:
            self.buckets.append([lo, hi, members, members])
        self.handles = []
        self.launch_order = []
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_hook) for p, _, _ in grad_slices]
        # conv weight gradients enqueued on the side stream (ops.wgrad_async) bypass autograd's accumulation and report
        # through the listener; the all-reduce of a bucket is then launched BEHIND the side stream.  autograd still calls
        # the post-accumulate hook of such a parameter (with an undefined gradient), so the hook skips parameters that
        # have already reported in this backward pass.
        self._async_reported = set()
        self._ops = None
        if flat_grad.is_cuda:
            from . import ops
            self._ops = ops
            ops.async_grad_listener = self._on_async

    def _on_async(self, param):
        self._async_reported.add(id(param))
        self._on_grad(param)

    def _on_hook(self, param):
        if id(param) in self._async_reported:
            return
        self._on_grad(param)

    def _on_grad(self, param):
        b = self.bucket_of[id(param)]
        bucket = self.buckets[b]
        bucket[2] -= 1
        if bucket[2] == 0:
            self._launch(b)

    def _launch(self, b):
        lo, hi = self.buckets[b][0], self.buckets[b][1]
        self.launch_order.append(b)
        if self.world > 1:
            if self._ops is not None:
                # the bucket holds gradients written on the main stream (autograd) and on the side stream (async wgrad):
                # issue the collective from the side stream after it has caught up with the main stream
                side = self._ops.side_stream(self.flat.device)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    h = dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                self.handles.append(h)
            else:
                self.handles.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Call after backward: flushes buckets whose parameters got no gradient, waits for the collectives
        and re-arms the counters.  Returns the scale (1/world) the optimiser must apply."""
        for b, bucket in enumerate(self.buckets):
            if bucket[2] != 0:
                self._launch(b)
        for h in self.handles:
            h.wait()
        self.handles.clear()
        self._async_reported.clear()
        if self._ops is not None:
            self._ops.sync_async_wgrad()
        self.launch_order.clear()
        for bucket in self.buckets:
            bucket[2] = bucket[3]
        return 1.0 / self.world

    def remove(self):
        for h in self._hooks:
            h.remove()
        if self._ops is not None and self._ops.async_grad_listener == self._on_async:
            self._ops.async_grad_listener = None


def init_distributed(backend=None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun); returns (rank, local_rank, world)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, local, world


def rank_and_world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def epoch_order(n, shuffle, epoch, seed=0, rank=0, world=1):
    """Sample order of one epoch for one rank: a permutation every rank derives identically (torch generator seeded
    with seed + epoch, the rule of torch's DistributedSampler), truncated to a multiple of ``world`` so that all ranks
    run the same number of steps (their collectives pair up), then taken with stride ``world``."""
    if shuffle:
        g = torch.Generator()
        g.manual_seed(int(seed) + int(epoch))
        order = torch.randperm(n, generator=g).numpy()
    else:
        import numpy as np
        order = np.arange(n)
    if world > 1:
        order = order[:n - n % world][rank::world]
    return order


def sharded_loader(dataset, batch_size, num_workers=0, shuffle=True, epoch=0, seed=0, rank=None, world=None):
    """The loader a task module's train/val_dataloader hook returns (segmentation.py:122-132): patch batches are the
    unit of data parallelism, so each rank must see its own share of an epoch.  A device-resident dataset
    (``sampler.GpuMedDataset``) iterates itself; anything else goes through a DataLoader over this rank's indices."""
    from torch.utils.data import DataLoader, Subset
    if rank is None or world is None:
        rank, world = rank_and_world()
    if hasattr(dataset, "loader"):
        return dataset.loader(batch_size, shuffle=shuffle, epoch=epoch, seed=seed, rank=rank, world=world)
    if world == 1:
        return DataLoader(dataset, batch_size=batch_size, num_workers=num_workers, shuffle=shuffle)
    order = epoch_order(len(dataset), shuffle, epoch, seed, rank, world)
    return DataLoader(Subset(dataset, order.tolist()), batch_size=batch_size, num_workers=num_workers, shuffle=False)

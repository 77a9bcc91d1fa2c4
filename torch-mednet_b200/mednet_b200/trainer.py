"""A minimal training loop calling the PL-0.9 hooks of the task modules with the reference's Trainer
arguments (examples/train_seg.py:126-132): gpus, max_epochs, default_root_dir, resume_from_checkpoint.
pytorch-lightning itself is not in this image; device placement, data parallelism and checkpointing it
provided are done here.
"""
from __future__ import annotations

import os

import torch

from .parallel import BucketedAllReduce, init_distributed


def _to_device(batch, dev):
    return {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in batch.items()}


class Trainer:
    def __init__(self, gpus=1, max_epochs=1, default_root_dir=None, resume_from_checkpoint=None, logger=None,
                 max_steps=None, log_every=10):
        self.gpus, self.max_epochs, self.root = gpus, max_epochs, default_root_dir
        self.resume, self.max_steps, self.log_every = resume_from_checkpoint, max_steps, log_every
        self.history = []

    def fit(self, model):
        rank, local, world = init_distributed()
        dev = torch.device("cuda", local)
        model.to(dev)
        opt = model.configure_optimizers()
        start_epoch, step = 0, 0
        if self.resume:
            # pytorch-lightning's resume_from_checkpoint (examples/train_seg.py:122-131): weights, epoch / step counters
            # AND the optimiser state (Adam moments + step count), so the resumed run continues the uninterrupted one
            ckpt = torch.load(self.resume, map_location="cpu", weights_only=False)
            model.load_state_dict(ckpt["state_dict"])
            start_epoch, step = ckpt.get("epoch", 0), ckpt.get("global_step", 0)
            states = ckpt.get("optimizer_states")
            if states:
                opt.load_state_dict(states[0])
        reducer = None
        if world > 1:
            import torch.distributed as dist
            flat = opt.flat_grad                               # materialises the flat parameter / gradient buffers
            dist.broadcast(opt._flat, src=0)                   # replicas start from rank 0's weights (PL's DDP does this)
            dist.broadcast(opt._m, src=0)
            dist.broadcast(opt._v, src=0)
            reducer = BucketedAllReduce(opt.grad_slices(), flat)
        opt.zero_grad()
        for epoch in range(start_epoch, self.max_epochs):
            model.current_epoch = epoch
            model.train()
            for i, batch in enumerate(model.train_dataloader()):
                out = model.training_step(_to_device(batch, dev), i)
                out["loss"].backward()
                if reducer is not None:
                    opt.grad_scale = reducer.finish()
                opt.step()
                opt.zero_grad()
                step += 1
                model.global_step = step
                if rank == 0 and step % self.log_every == 0:
                    self.history.append({k: float(v) for k, v in out["log"].items()})
                if self.max_steps and step >= self.max_steps:
                    break
            if getattr(model, "validation_dataset", None) is not None:
                model.eval()
                outs = []
                with torch.no_grad():
                    for i, batch in enumerate(model.val_dataloader()):
                        outs.append(model.validation_step(_to_device(batch, dev), i))
                if outs:
                    logs = model.validation_epoch_end(outs)["log"]
                    if world > 1:                              # every rank validated its own share: average them
                        import torch.distributed as dist
                        vals = torch.stack([torch.as_tensor(v, dtype=torch.float32, device=dev) for v in logs.values()])
                        dist.all_reduce(vals)
                        logs = dict(zip(logs.keys(), vals / world))
                    self.history.append({k: float(v) for k, v in logs.items()})
            if rank == 0 and self.root:
                os.makedirs(self.root, exist_ok=True)
                model.save_checkpoint(os.path.join(self.root, f"epoch={epoch}.ckpt"), opt, epoch + 1, step)
            if self.max_steps and step >= self.max_steps:
                break
        return model

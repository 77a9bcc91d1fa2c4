"""Fused Adam over one flat fp32 bucket (ref: torch.optim.Adam(self.parameters(), lr) at
midasmednet/segmentation.py:119-120, landmarks.py:176-177 -- PyTorch defaults, no weight decay).

All parameters are re-pointed into one contiguous buffer (and so are their gradients), so the optimiser
step is a single kernel launch and the data-parallel all-reduce can work on bucket views of that buffer.

Asynchronous weight gradients (default): the conv weights' ``.grad`` views are marked for ``ops.wgrad_async``, i.e.
their gradients are accumulated into the flat buffer by kernels on a side stream that overlap the rest of backward.
``step()``, ``zero_grad()`` and the data-parallel reducer join that stream themselves; code that reads ``p.grad``
between ``backward()`` and ``step()`` (gradient clipping, logging) must call ``optimizer.sync_gradients()`` first, or
construct the optimiser with ``async_wgrad=False``.
"""
from __future__ import annotations

import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, async_wgrad=True):
        if weight_decay != 0:
            raise NotImplementedError("the reference uses Adam without weight decay")
        params = list(params)
        if params and isinstance(params[0], dict):
            if len(params) > 1:
                raise NotImplementedError("FusedAdam updates one flat bucket with one set of hyper-parameters: pass a single "
                                          "parameter group (the reference does: Adam(self.parameters(), lr))")
            params = list(params[0]["params"])
        params = [p for p in params if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._params = params
        self._flat = self._flat_grad = self._m = self._v = None
        self._step = 0
        self.grad_scale = 1.0
        # weight gradients are written into the flat buffer on a side stream, overlapped with the rest of backward;
        # every reader of the buffer (step, zero_grad, the all-reduce) first joins that stream
        self.async_wgrad = bool(async_wgrad) and torch.cuda.is_available()

    # -- flat storage ---------------------------------------------------------------------------
    def _materialize(self):
        if self._flat is not None:
            return
        dev = self._params[0].device
        total = sum(p.numel() for p in self._params)
        self._flat = torch.empty(total, dtype=torch.float32, device=dev)
        self._flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self._m = torch.zeros(total, dtype=torch.float32, device=dev)
        self._v = torch.zeros(total, dtype=torch.float32, device=dev)
        self._offsets = []
        off = 0
        with torch.no_grad():
            for p in self._params:
                n = p.numel()
                view = self._flat[off:off + n].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self._flat_grad[off:off + n].view_as(p)
                if self.async_wgrad and p.dim() == 5:
                    ops.mark_async_grad(p)            # conv weights: wgrad goes to a side stream (ops.wgrad_async)
                self._offsets.append((off, n))
                off += n

    @property
    def flat_grad(self):
        self._materialize()
        return self._flat_grad

    def grad_slices(self):
        """[(param, offset, numel)] in registration order (used to build all-reduce buckets)."""
        self._materialize()
        return [(p, o, n) for p, (o, n) in zip(self._params, self._offsets)]

    def sync_gradients(self):
        """Join the side stream: after this the .grad tensors are complete on the current stream."""
        ops.sync_async_wgrad()

    def _realias(self):
        """Parameters must still be views of the flat buffer: a `model.to()` / `.float()` / `load_state_dict(assign=True)`
        after the first step replaces `p.data` and the kernel would silently update storage nobody reads."""
        base = self._flat.data_ptr()
        for p, (off, n) in zip(self._params, self._offsets):
            if p.data_ptr() != base + 4 * off:
                if p.device != self._flat.device or p.dtype != torch.float32:
                    raise RuntimeError("FusedAdam: a parameter was moved or cast after the optimiser took ownership of it "
                                       f"(now {p.device}/{p.dtype}); rebuild the optimiser after moving the model")
                view = self._flat[off:off + n].view_as(p)
                view.copy_(p.data)
                p.data = view

    # -- checkpointing: torch.optim.Adam's own state_dict format (what pytorch-lightning stores under
    #    'optimizer_states', examples/train_seg.py:122-131 resume_from_checkpoint) ------------------------------------
    def state_dict(self):
        groups = [{**{k: v for k, v in g.items() if k != "params"}, "params": list(range(len(self._params)))}
                  for g in self.param_groups]
        state = {}
        if self._flat is not None and self._step > 0:
            for i, (p, (off, n)) in enumerate(zip(self._params, self._offsets)):
                state[i] = {"step": torch.tensor(float(self._step)),
                            "exp_avg": self._m[off:off + n].view_as(p).clone(),
                            "exp_avg_sq": self._v[off:off + n].view_as(p).clone()}
        return {"state": state, "param_groups": groups}

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        self._materialize()
        groups = state_dict["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self._params):
            raise ValueError("loaded state dict does not match this optimiser's single parameter group")
        for k in ("lr", "betas", "eps"):
            if k in groups[0]:
                self.param_groups[0][k] = groups[0][k]
        state = state_dict["state"]
        self._m.zero_()
        self._v.zero_()
        self._step = 0
        steps = set()
        for i, (p, (off, n)) in enumerate(zip(self._params, self._offsets)):
            st = state.get(i, state.get(str(i)))
            if st is None:
                continue
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError(f"optimizer state {i}: shape {tuple(st['exp_avg'].shape)} vs parameter {tuple(p.shape)}")
            self._m[off:off + n].view_as(p).copy_(st["exp_avg"])
            self._v[off:off + n].view_as(p).copy_(st["exp_avg_sq"])
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"FusedAdam keeps one step count for the bucket; the loaded state has {sorted(steps)}")
        if steps:
            self._step = steps.pop()

    def zero_grad(self, set_to_none=False):
        self._materialize()
        ops.sync_async_wgrad()
        self._flat_grad.zero_()
        for p, (off, n) in zip(self._params, self._offsets):
            if p.grad is None or p.grad.data_ptr() != self._flat_grad.data_ptr() + 4 * off:
                p.grad = self._flat_grad[off:off + n].view_as(p)

    @torch.no_grad()
    def step(self, closure=None):
        self._materialize()
        ops.sync_async_wgrad()
        self._realias()
        # gradients written by autograd into fresh tensors (first backward) are folded into the flat buffer
        for p, (off, n) in zip(self._params, self._offsets):
            if p.grad is not None and p.grad.data_ptr() != self._flat_grad.data_ptr() + 4 * off:
                self._flat_grad[off:off + n].view_as(p).copy_(p.grad)
                p.grad = self._flat_grad[off:off + n].view_as(p)
        self._step += 1
        g = self.param_groups[0]
        ops.k_adam(self._flat, self._flat_grad, self._m, self._v, self._step, g["lr"], g["betas"][0], g["betas"][1],
                   g["eps"], self.grad_scale)
        return None

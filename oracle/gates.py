"""Parity gates shared by tests/, __graft_entry__.smoke() (test infrastructure; see oracle/__init__.py).

North-star gates: logits rel. err <= 1e-2 (bf16) / <= 1e-4 (fp32 validation mode), loss within 1e-3, per-parameter
gradient cosine >= 0.999.  The gradient gate needs care in bf16 for the ReLU network ('gcr' UNet3D):

  * ReLU'(x) and the max-pool routing are discontinuous.  A storage rounding of relative size eps flips a fraction
    ~eps of those decisions and every flip changes the back-propagated signal by O(1), so the gradient error grows
    like sqrt(eps) * depth, not eps: the REFERENCE ITSELF evaluated with bf16 storage (oracle.unet.Storage.bf16 ==
    `model.bfloat16()` of the reference) reaches only cosine 0.94-0.97 against its own fp32 gradients on the early
    encoder layers, at every volume size, and fp16 storage (8x finer) still only 0.96
    (tests/test_oracle_golden.py::test_gradient_sensitivity_is_decision_flip_noise pins this on the CPU).
    The smooth 'cge' ResidualUNet3D (ELU) keeps >= 0.999 under the same storage.
  * So no bf16 implementation can meet cos >= 0.999 against the FP32 gradients of the ReLU net; what a correct bf16
    implementation can do is be as close to the fp32 gradient as the bf16-storage reference is.  The gate therefore
    requires, per parameter tensor,   1 - cos(ours, fp32)  <=  max(1e-3, 2 * (1 - cos(bf16-storage reference, fp32)))
    i.e. >= 0.999 wherever the format allows it and never more than twice the format's own angular error elsewhere.
  * A cosine over a handful of numbers is noise (GroupNorm(1,1) of the input image has ONE gamma: cos = +-1), so
    tensors with fewer than `min_numel` elements are pooled per kind (all such '.weight' / '.bias' vectors
    concatenated) and gated as one vector.
The strict cos >= 0.999 (in fact 0.9999) gate is enforced per operator against fp32 PyTorch (tests/test_ops_gpu.py,
tests/test_tcgen05_gpu.py) and for whole networks in the fp32 validation mode (tests/test_model_gpu.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def cosine(a, b):
    a = torch.as_tensor(a).detach().cpu().double().flatten()
    b = torch.as_tensor(b).detach().cpu().double().flatten()
    # not F.cosine_similarity: it clamps the norm product at 1e-8, which small-magnitude gradients fall below
    return (torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-300)).item()


def relerr(a, b):
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _pooled(grads, names):
    return torch.cat([torch.as_tensor(grads[k]).detach().cpu().double().flatten() for k in names])


def format_aware_gradient_gate(ours, ref, fmt, min_numel=64, floor=1e-3, factor=2.0):
    """ours / ref / fmt: {name: gradient}.  Returns (failures, report) where report rows are
    (name, cos(ours, ref), cos(fmt, ref), numel); failures is the subset violating the gate above."""
    rows, small = [], {}
    for k in ref:
        n = torch.as_tensor(ref[k]).numel()
        if n < min_numel:
            small.setdefault(k.rsplit(".", 1)[-1], []).append(k)
            continue
        rows.append((k, cosine(ours[k], ref[k]), cosine(fmt[k], ref[k]), n))
    for kind, names in small.items():
        r = _pooled(ref, names)
        rows.append((f"<pooled {len(names)} small '{kind}' tensors>", cosine(_pooled(ours, names), r),
                     cosine(_pooled(fmt, names), r), r.numel()))
    failures = [row for row in rows if (1.0 - row[1]) > max(floor, factor * (1.0 - row[2]))]
    return failures, rows

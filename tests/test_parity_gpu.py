"""bf16 / tensor-core parity of the production UNet3D against the oracle at BASELINE.json's sizes (cfg-3: f = 64,
128^3 patches) -- the evidence behind DESIGN.md section 4.

1. LAYER-WISE, teacher-forced: every one of the 14 GroupNorm+conv+ReLU layers, the pools, the three virtual-concat
   joins, the 1x1x1 head and the Dice loss reproduce the bf16-storage oracle's output, input gradient and parameter
   gradients on the oracle's OWN layer inputs to <= 5e-3 (measured ~1e-3 and below): the kernels are right; what an
   end-to-end comparison shows beyond that is propagation through the network.
2. END-TO-END on a trained-like state (the product's own bf16 training steps on a learnable target), against the
   FP32 reference: logits <= 1e-2 and loss within 1e-3 literally as BASELINE.json:north_star states them; parameter
   gradients next to `reference.bfloat16()` on cuDNN (the like-for-like comparator): the product must be at least as
   close to the fp32 gradients as that.
3. END-TO-END at random initialisation (what bench.py runs): forward AND backward at full size, same comparators.

Set MEDNET_PARITY_REPORT=<dir> to also write the measured tables as JSON (committed under profiles/)."""
import json
import os

import pytest
import torch

import parity_lib as pl
from mednet_b200.unet.model import UNet3D

pytestmark = pytest.mark.gpu
DEV = "cuda"
W4 = torch.tensor([0.2, 1.0, 0.7, 0.5])


def _dump(name, obj):
    d = os.environ.get("MEDNET_PARITY_REPORT")
    if d:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, name + ".json"), "w") as f:
            json.dump(obj, f, indent=1)


def _print_rows(rows):
    for r in rows:
        print("  %-48s %-18s rel %.2e  cos %.6f" % r)


def _print_e2e(tag, rep):
    print(f"[{tag}] logits rel. error vs fp32 reference: ours {rep['logits_rel_ours_vs_fp32']:.2e}, bf16-storage oracle "
          f"{rep['logits_rel_bf16storage_vs_fp32']:.2e}, reference.bfloat16() on cuDNN {rep['logits_rel_native_bf16_vs_fp32']:.2e}; "
          f"ours vs bf16-storage oracle {rep['logits_rel_ours_vs_bf16storage']:.2e}")
    print(f"[{tag}] loss fp32 {rep['loss_fp32']:.6f} ours {rep['loss_ours']:.6f} native bf16 {rep['loss_native_bf16']:.6f}; label flips vs fp32: "
          f"ours {rep['label_flips_ours_vs_fp32']:.5f} native {rep['label_flips_native_vs_fp32']:.5f}")
    print(f"[{tag}] gradient cosine vs fp32: min ours {rep['grad_cos_min_ours']:.5f}, bf16-storage {rep['grad_cos_min_bf16storage']:.5f}, "
          f"native bf16 {rep['grad_cos_min_native_bf16']:.5f}; tensors >= 0.999: ours {rep['grad_tensors_ge_0.999_ours']}/{rep['grad_tensors']}, "
          f"native {rep['grad_tensors_ge_0.999_native_bf16']}/{rep['grad_tensors']}")
    for r in rep["grad_rows"]:
        print("    %-52s ours %.5f  bf16-storage %.5f  native-bf16 %.5f  (%d)" % r)


def _gates_vs_native(rep, slack=1.5):
    """The product is at least as close to the fp32 reference as `reference.bfloat16()` on cuDNN is (with `slack` for
    run-to-run summation-order noise), tensor by tensor."""
    assert rep["logits_rel_ours_vs_fp32"] <= slack * rep["logits_rel_native_bf16_vs_fp32"] + 1e-3
    bad = [r for r in rep["grad_rows"] if (1 - r[1]) > max(1e-3, slack * (1 - r[3]))]
    assert not bad, bad


@pytest.mark.parametrize("state", ["init", "trained"])
def test_layerwise_teacher_forced_parity_cfg3_size(state):
    torch.manual_seed(0)
    net = UNet3D(1, 4, False).to(DEV)                              # f = 64, 4 levels: the cfg-3 network
    if state == "trained":
        pl.train_state(net, 40, (128,) * 3, 4, W4.to(DEV))
    x, y = pl.learnable_batch(1, (128,) * 3, 4, seed=7)
    rows = pl.layerwise_report(net, x, y, W4.to(DEV), 64)
    _print_rows(rows)
    _dump(f"parity_layerwise_{state}", [list(r) for r in rows])
    assert len([r for r in rows if r[1] == "y"]) == 14
    for layer, what, rel, c in rows:
        if "exact" in what:
            assert rel == 0.0, (layer, what, rel)
        elif "abs diff" in what:
            assert rel < 1e-5, (layer, what, rel)
        elif "not gated" in what:                       # the comparator's own noise, reported next to the gated row
            continue
        else:
            assert rel <= 5e-3 and c >= 0.9999, (layer, what, rel, c)


def test_trained_state_meets_the_north_star_gates_literally_full_size():
    torch.manual_seed(0)
    net = UNet3D(1, 4, False).to(DEV)
    losses = pl.train_state(net, 60, (128,) * 3, 4, W4.to(DEV))
    print("training loss trajectory (product, bf16):", ["%.4f" % v for v in losses])
    assert losses[-1] < 0.9 * losses[0]
    x, y = pl.learnable_batch(1, (128,) * 3, 4, seed=7)            # a batch the training never saw
    rep = pl.end_to_end_report(net, x, y, W4.to(DEV), 64)
    _print_e2e("trained f=64 128^3", rep)
    _dump("parity_e2e_trained_f64_128", rep)
    assert rep["logits_rel_ours_vs_fp32"] <= 1e-2                  # north star, literally, against the FP32 reference
    assert abs(rep["loss_ours"] - rep["loss_fp32"]) <= 1e-3
    assert rep["label_flips_ours_vs_fp32"] <= 5e-3
    _gates_vs_native(rep)


def test_random_init_forward_and_backward_full_size():
    torch.manual_seed(0)
    net = UNet3D(1, 4, False).to(DEV)
    x, y = pl.learnable_batch(1, (128,) * 3, 4, seed=7)
    rep = pl.end_to_end_report(net, x, y, W4.to(DEV), 64)
    _print_e2e("random init f=64 128^3", rep)
    _dump("parity_e2e_init_f64_128", rep)
    assert abs(rep["loss_ours"] - rep["loss_fp32"]) <= 1e-3
    _gates_vs_native(rep)


def test_trained_state_small_net():
    """The same literal gates on the smoke-sized network (what __graft_entry__.smoke() prints)."""
    torch.manual_seed(0)
    f_maps = [16, 32, 64]
    w = torch.tensor([0.2, 1.0, 0.7]).to(DEV)
    net = UNet3D(1, 3, False, f_maps=f_maps).to(DEV)
    pl.train_state(net, 60, (16, 32, 16), 3, w, batch=2)
    x, y = pl.learnable_batch(2, (16, 32, 16), 3, seed=7)
    rep = pl.end_to_end_report(net, x, y, w, f_maps)
    _print_e2e("trained [16,32,64] 16x32x16", rep)
    _dump("parity_e2e_trained_small", rep)
    assert rep["logits_rel_ours_vs_fp32"] <= 1e-2
    assert abs(rep["loss_ours"] - rep["loss_fp32"]) <= 1e-3
    _gates_vs_native(rep, slack=2.0)

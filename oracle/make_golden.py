"""Generate tests/golden/*.npz from the LIVE reference (run in the build container only).

    python oracle/make_golden.py            # needs /root/reference; writes tests/golden/

The reference's model/loss modules are imported unchanged from /root/reference with a 2-line
``pytorch_lightning`` stub (LightningModule = torch.nn.Module).
``grid_patch_generator`` cannot be imported (dataset.py pulls in h5py/zarr), so its source text is
extracted with ``ast`` and executed against NumPy -- the function body itself is unmodified.

The GPU box has no /root/reference: tests only read the committed .npz files.
"""
from __future__ import annotations

import ast
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _import_reference():
    stub = types.ModuleType("pytorch_lightning")
    stub.LightningModule = torch.nn.Module
    sys.modules["pytorch_lightning"] = stub
    sys.path.insert(0, REF)
    from midasmednet.unet import loss as rloss      # noqa: E402
    from midasmednet.unet import model as rmodel    # noqa: E402
    return rmodel, rloss


def _np_sd(module):
    return {k: v.detach().cpu().numpy() for k, v in module.state_dict().items()}


def _grads(module, value):
    module.zero_grad()
    value.backward()
    return {"grad." + k: p.grad.detach().cpu().numpy() for k, p in module.named_parameters()}


def golden_unet3d(rmodel, rloss):
    torch.manual_seed(1234)
    net = rmodel.UNet3D(1, 2, False, f_maps=[8, 16, 32])
    # perturb the affine GroupNorm parameters so that gamma/beta gradients and usage are exercised
    with torch.no_grad():
        for k, p in net.named_parameters():
            if "groupnorm" in k:
                p.add_(0.25 * torch.randn_like(p))
    x = torch.randn(2, 1, 16, 16, 16)
    y = torch.randint(0, 2, (2, 16, 16, 16))
    logits = net(x)
    dice = rloss.DiceLoss(weight=torch.tensor([0.05, 1.0]))(logits, y)
    ce = torch.nn.CrossEntropyLoss(weight=torch.tensor([0.3, 0.7]))(logits, y)
    metric = rloss.dice_metric(logits, y)
    out = {"x": x.numpy(), "y": y.numpy(), "logits": logits.detach().numpy(), "dice": dice.detach().numpy(),
           "ce": ce.detach().numpy(), "dice_metric": metric.detach().numpy()}
    out.update({"sd." + k: v for k, v in _np_sd(net).items()})
    out.update(_grads(net, dice))
    # odd size (pool floors, interpolate(size=...) restores -- SURVEY.md section 7.2)
    xo = torch.randn(1, 1, 13, 14, 15)
    out["x_odd"] = xo.numpy()
    out["logits_odd"] = net(xo).detach().numpy()
    # test-time activation (model.py:107-108)
    net.testing = True
    out["probs"] = net(x).detach().numpy()
    np.savez_compressed(os.path.join(OUT, "unet3d_small.npz"), **out)


def golden_unet3d_orders(rmodel):
    out = {}
    for order in ("crg", "cl", "gce"):
        torch.manual_seed(77)
        net = rmodel.UNet3D(2, 3, False, f_maps=[8, 16], layer_order=order)
        x = torch.randn(1, 2, 8, 12, 8)
        out[f"{order}.x"] = x.numpy()
        out[f"{order}.logits"] = net(x).detach().numpy()
        out.update({f"{order}.sd." + k: v for k, v in _np_sd(net).items()})
    np.savez_compressed(os.path.join(OUT, "unet3d_orders.npz"), **out)


def golden_residual(rmodel, rloss):
    torch.manual_seed(4321)
    L, K = 2, 2
    net = rmodel.ResidualUNet3D(1, L + K, False, f_maps=[8, 16, 32])
    with torch.no_grad():
        for k, p in net.named_parameters():
            if "groupnorm" in k:
                p.add_(0.25 * torch.randn_like(p))
    x = torch.randn(2, 1, 16, 16, 16)
    label = torch.randint(0, K, (2, 16, 16, 16))
    hm = torch.randint(0, 256, (2, L, 16, 16, 16)).float()
    outputs = net(x)
    # LandmarkNet.loss, landmarks.py:125-134 (restated: landmarks.py is not importable here)
    class_loss = rloss.DiceLoss(weight=torch.tensor([0.05, 1.0]))(outputs[:, L:], label)
    reg_w = [0.001, 0.015]
    regression = torch.tensor(0.0)
    for c in range(L):
        regression = regression + reg_w[c] * torch.nn.MSELoss()(outputs[:, c], hm[:, c])
    total = regression + class_loss
    out = {"x": x.numpy(), "label": label.numpy(), "heatmaps": hm.numpy(), "outputs": outputs.detach().numpy(),
           "class_loss": class_loss.detach().numpy(), "regression_loss": regression.detach().numpy(),
           "loss": total.detach().numpy(), "reg_w": np.array(reg_w, dtype=np.float32)}
    out.update({"sd." + k: v for k, v in _np_sd(net).items()})
    out.update(_grads(net, total))
    # plain segmentation use (SegmentationNet wiring, segmentation.py:43-49)
    seg = rloss.DiceLoss(weight=torch.tensor([0.05, 1.0, 1.0, 1.0]))(outputs.detach(), torch.randint(0, 4, (2, 16, 16, 16)))
    out["seg_dice_4class"] = seg.numpy()
    np.savez_compressed(os.path.join(OUT, "residual_small.npz"), **out)


def golden_tiling():
    src = open(os.path.join(REF, "midasmednet", "dataset.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "grid_patch_generator")
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "dataset.py", "exec"), ns)
    gen = ns["grid_patch_generator"]
    rng = np.random.default_rng(5)
    out = {}
    for tag, shape, p, o in (("a", (50, 41, 37), 24, 4), ("b", (32, 32, 48), 16, 4), ("c", (20, 33, 17), 16, 2)):
        img = rng.standard_normal((2,) + shape).astype(np.float32)
        pos, sums = [], []
        for patch, idx, count in gen(img, [p] * 3, [o] * 3, mode="constant"):
            pos.append(idx)
            sums.append(patch.astype(np.float64).sum())
        out[f"{tag}.img"] = img
        out[f"{tag}.patch"] = np.array([p, o])
        out[f"{tag}.pos"] = np.array(pos)
        out[f"{tag}.sums"] = np.array(sums)
    np.savez_compressed(os.path.join(OUT, "tiling.npz"), **out)


def golden_sampling():
    """Patch positions from the reference's own get_labeled_position / get_random_patch_indices (dataset.py:18-88),
    function bodies extracted with ast; ``np.int`` (removed from NumPy) is mapped to ``int``."""
    src = open(os.path.join(REF, "midasmednet", "dataset.py")).read()
    fns = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and
           n.name in ("get_labeled_position", "get_random_patch_indices")]
    np_shim = types.SimpleNamespace(**{k: getattr(np, k) for k in dir(np) if not k.startswith("__")})
    np_shim.int = int
    ns = {"np": np_shim}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "dataset.py", "exec"), ns)
    rng = np.random.default_rng(5)
    label = np.zeros((40, 37, 29), dtype=np.uint8)
    label[5:12, 20:30, 3:9] = 1
    label[30:38, 2:9, 15:27] = 2
    label[(rng.random(label.shape) < 0.002)] = 1
    patch = np.array([16, 12, 10])
    probs = np.array([0.2, 0.5, 0.3])
    np.random.seed(77)
    ini, cls = [], []
    for _ in range(64):
        c = np.random.choice(range(len(probs)), p=probs / probs.sum())           # dataset.py:301-302
        pos = ns["get_labeled_position"](label, c, label_any=np.any(label == c, axis=2)) if c > 0 else None
        a, b = ns["get_random_patch_indices"](patch, np.array(label.shape), pos=pos)
        assert np.array_equal(b - a, patch)
        ini.append(a)
        cls.append(c)
    np.savez_compressed(os.path.join(OUT, "sampling.npz"), label=label, patch=patch, probs=probs, seed=77,
                        index_ini=np.array(ini), selected_class=np.array(cls))


def golden_loss_variants(rloss):
    """Loss configurations the small-net fixtures do not reach: sigmoid normalisation, unweighted Dice, unweighted CE, a
    class that never occurs, the fp32 one-hot -- values and d(loss)/d(logits) from the reference's own classes."""
    torch.manual_seed(99)
    logits = (2.0 * torch.randn(2, 3, 6, 5, 4)).requires_grad_(True)
    labels = torch.randint(0, 2, (2, 6, 5, 4))                   # class 2 never occurs
    w = torch.tensor([0.2, 1.0, 0.5])
    cases = {"dice_softmax_unweighted": rloss.DiceLoss(), "dice_softmax_weighted": rloss.DiceLoss(weight=w),
             "dice_sigmoid_unweighted": rloss.DiceLoss(sigmoid_normalization=True),
             "dice_sigmoid_weighted": rloss.DiceLoss(weight=w, sigmoid_normalization=True),
             "ce_unweighted": torch.nn.CrossEntropyLoss(), "ce_weighted": torch.nn.CrossEntropyLoss(weight=w)}
    out = {"logits": logits.detach().numpy(), "labels": labels.numpy(), "weight": w.numpy(),
           "one_hot": rloss.expand_as_one_hot(labels, 3).numpy(),
           "dice_metric": rloss.dice_metric(logits.detach(), labels).numpy()}
    for name, fn in cases.items():
        value = fn(logits, labels)
        grad, = torch.autograd.grad(value, logits)
        out[name] = value.detach().numpy()
        out[name + ".dlogits"] = grad.numpy()
    np.savez_compressed(os.path.join(OUT, "loss_variants.npz"), **out)


def golden_semantics():
    """KATs probed by the survey (last row), recorded from live torch."""
    import torch.nn.functional as F
    up = F.interpolate(torch.arange(12.0).view(1, 1, 12, 1, 1), size=(25, 1, 1), mode="nearest").flatten()
    x = torch.zeros(1, 1, 2, 2, 2)
    _, idx_eq = F.max_pool3d(x, 2, return_indices=True)
    x2 = torch.zeros(1, 1, 2, 2, 2)
    x2[0, 0, 1, 0, 1] = float("nan")
    v_nan, idx_nan = F.max_pool3d(x2, 2, return_indices=True)
    am = torch.argmax(torch.tensor([[1.0, 3.0, 3.0, 2.0]]), dim=1)
    logits = torch.tensor([[[[[0.2]]], [[[1.5]]]], [[[[-0.3]]], [[[0.1]]]]])     # (2,2,1,1,1)
    lab = torch.tensor([[[[1]]], [[[0]]]])
    w = torch.tensor([0.25, 2.0])
    ce = F.cross_entropy(logits, lab, weight=w)
    np.savez_compressed(os.path.join(OUT, "semantics.npz"), nearest_12_to_25=up.numpy(),
                        pool_equal_idx=idx_eq.numpy(), pool_nan_idx=idx_nan.numpy(), pool_nan_val=v_nan.numpy(),
                        argmax_tie=am.numpy(), ce_logits=logits.numpy(), ce_label=lab.numpy(), ce_w=w.numpy(),
                        ce_value=ce.numpy())


def main():
    os.makedirs(OUT, exist_ok=True)
    rmodel, rloss = _import_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    golden_unet3d(rmodel, rloss)
    golden_unet3d_orders(rmodel)
    golden_residual(rmodel, rloss)
    golden_tiling()
    golden_semantics()
    golden_sampling()
    golden_loss_variants(rloss)
    with open(os.path.join(OUT, "PROVENANCE.txt"), "w") as f:
        f.write(f"generated by oracle/make_golden.py from {REF} with torch {torch.__version__}, numpy {np.__version__}\n")
    print("golden vectors written to", os.path.abspath(OUT))


if __name__ == "__main__":
    main()

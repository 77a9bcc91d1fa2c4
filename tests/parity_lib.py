"""Shared machinery of the bf16 parity tests (test infrastructure; imports the oracle).

Three comparators for the production (bf16 / tensor-core) UNet3D, all evaluated on the GPU so that the BASELINE.json
sizes (128^3 patches, f = 64) finish in seconds; torch's CUDA kernels are used purely as the checker's arithmetic
engine (TF32 off), exactly what the reference itself would execute on this machine:

  fp32      oracle.unet.unet3d_forward on fp32 tensors                      = the reference as shipped
  bf16-st   the same with bf16 STORAGE (oracle.unet.Storage.bf16)           = same roundings as the product, fp32 math
  native    the same functional network on `.bfloat16()` tensors           = `reference_model.bfloat16()` on cuDNN

plus the LAYER-WISE (teacher-forced) walk: every SingleConv layer / pool / join / head of the product is fed the
bf16-storage oracle's own input and output gradient of that layer and must reproduce the oracle's output, input
gradient and parameter gradients to <= 5e-3, so that an end-to-end drift can be attributed to propagation through
the network, not to a kernel.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from mednet_b200 import ops
from mednet_b200.unet.loss import DiceLoss
from oracle import loss as oloss
from oracle import unet as ounet

DEV = "cuda"


def strict_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def relerr(a, b):
    a, b = a.detach().double(), b.detach().double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def cosine(a, b):
    # not F.cosine_similarity: it clamps the norm product at 1e-8, which deep-layer gradients (norms ~1e-9) fall below
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-300)).item()


def r16(t):
    return t.to(torch.bfloat16).to(torch.float32)


def nd(t, dtype=torch.bfloat16):
    """(N,C,D,H,W) -> NDHWC contiguous in the compute dtype."""
    return t.detach().permute(0, 2, 3, 4, 1).contiguous().to(dtype)


def ncdhw(t):
    return t.detach().permute(0, 4, 1, 2, 3).float()


def learnable_batch(n, shape, classes, seed, device=DEV):
    """Synthetic image + a target the network can learn (thresholds of the smoothed image)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn((n, 1) + tuple(shape), generator=g).to(device)
    s = F.avg_pool3d(x, 3, 1, 1)[:, 0]
    y = torch.zeros_like(s, dtype=torch.long)
    for t in torch.linspace(-0.2, 0.2, classes - 1).tolist() if classes > 2 else [0.0]:
        y += (s > t).long()
    return x, y


def train_state(net, steps, shape, classes, weight, lr=2e-3, batch=1, seed0=100):
    """Brings a randomly initialised product network to a trained-like state with its own bf16 training step."""
    from mednet_b200.optim import FusedAdam
    opt = FusedAdam(net.parameters(), lr=lr)
    opt.zero_grad()
    crit = DiceLoss(weight=weight)
    losses = []
    for it in range(steps):
        x, y = learnable_batch(batch, shape, classes, seed0 + it % 8)
        loss = crit(net(x), y)
        loss.backward()
        opt.step()
        opt.zero_grad()
        if it % 10 == 0 or it == steps - 1:
            losses.append(float(loss))
    torch.cuda.synchronize()
    # hand the parameters back to ordinary autograd gradients for the parity passes that follow
    for p in net.parameters():
        p._mednet_async_grad = False
        p.grad = None
    return losses


# ------------------------------------------------------------------------------------------------ oracle passes
def oracle_pass(sd, x, y, weight, f_maps, storage=None, native=False, trace=False):
    """One forward + loss + backward of the oracle.  Returns dict(logits, loss, grads{name: fp32}, trace, tgrads)."""
    dt = torch.bfloat16 if native else torch.float32
    leaf = {k: v.detach().clone().to(dt).requires_grad_(True) for k, v in sd.items()}
    tr = [] if trace else None
    kw = {} if storage is None else {"storage": storage}
    logits = ounet.unet3d_forward(leaf, x.to(dt), f_maps=f_maps, trace=tr, **kw)
    loss = oloss.dice_loss(logits.float(), y, weight=None if weight is None else weight.to(logits.device))
    tensors, seen = [], set()
    if trace:
        for e in tr:
            for t in e[1:]:
                if torch.is_tensor(t) and t.requires_grad and id(t) not in seen:
                    seen.add(id(t))
                    tensors.append(t)
    names = list(leaf)
    g = torch.autograd.grad(loss, [leaf[k] for k in names] + tensors)
    grads = {k: v.float() for k, v in zip(names, g[:len(names)])}
    tgrads = {id(t): v for t, v in zip(tensors, g[len(names):])}
    return dict(logits=logits.detach().float(), loss=float(loss), grads=grads, trace=tr, tgrads=tgrads)


def product_pass(net, x, y, weight):
    for p in net.parameters():
        p.grad = None
    logits = net(x)
    loss = DiceLoss(weight=weight)(logits, y)
    loss.backward()
    torch.cuda.synchronize()
    return dict(logits=logits.detach(), loss=float(loss), grads={k: p.grad.detach().float().clone() for k, p in net.named_parameters()})


def end_to_end_report(net, x, y, weight, f_maps, min_numel=64):
    """Product vs the three comparators.  Returns a dict of scalars plus per-tensor cosine rows."""
    strict_fp32()
    sd = {k: v.detach().float() for k, v in net.state_dict().items()}
    ours = product_pass(net, x, y, weight)
    ref32 = oracle_pass(sd, x, y, weight, f_maps)
    st16 = oracle_pass(sd, x, y, weight, f_maps, storage=ounet.Storage.bf16(master_weights=True))
    nat = oracle_pass(sd, x, y, weight, f_maps, native=True)
    rep = {
        "logits_rel_ours_vs_fp32": relerr(ours["logits"], ref32["logits"]),
        "logits_rel_bf16storage_vs_fp32": relerr(st16["logits"], ref32["logits"]),
        "logits_rel_native_bf16_vs_fp32": relerr(nat["logits"], ref32["logits"]),
        "logits_rel_ours_vs_bf16storage": relerr(ours["logits"], st16["logits"]),
        "loss_fp32": ref32["loss"], "loss_ours": ours["loss"], "loss_bf16storage": st16["loss"], "loss_native_bf16": nat["loss"],
        "label_flips_ours_vs_fp32": float((ours["logits"].argmax(1) != ref32["logits"].argmax(1)).float().mean()),
        "label_flips_native_vs_fp32": float((nat["logits"].argmax(1) != ref32["logits"].argmax(1)).float().mean()),
        "label_flips_ours_vs_bf16storage": float((ours["logits"].argmax(1) != st16["logits"].argmax(1)).float().mean()),
    }
    rows = []
    for k, g in ref32["grads"].items():
        if g.numel() < min_numel:
            continue
        rows.append((k, cosine(ours["grads"][k], g), cosine(st16["grads"][k], g), cosine(nat["grads"][k], g), g.numel()))
    rep["grad_rows"] = rows
    rep["grad_cos_min_ours"] = min(r[1] for r in rows)
    rep["grad_cos_min_bf16storage"] = min(r[2] for r in rows)
    rep["grad_cos_min_native_bf16"] = min(r[3] for r in rows)
    rep["grad_tensors_ge_0.999_ours"] = sum(r[1] >= 0.999 for r in rows)
    rep["grad_tensors_ge_0.999_native_bf16"] = sum(r[3] >= 0.999 for r in rows)
    rep["grad_tensors"] = len(rows)
    return rep


# ------------------------------------------------------------------------------------------------ layer-wise walk
RELU = ops.ACT["r"]


def layerwise_report(net, x, y, weight, f_maps):
    """Teacher-forced parity of every layer of the 'gcr' UNet3D against the bf16-storage oracle (see module docstring).
    Returns rows (layer, quantity, relative error, cosine)."""
    strict_fp32()
    sd = {k: v.detach().float() for k, v in net.state_dict().items()}
    o = oracle_pass(sd, x, y, weight, f_maps, storage=ounet.Storage.bf16(master_weights=True), trace=True)
    tg = lambda t: o["tgrads"][id(t)]
    rows = []

    def add(layer, what, got, want, rounded=True):
        want = r16(want) if rounded else want
        rows.append((layer, what, relerr(got, want), cosine(got, want)))

    def param_rows(prefix, mod, exact=None):
        for pname, p in mod.named_parameters():
            got, want = p.grad.float(), o["grads"][prefix + pname]
            if exact is not None and pname in exact:
                rows.append((prefix, "d" + pname + " (vs unrounded dxn, f64)", relerr(got, exact[pname]), 1.0))
                rows.append((prefix, "d" + pname + " (bf16-storage oracle vs the same f64 value: its own rounding noise, not gated)",
                             relerr(want, exact[pname]), 1.0))
            elif want.numel() < 8 and pname == "groupnorm.weight" and hasattr(mod, "conv"):
                # A handful of numbers that are ~0 by construction: the next GroupNorm makes the loss invariant to the
                # scale of this layer's (ReLU) output, so dgamma = sum_v dxn * xhat = sum W (.) dW cancels to ~1e-5 of its
                # terms.  Its error is measured against the size of those terms (Cauchy-Schwarz bound ||W|| ||dW||).
                scale = (mod.conv.weight.detach().float().norm() * o["grads"][prefix + "conv.weight"].norm()).item()
                rows.append((prefix, "d" + pname + " /|W||dW|", (got - want).norm().item() / scale, 1.0))
            else:
                add(prefix, "d" + pname, got, want, rounded=False)
            p.grad = None

    after_pool = False
    join = None
    for e in o["trace"]:
        kind = e[0]
        if kind == "pool":
            _, i, xin, yout = e
            got, _ = ops.k_pool_fwd(nd(xin))
            rows.append((f"encoders.{i}.pooling", "y (exact)", float((ncdhw(got) != yout).float().mean()), 1.0))
            after_pool = True
        elif kind == "join":
            join = e
        elif kind == "layer":
            _, prefix, xin, yout = e
            mod = net.get_submodule(prefix[:-1])
            dpre = nd(tg(yout) * (yout > 0))                       # what the consumers of y hand back (deferred ReLU')
            if join is not None:                                  # decoder SingleConv1: GroupNorm over the virtual concat
                _, j, skip, low, cat = join
                s_, l_ = nd(skip).requires_grad_(), nd(low).requires_grad_()
                gn = mod.groupnorm
                if ops.upconv_supported(s_, l_, mod.conv.weight):      # the production branch of Decoder.run
                    sn, ln = ops.UpcatGroupNormSplitFn.apply(s_, l_, gn.weight, gn.bias, gn.num_groups, RELU, RELU)
                    yy = ops.UpConvJoinFn.apply(sn, ln, mod.conv.weight, RELU, True)
                else:
                    xn = ops.UpcatGroupNormFn.apply(s_, l_, gn.weight, gn.bias, gn.num_groups, RELU, RELU)
                    yy = mod.run(xn, defer=True, skip_first=True)
                add(prefix, "y", ncdhw(yy), yout)
                yy.backward(dpre)
                cs = skip.shape[1]
                gcat = tg(cat)
                add(prefix, "dskip", ncdhw(s_.grad), gcat[:, :cs] * (skip > 0))
                add(prefix, "dlow", ncdhw(l_.grad), F.avg_pool3d(gcat[:, cs:], 2) * 8.0 * (low > 0))
                join = None
            else:
                first = not xin.requires_grad                      # the image: no input gradient
                in_act = 0 if (first or after_pool) else RELU
                xi = nd(xin)
                if not first:
                    xi.requires_grad_()
                yy = mod.run(xi, in_act=in_act, defer=True)
                add(prefix, "y", ncdhw(yy), yout)
                yy.backward(dpre)
                if not first:
                    gx = tg(xin)
                    add(prefix, "dx", ncdhw(xi.grad), gx * (xin > 0) if in_act else gx)
                else:
                    # One-channel GroupNorm of the image: dbeta = sum_v dxn[v] is ONE number, a sum of 2 M signed terms.  The
                    # bf16-storage oracle rounds dxn to bf16 before summing (noise ~2^-9 ||dxn||_2 / sqrt 3: 1e-4 .. 1e-2 of
                    # the sum, depending on how much of it cancels in the state at hand -- 1.5e-4, 1.4e-3 and 6.9e-3 were
                    # seen on three trained states); the product never materialises dxn, so it has no such
                    # rounding and is compared with the unrounded evaluation of the same formula in float64.
                    g64 = (tg(yout) * (yout > 0)).double()
                    wq = r16(mod.conv.weight.detach().float()).double()
                    dxn = torch.nn.grad.conv3d_input(list(xin.shape), wq, g64, padding=1)
                    first_exact = {"groupnorm.bias": dxn.sum(dim=(0, 2, 3, 4)).float()}
                    del g64, dxn
                    param_rows(prefix, mod, first_exact)
                    after_pool = False
                    continue
            param_rows(prefix, mod)
            after_pool = False
        elif kind == "final":
            _, xin, logits = e
            xi = nd(xin).requires_grad_()
            fc = net.final_conv
            got = ops.Conv1x1Fn.apply(xi, fc.weight, fc.bias, RELU)
            add("final_conv.", "logits", got, logits, rounded=False)
            lg = logits.detach().clone().requires_grad_()
            loss = DiceLoss(weight=weight)(lg, y)
            rows.append(("loss", "value (abs diff)", abs(float(loss) - o["loss"]), 1.0))
            loss.backward()
            add("loss", "dlogits", lg.grad, tg(logits), rounded=False)
            got.backward(tg(logits).contiguous())
            add("final_conv.", "dx", ncdhw(xi.grad), tg(xin) * (xin > 0))
            param_rows("final_conv.", fc)
    torch.cuda.synchronize()
    return rows

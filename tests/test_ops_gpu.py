"""Op-level parity of every CUDA kernel (through the C ABI) against the arithmetic engine of the reference:
the matching torch.nn.functional call on CPU fp32.  Tolerances: fp32 kernels
1e-4 relative; index / label / uint8 outputs bit-exact."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from mednet_b200 import heatmaps as hm
from mednet_b200 import ops
from oracle import heatmaps as ohm
from oracle import loss as oloss
from oracle import tiling as otiling

pytestmark = pytest.mark.gpu
DEV = "cuda"


def ndhwc(x, dtype=torch.float32):          # (N,C,D,H,W) cpu -> (N,D,H,W,C) cuda
    return x.permute(0, 2, 3, 4, 1).contiguous().to(DEV, dtype)


def ncdhw(y):
    return y.float().permute(0, 4, 1, 2, 3).contiguous().cpu()


def relerr(a, b):
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def test_layout_roundtrip():
    x = torch.randn(2, 5, 3, 4, 6)
    y = ops.k_layout(x.to(DEV), True, torch.float32)
    assert torch.equal(y.cpu(), x.permute(0, 2, 3, 4, 1).contiguous())
    z = ops.k_layout(y, False, torch.float32)
    assert torch.equal(z.cpu(), x)
    yb = ops.k_layout(x.to(DEV), True, torch.bfloat16)
    assert torch.equal(yb.cpu(), x.permute(0, 2, 3, 4, 1).contiguous().bfloat16())
    x1 = torch.randn(2, 1, 4, 4, 4)
    assert torch.equal(ops.k_layout(x1.to(DEV), True, torch.float32).cpu().flatten(), x1.flatten())


@pytest.mark.parametrize("cin,cout,shape", [(1, 4, (5, 6, 7)), (3, 8, (4, 9, 5)), (8, 20, (6, 6, 6)), (16, 16, (8, 8, 8)),
                                            (20, 7, (3, 4, 5))])
@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_conv3_simt_fp32_fwd_bwd(cin, cout, shape, act):
    torch.manual_seed(cin * 100 + cout)
    x = torch.randn(2, cin, *shape, requires_grad=True)
    w = (torch.randn(cout, cin, 3, 3, 3) * 0.2).requires_grad_()
    b = torch.randn(cout, requires_grad=True)
    add = torch.randn(2, cout, *shape, requires_grad=True)
    pre = F.conv3d(x, w, b, padding=1) + add
    ref = [pre, F.relu(pre), F.leaky_relu(pre, 0.1), F.elu(pre)][act]
    g = torch.randn_like(ref)
    ref.backward(g)
    xg = ndhwc(x.detach()).requires_grad_()
    wg, bg = w.detach().to(DEV).requires_grad_(), b.detach().to(DEV).requires_grad_()
    ag = ndhwc(add.detach()).requires_grad_()
    y = ops.Conv3x3Fn.apply(xg, wg, bg, ag, act, "simt")
    y.backward(ndhwc(g))
    assert relerr(ncdhw(y.detach()), ref.detach()) < 1e-5
    assert relerr(ncdhw(xg.grad), x.grad) < 1e-5
    assert relerr(wg.grad.cpu(), w.grad) < 1e-5
    assert relerr(bg.grad.cpu(), b.grad) < 1e-5
    assert relerr(ncdhw(ag.grad), add.grad) < 1e-5


@pytest.mark.parametrize("cin,cout,shape", [(1, 32, (5, 9, 70)), (1, 8, (3, 4, 5)), (2, 16, (4, 6, 7)), (4, 64, (2, 5, 66)),
                                            (1, 24, (4, 4, 9)), (1, 16, (3, 5, 6)), (1, 64, (2, 3, 133)), (3, 16, (3, 4, 3)),
                                            (1, 32, (3, 3, 2))])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_first_layer_small_channel_kernels(cin, cout, shape, dtype):
    """in_channels = 1 layers (segmentation.py:30-31): few-input-channel fprop, few-output-channel dgrad and the
    small-channel wgrad specialisations of the CUDA-core path (w-segments longer than 64 voxels included)."""
    torch.manual_seed(cin * 10 + cout)
    q = (lambda t: t.to(dtype).float())
    x = q(torch.randn(2, cin, *shape)).requires_grad_()
    w = q(torch.randn(cout, cin, 3, 3, 3) * 0.2).requires_grad_()
    ref = F.elu(F.conv3d(x, w, None, padding=1))
    g = q(torch.randn_like(ref))
    ref.backward(g)
    xg = ndhwc(x.detach(), dtype).requires_grad_()
    wg = w.detach().to(DEV).requires_grad_()
    y = ops.Conv3x3Fn.apply(xg, wg, None, None, 3, "auto")
    y.backward(ndhwc(g, dtype))
    tol = 1e-5 if dtype == torch.float32 else 6e-3
    assert relerr(ncdhw(y.detach()), ref.detach()) < tol
    assert relerr(ncdhw(xg.grad), x.grad) < (1e-5 if dtype == torch.float32 else 2e-2)
    assert relerr(wg.grad.cpu(), w.grad) < (1e-5 if dtype == torch.float32 else 1e-2)


def test_conv3_simt_bf16_matches_bf16_rounded_reference():
    torch.manual_seed(3)
    x = torch.randn(1, 16, 6, 7, 8).bfloat16().float()
    w = (torch.randn(24, 16, 3, 3, 3) * 0.1).bfloat16().float()
    ref = F.relu(F.conv3d(x, w, None, padding=1))
    y = ops.Conv3x3Fn.apply(ndhwc(x, torch.bfloat16), w.to(DEV), None, None, 1, "simt")
    assert relerr(ncdhw(y), ref) < 4e-3


@pytest.mark.parametrize("cin,cout,shape", [(4, 6, (3, 4, 5)), (16, 8, (4, 4, 4)), (5, 3, (2, 3, 2))])
def test_conv_transpose_fwd_bwd(cin, cout, shape):
    torch.manual_seed(cin + cout)
    x = torch.randn(2, cin, *shape, requires_grad=True)
    w = (torch.randn(cin, cout, 3, 3, 3) * 0.2).requires_grad_()
    b = torch.randn(cout, requires_grad=True)
    skip = torch.randn(2, cout, *[2 * s for s in shape], requires_grad=True)
    ref = F.conv_transpose3d(x, w, b, stride=2, padding=1, output_padding=1) + skip
    g = torch.randn_like(ref)
    ref.backward(g)
    xg = ndhwc(x.detach()).requires_grad_()
    wg, bg = w.detach().to(DEV).requires_grad_(), b.detach().to(DEV).requires_grad_()
    sg = ndhwc(skip.detach()).requires_grad_()
    y = ops.ConvTranspose3x3Fn.apply(xg, wg, bg, sg, "auto")
    y.backward(ndhwc(g))
    assert relerr(ncdhw(y.detach()), ref.detach()) < 1e-5
    assert relerr(ncdhw(xg.grad), x.grad) < 1e-5
    assert relerr(wg.grad.cpu(), w.grad) < 1e-5
    assert relerr(bg.grad.cpu(), b.grad) < 1e-5
    assert torch.equal(ncdhw(sg.grad), g)


@pytest.mark.parametrize("c,groups,shape", [(1, 1, (9, 7, 5)), (8, 8, (4, 4, 4)), (24, 8, (5, 6, 7)), (192, 8, (4, 4, 6)),
                                            (20, 4, (3, 5, 2)), (1024, 8, (2, 2, 2))])
@pytest.mark.parametrize("act,res", [(0, False), (1, False), (3, True), (2, True)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_groupnorm_act_fwd_bwd(c, groups, shape, act, res, dtype):
    torch.manual_seed(c + act)
    q = (lambda t: t.to(dtype).float())
    x = q(torch.randn(2, c, *shape) * 2 + 0.5).requires_grad_()
    gamma = (torch.rand(c) + 0.5).requires_grad_()
    beta = torch.randn(c).requires_grad_()
    r = q(torch.randn(2, c, *shape)).requires_grad_() if res else None
    pre = F.group_norm(x, groups, gamma, beta, 1e-5)
    if res:
        pre = pre + r
    ref = [pre, F.relu(pre), F.leaky_relu(pre, 0.1), F.elu(pre)][act]
    g = q(torch.randn_like(ref))
    ref.backward(g)
    xg = ndhwc(x.detach(), dtype).requires_grad_()
    gg, bg = gamma.detach().to(DEV).requires_grad_(), beta.detach().to(DEV).requires_grad_()
    rg = ndhwc(r.detach(), dtype).requires_grad_() if res else None
    y = ops.GroupNormActFn.apply(xg, gg, bg, groups, act, rg)
    y.backward(ndhwc(g, dtype))
    tol = 1e-5 if dtype == torch.float32 else 1.5e-2
    assert relerr(ncdhw(y.detach()), ref.detach()) < tol
    assert relerr(ncdhw(xg.grad), x.grad) < (2e-4 if dtype == torch.float32 else 3e-2)
    assert relerr(gg.grad.cpu(), gamma.grad) < (1e-4 if dtype == torch.float32 else 3e-2)
    assert relerr(bg.grad.cpu(), beta.grad) < (1e-4 if dtype == torch.float32 else 3e-2)
    if res:
        assert relerr(ncdhw(rg.grad), r.grad) < tol


@pytest.mark.parametrize("act", [1, 2, 3])
def test_standalone_activation(act):
    x = torch.randn(3, 4, 5, 6, 7, requires_grad=True)
    ref = [None, F.relu(x), F.leaky_relu(x, 0.1), F.elu(x)][act]
    g = torch.randn_like(ref)
    ref.backward(g)
    xg = x.detach().to(DEV).requires_grad_()
    y = ops.ActFn.apply(xg, act)
    y.backward(g.to(DEV))
    assert relerr(y.detach().cpu(), ref.detach()) < 1e-6 and relerr(xg.grad.cpu(), x.grad) < 1e-6


@pytest.mark.parametrize("c,shape", [(1, (4, 4, 4)), (8, (6, 8, 10)), (12, (5, 7, 9)), (64, (4, 4, 4)), (3, (2, 3, 2))])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_maxpool_values_indices_and_backward_bit_exact(c, shape, dtype):
    torch.manual_seed(c)
    x = torch.randint(-3, 4, (2, c, *shape)).float()          # many ties -> exercises the first-max rule
    x[0, 0, 1, 0, 1] = float("nan")                          # NaN wins
    ref, ridx = F.max_pool3d(x, 2, return_indices=True)
    xg = ndhwc(x, dtype).requires_grad_()
    y = ops.MaxPoolFn.apply(xg)
    yi, idx = ops.k_pool_fwd(xg.detach())
    assert torch.equal(torch.nan_to_num(ncdhw(y.detach()), nan=123.0), torch.nan_to_num(ref, nan=123.0))
    assert torch.equal(ops.k_pool_indices_i64(idx, tuple(xg.shape)).cpu(), ridx)
    g = torch.randn_like(ref).to(dtype).float()
    xr = x.clone().requires_grad_()
    F.max_pool3d(xr, 2).backward(g)
    y.backward(ndhwc(g, dtype))
    assert torch.equal(ncdhw(xg.grad), xr.grad)


@pytest.mark.parametrize("cs,cl,big,small", [(8, 16, (8, 8, 8), (4, 4, 4)), (4, 4, (25, 7, 9), (12, 3, 4)),
                                             (64, 128, (4, 6, 4), (2, 3, 2)), (3, 5, (5, 5, 5), (2, 2, 2)),
                                             (8, 8, (13, 14, 15), (6, 7, 7))])
def test_upsample_concat_fwd_bwd_exact(cs, cl, big, small):
    torch.manual_seed(cs)
    skip = torch.randn(2, cs, *big, requires_grad=True)
    low = torch.randn(2, cl, *small, requires_grad=True)
    ref = torch.cat((skip, F.interpolate(low, size=big, mode="nearest")), dim=1)
    g = torch.randint(-4, 5, ref.shape).float()              # integer grads -> exact sums
    ref.backward(g)
    sg, lg = ndhwc(skip.detach()).requires_grad_(), ndhwc(low.detach()).requires_grad_()
    y = ops.UpsampleConcatFn.apply(sg, lg)
    y.backward(ndhwc(g))
    assert torch.equal(ncdhw(y.detach()), ref.detach())
    assert torch.equal(ncdhw(sg.grad), skip.grad) and torch.equal(ncdhw(lg.grad), low.grad)


@pytest.mark.parametrize("cin,cout", [(8, 2), (64, 4), (32, 10), (20, 3)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_final_conv1x1(cin, cout, dtype):
    torch.manual_seed(cin)
    x = (torch.randn(2, cin, 5, 6, 7)).to(dtype).float().requires_grad_()
    w = (torch.randn(cout, cin, 1, 1, 1) * 0.3).requires_grad_()
    b = torch.randn(cout, requires_grad=True)
    ref = F.conv3d(x, w, b)
    g = torch.randn_like(ref)
    ref.backward(g)
    xg = ndhwc(x.detach(), dtype).requires_grad_()
    wg, bg = w.detach().to(DEV).requires_grad_(), b.detach().to(DEV).requires_grad_()
    y = ops.Conv1x1Fn.apply(xg, wg, bg)
    assert y.dtype == torch.float32 and y.shape == ref.shape
    y.backward(g.to(DEV))
    assert relerr(y.detach().cpu(), ref.detach()) < 1e-5
    assert relerr(ncdhw(xg.grad), x.grad) < (1e-5 if dtype == torch.float32 else 5e-3)
    assert relerr(wg.grad.cpu(), w.grad) < 1e-5 and relerr(bg.grad.cpu(), b.grad) < 1e-5


@pytest.mark.parametrize("c", [2, 4, 7, 12])
@pytest.mark.parametrize("label_dtype", [torch.int64, torch.uint8])
def test_dice_and_ce_losses(c, label_dtype):
    torch.manual_seed(c)
    logits = (torch.randn(2, c, 6, 7, 8) * 2).requires_grad_()
    labels = torch.randint(0, c, (2, 6, 7, 8))
    w = torch.rand(c) + 0.05
    for name, fn_ref, fn_mine in (
            ("dice", lambda z: oloss.dice_loss(z, labels, weight=w),
             lambda z: ops.DiceLossFn.apply(z, labels.to(DEV, label_dtype), w.to(DEV), 1e-5, False)[0]),
            ("dice_sigmoid", lambda z: oloss.dice_loss(z, labels, weight=None, sigmoid_normalization=True),
             lambda z: ops.DiceLossFn.apply(z, labels.to(DEV, label_dtype), None, 1e-5, True)[0]),
            ("ce", lambda z: oloss.weighted_cross_entropy(z, labels, w),
             lambda z: ops.CrossEntropyFn.apply(z, labels.to(DEV, label_dtype), w.to(DEV)))):
        logits.grad = None
        ref = fn_ref(logits)
        ref.backward()
        zg = logits.detach().to(DEV).requires_grad_()
        mine = fn_mine(zg)
        (mine * 1.0).backward()
        assert abs(mine.item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item())), name
        assert relerr(zg.grad.cpu(), logits.grad) < 1e-4, name
    dm = ops.k_dice_fwd(logits.detach().to(DEV), labels.to(DEV), None, 1e-5, False)[1]
    np.testing.assert_allclose(dm.cpu().numpy(), oloss.dice_metric(logits.detach(), labels).numpy(), rtol=1e-5)


def test_dice_on_channel_slice_and_fused_landmark_loss():
    torch.manual_seed(5)
    L, K = 3, 2
    out = (torch.randn(2, L + K, 6, 6, 6) * 3).requires_grad_()
    labels = torch.randint(0, K, (2, 6, 6, 6))
    hmaps = torch.randint(0, 256, (2, L, 6, 6, 6), dtype=torch.uint8)
    cw, rw = torch.tensor([0.05, 1.0]), [0.001, 0.015, 0.02]
    for lc, lr in (("DICE", "L2"), ("CE", "L1")):
        out.grad = None
        tot, cl, rg = oloss.landmark_loss(out[:, L:], out[:, :L], labels, hmaps.float(), cw, rw, lc, lr)
        tot.backward()
        og = out.detach().to(DEV).requires_grad_()
        t2, c2, r2 = ops.LandmarkLossFn.apply(og, labels.to(DEV, torch.uint8), hmaps.to(DEV), cw.to(DEV),
                                              torch.tensor(rw).to(DEV), lc == "CE", lr == "L1", 1e-5)
        t2.backward()
        assert abs(t2.item() - tot.item()) < 1e-4 * abs(tot.item()) and abs(c2.item() - cl.item()) < 1e-5
        assert abs(r2.item() - rg.item()) < 1e-4 * abs(rg.item())
        assert relerr(og.grad.cpu(), out.grad) < 1e-4
        # un-fused path on slices goes through autograd's slice backward and must agree
        og2 = out.detach().to(DEV).requires_grad_()
        d = ops.DiceLossFn.apply(og2[:, L:], labels.to(DEV), cw.to(DEV), 1e-5, False)[0] if lc == "DICE" else \
            ops.CrossEntropyFn.apply(og2[:, L:], labels.to(DEV), cw.to(DEV))
        h, _ = ops.HeatmapLossFn.apply(og2[:, :L], hmaps.to(DEV), torch.tensor(rw).to(DEV), lr == "L1")
        (d + h).backward()
        assert relerr(og2.grad.cpu(), out.grad) < 1e-4


def test_predict_epilogue_bit_exact():
    torch.manual_seed(9)
    logits = torch.randn(3, 7, 5, 6, 7) * 120
    logits[0, 4:, 0, 0, 0] = 1.5                                  # tie -> first index
    want = otiling.predict_epilogue(logits.numpy(), 4)
    got = ops.k_predict_epilogue(logits.to(DEV), 4).cpu().numpy()
    np.testing.assert_array_equal(got, want)


def test_final_activation():
    z = torch.randn(2, 5, 3, 4, 5)
    np.testing.assert_allclose(ops.k_final_activation(z.to(DEV), False).cpu().numpy(), torch.softmax(z, 1).numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(ops.k_final_activation(z.to(DEV), True).cpu().numpy(), torch.sigmoid(z).numpy(), rtol=1e-5, atol=1e-7)


def test_fused_adam_matches_torch():
    torch.manual_seed(1)
    from mednet_b200.optim import FusedAdam
    ps = [torch.randn(33, 7), torch.randn(129), torch.randn(4, 4, 3, 3, 3)]
    ref = [p.clone().requires_grad_() for p in ps]
    mine = [torch.nn.Parameter(p.clone().to(DEV)) for p in ps]
    o_ref, o_mine = torch.optim.Adam(ref, lr=1e-2), FusedAdam(mine, lr=1e-2)
    for it in range(5):
        o_ref.zero_grad()
        o_mine.zero_grad()
        gs = [torch.randn_like(p) for p in ps]
        for p, g in zip(ref, gs):
            p.grad = g.clone()
        for p, g in zip(mine, gs):
            p.grad.copy_(g.to(DEV)) if it % 2 == 0 else setattr(p, "grad", g.to(DEV))   # both write paths
        o_ref.step()
        o_mine.step()
    for a, b in zip(mine, ref):
        assert relerr(a.detach().cpu(), b.detach()) < 1e-5


def test_heatmap_render_and_landmark_extraction():
    torch.manual_seed(2)
    pts = torch.rand(2, 3, 3) * torch.tensor([10.0, 12.0, 14.0])
    sig = [1.5, 2.0, 3.0]
    got = hm.render_heatmaps(pts.to(DEV), sig, (10, 12, 14)).cpu().numpy()
    want = ohm.render_heatmaps(pts.numpy(), sig, (10, 12, 14))
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1 and (got != want).mean() < 2e-3   # expf ulp at truncation edges
    x = torch.randint(0, 5, (2, 3, 6, 7, 8)).float()             # ties everywhere -> first-max rule
    arg, soft = hm.extract_landmarks(x.to(DEV), soft=True, beta=0.7)
    assert torch.equal(arg.cpu(), ohm.argmax_landmarks(x))
    np.testing.assert_allclose(soft.cpu().numpy(), ohm.soft_argmax_landmarks(x, 0.7).numpy(), rtol=1e-4, atol=1e-4)
    u8 = torch.from_numpy(want)
    assert torch.equal(hm.extract_landmarks(u8.to(DEV)).cpu(), ohm.argmax_landmarks(u8))


# ---------------------------------------------------------------------------------------------------------------
# Deferred activation derivative (include/mednet_b200.h): conv(+act) with defer_act=True followed by a consumer that
# is told `in_act` must give the same gradients as the reference chain  act(conv(x)) -> consumer.
# ---------------------------------------------------------------------------------------------------------------
def _act_ref(pre, act):
    return [pre, F.relu(pre), F.leaky_relu(pre, 0.1), F.elu(pre)][act]


@pytest.mark.parametrize("act", [1, 2, 3])
@pytest.mark.parametrize("consumer", ["groupnorm", "maxpool", "upcat_skip", "upcat_low", "conv1x1"])
def test_deferred_activation_derivative(act, consumer):
    torch.manual_seed(act)
    cin, c = 6, 8
    x = torch.randn(2, cin, 6, 8, 4, requires_grad=True)
    w = (torch.randn(c, cin, 3, 3, 3) * 0.2).requires_grad_()
    y = _act_ref(F.conv3d(x, w, None, padding=1), act)
    xg = ndhwc(x.detach()).requires_grad_()
    wg = w.detach().to(DEV).requires_grad_()
    yg = ops.Conv3x3Fn.apply(xg, wg, None, None, act, "simt", True)
    extra_ref, extra_got = [], []
    if consumer == "groupnorm":
        gamma, beta = (torch.rand(c) + 0.5).requires_grad_(), torch.randn(c).requires_grad_()
        out = F.group_norm(y, 4, gamma, beta, 1e-5)
        gg, bg = gamma.detach().to(DEV).requires_grad_(), beta.detach().to(DEV).requires_grad_()
        got = ncdhw_keep(ops.GroupNormActFn.apply(yg, gg, bg, 4, 0, None, act))
        extra_ref, extra_got = [gamma, beta], [gg, bg]
    elif consumer == "maxpool":
        out = F.max_pool3d(y, 2)
        got = ncdhw_keep(ops.MaxPoolFn.apply(yg, act))
    elif consumer == "upcat_skip":
        low = torch.randn(2, 4, 3, 4, 2)
        out = torch.cat((y, F.interpolate(low, size=y.shape[2:], mode="nearest")), dim=1)
        got = ncdhw_keep(ops.UpsampleConcatFn.apply(yg, ndhwc(low), act, 0))
    elif consumer == "upcat_low":
        skip = torch.randn(2, 4, 12, 16, 8)
        out = torch.cat((skip, F.interpolate(y, size=skip.shape[2:], mode="nearest")), dim=1)
        got = ncdhw_keep(ops.UpsampleConcatFn.apply(ndhwc(skip), yg, 0, act))
    else:
        w1, b1 = (torch.randn(3, c, 1, 1, 1) * 0.3).requires_grad_(), torch.randn(3, requires_grad=True)
        out = F.conv3d(y, w1, b1)
        w1g, b1g = w1.detach().to(DEV).requires_grad_(), b1.detach().to(DEV).requires_grad_()
        got = ops.Conv1x1Fn.apply(yg, w1g, b1g, act)
        extra_ref, extra_got = [w1, b1], [w1g, b1g]
    g = torch.randn_like(out)
    out.backward(g)
    got.backward(g.to(DEV))
    assert relerr(got.detach().cpu(), out.detach()) < 1e-5
    assert relerr(ncdhw(xg.grad), x.grad) < 1e-4
    assert relerr(wg.grad.cpu(), w.grad) < 1e-4
    for a, b in zip(extra_ref, extra_got):
        assert relerr(b.grad.cpu(), a.grad) < 1e-4


def ncdhw_keep(y):                           # (N,D,H,W,C) cuda -> (N,C,D,H,W) view that keeps the autograd graph
    return y.permute(0, 4, 1, 2, 3)


@pytest.mark.parametrize("cs,cl,groups,small", [(8, 16, 8, (3, 4, 2)), (64, 128, 8, (2, 2, 3)), (16, 8, 4, (4, 3, 5)),
                                                (4, 4, 1, (2, 2, 2)), (6, 10, 8, (2, 3, 2))])
@pytest.mark.parametrize("acts", [(0, 0), (1, 1), (3, 2)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_groupnorm_over_virtual_concat(cs, cl, groups, small, acts, dtype):
    """GroupNorm(cat(skip, up2(low))) without the concat tensor == F.interpolate + cat + group_norm of the reference
    (components.py:277-280, :57), including groups that straddle the skip/low boundary (192 = 64 + 128 channels in
    groups of 24) and the deferred activation derivatives of both producers."""
    torch.manual_seed(cs + cl)
    q = (lambda t: t.to(dtype).float())
    big = tuple(2 * v for v in small)
    C = cs + cl
    sp = q(torch.randn(2, cs, *big) * 1.5).requires_grad_()      # pre-activations of the two producers
    lp = q(torch.randn(2, cl, *small) + 0.3).requires_grad_()
    skip, low = q(_act_ref(sp, acts[0])), q(_act_ref(lp, acts[1]))
    # straight-through for the storage rounding of the activated tensors
    skip = _act_ref(sp, acts[0]) + (skip - _act_ref(sp, acts[0])).detach()
    low = _act_ref(lp, acts[1]) + (low - _act_ref(lp, acts[1])).detach()
    gamma, beta = (torch.rand(C) + 0.5).requires_grad_(), torch.randn(C).requires_grad_()
    ref = F.group_norm(torch.cat((skip, F.interpolate(low, size=big, mode="nearest")), dim=1), groups, gamma, beta, 1e-5)
    g = q(torch.randn_like(ref))
    ref.backward(g)
    sg, lg = ndhwc(skip.detach(), dtype).requires_grad_(), ndhwc(low.detach(), dtype).requires_grad_()
    gg, bg = gamma.detach().to(DEV).requires_grad_(), beta.detach().to(DEV).requires_grad_()
    assert ops.upcat_gn_supported(sg, lg)
    y = ops.UpcatGroupNormFn.apply(sg, lg, gg, bg, groups, acts[0], acts[1])
    y.backward(ndhwc(g, dtype))
    f32 = dtype == torch.float32
    assert relerr(ncdhw(y.detach()), ref.detach()) < (1e-5 if f32 else 1.5e-2)
    assert relerr(ncdhw(sg.grad), sp.grad) < (2e-4 if f32 else 3e-2)
    assert relerr(ncdhw(lg.grad), lp.grad) < (2e-4 if f32 else 3e-2)
    assert relerr(gg.grad.cpu(), gamma.grad) < (1e-4 if f32 else 3e-2)
    assert relerr(bg.grad.cpu(), beta.grad) < (1e-4 if f32 else 3e-2)
    # identical to the unfused composition of the same kernels (upsample_concat -> groupnorm)
    y2 = ops.GroupNormActFn.apply(ops.UpsampleConcatFn.apply(sg.detach(), lg.detach()), gg.detach(), bg.detach(), groups, 0, None)
    assert relerr(y.detach().float().cpu(), y2.float().cpu()) < (1e-6 if f32 else 8e-3)
    assert not ops.upcat_gn_supported(sg, lg[:, :1])             # non-2x geometry -> unfused path


@pytest.mark.parametrize("c,shape", [(8, (6, 8, 10)), (12, (5, 7, 9)), (64, (4, 4, 4))])
@pytest.mark.parametrize("act", [0, 1])
def test_maxpool_with_fused_skip_gradient(c, shape, act):
    """MaxPoolSkipFn returns (pooled, x) and adds the skip path's gradient inside the pool-backward kernel: same
    gradient as autograd's pool backward + accumulate (components.py:210,224 + the encoder features of model.py:90-95)."""
    torch.manual_seed(c)
    pre = torch.randn(2, c, *shape, requires_grad=True)
    x = _act_ref(pre, act)
    out = F.max_pool3d(x, 2)
    g1, g2 = torch.randn_like(out), torch.randn_like(x)
    (out * g1).sum().backward(retain_graph=True)
    gp = pre.grad.clone()
    pre.grad = None
    ((out * g1).sum() + (x * g2).sum()).backward()
    xg = ndhwc(x.detach()).requires_grad_()
    y, skip = ops.MaxPoolSkipFn.apply(xg, act)
    assert torch.equal(ncdhw(y.detach()), out.detach()) and torch.equal(skip.detach(), xg.detach())
    # the skip consumer applies the deferred mask itself (here: by hand), exactly like the join kernels do
    mask = ndhwc((pre.detach() > 0).float() if act else torch.ones_like(pre))
    ((y * ndhwc(g1)).sum() + (skip * (ndhwc(g2) * mask)).sum()).backward()
    assert relerr(ncdhw(xg.grad), pre.grad) < 1e-6
    xg2 = ndhwc(x.detach()).requires_grad_()
    y2, _ = ops.MaxPoolSkipFn.apply(xg2, act)
    (y2 * ndhwc(g1)).sum().backward()                       # skip branch unused -> plain pool backward
    assert relerr(ncdhw(xg2.grad), gp) < 1e-6


@pytest.mark.parametrize("c,groups", [(1, 1), (16, 8)])
def test_groupnorm_large_offset_input_matches_welford(c, groups):
    """Un-normalised intensities (|mean| >> std, e.g. CT values) reach the first GroupNorm of the 'gcr' order
    (components.py:45-57).  torch's GroupNorm is Welford-based; raw fp32 sums of x and x^2 would cancel catastrophically
    here (variance 1 out of second moments ~1e6).  The kernels accumulate pivot-shifted sums: same 1e-4 as elsewhere."""
    torch.manual_seed(0)
    x = (1000.0 + torch.randn(2, c, 12, 20, 24)).requires_grad_()
    gamma, beta = torch.rand(c) + 0.5, torch.randn(c)
    ref = F.group_norm(x, groups, gamma, beta, eps=1e-5)
    xg = ndhwc(x.detach())
    y = ops.GroupNormActFn.apply(xg, gamma.to(DEV), beta.to(DEV), groups, 0, None)
    assert relerr(ncdhw(y), ref.detach()) < 1e-4
    # virtual-concat variant: skip and (upsampled) low with different large offsets
    if c >= 16:
        skip = 500.0 + torch.randn(1, c, 8, 8, 8)
        low = -300.0 + torch.randn(1, c, 4, 4, 4)
        cat = torch.cat((skip, F.interpolate(low, size=(8, 8, 8), mode="nearest")), dim=1)
        g2, b2 = torch.rand(2 * c) + 0.5, torch.randn(2 * c)
        ref2 = F.group_norm(cat, groups, g2, b2, eps=1e-5)
        y2 = ops.UpcatGroupNormFn.apply(ndhwc(skip), ndhwc(low), g2.to(DEV), b2.to(DEV), groups, 0, 0)
        assert relerr(ncdhw(y2), ref2) < 1e-4


@pytest.mark.parametrize("cin,cout,groups,shape", [(1, 32, 1, (6, 9, 20)), (1, 16, 1, (2, 2, 2)), (2, 8, 1, (5, 4, 7)),
                                                   (4, 64, 2, (3, 6, 5))])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_first_layer_without_data_gradient(cin, cout, groups, shape, dtype):
    """GroupNorm -> conv -> ReLU on the image (no input gradient): dgamma / dbeta / dW from the adjoint identity
    (ops.NormConvInputFn: no dgrad, no GroupNorm backward) against torch autograd of the reference ops
    (components.py:45-57, :8-9), with perturbed affine parameters and every border class populated."""
    torch.manual_seed(cin + cout)
    q = (lambda t: t.to(dtype).float()) if dtype == torch.bfloat16 else (lambda t: t)
    x = q(torch.randn(2, cin, *shape) * 2.0 + 3.0)                 # the image as the network stores it
    gamma = (torch.rand(cin) + 0.5).requires_grad_()
    beta = (torch.randn(cin) * 0.5).requires_grad_()
    w = (torch.randn(cout, cin, 3, 3, 3) * 0.2).requires_grad_()
    xn = F.group_norm(x, groups, gamma, beta, eps=1e-5)
    xn_q = xn + (q(xn) - xn).detach()                              # stored in the compute dtype, fp32 gradient
    wq = w + (q(w) - w).detach()
    ref = F.relu(F.conv3d(xn_q, wq, None, padding=1))
    g = q(torch.randn_like(ref))
    ref.backward(g)
    xg = ndhwc(x, dtype)
    assert not xg.requires_grad
    gd, bd, wd = (t.detach().to(DEV).requires_grad_() for t in (gamma, beta, w))
    assert ops.norm_conv_input_supported(xg, wd, None)
    y = ops.NormConvInputFn.apply(xg, gd, bd, groups, wd, 1, "auto")
    y.backward(ndhwc(g, dtype))
    tol, gtol = (2e-5, 2e-5) if dtype == torch.float32 else (6e-3, 2e-2)     # bf16: one rounding of xhat vs of xn per voxel
    assert relerr(ncdhw(y.detach()), ref.detach()) < tol
    assert relerr(wd.grad.cpu(), w.grad) < gtol
    assert relerr(gd.grad.cpu(), gamma.grad) < gtol
    assert relerr(bd.grad.cpu(), beta.grad) < gtol

"""tcgen05/TMEM implicit-GEMM convolution: descriptor calibration probe and parity against F.conv3d
(CPU fp32 on bf16-rounded operands).  Tolerance: relative error <= 1e-2 (north star, bf16 mode); with
fp32 accumulation and exact bf16 products the observed error is set by the bf16 output rounding (~3e-3)."""
import pytest
import torch
import torch.nn.functional as F

from mednet_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def relerr(a, b):
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def test_probe_finds_a_working_descriptor_variant():
    v = ops.calibrate_tcgen05(force=True)
    rep = v["report"]
    print({k: val for k, val in rep.items()})
    for rb in (128, 64, 32):
        assert rep[f"rb{rb}.aligned.bo0"], "canonical aligned K-major descriptor must address correctly"
        assert v[rb]["enabled"] == 1, f"no shifted-window variant works for {rb}-byte rows: {rep}"


CASES = [
    # N, K(Cin), Nout, (D,H,W)
    (1, 64, 64, (2, 16, 8)),        # exactly one brick, one chunk
    (2, 64, 64, (5, 20, 12)),       # ragged tiles in all three axes, odd depth
    (1, 32, 32, (4, 16, 16)),       # 64-byte rows (SW64)
    (1, 16, 32, (3, 9, 9)),         # 32-byte rows (SW32)
    (1, 96, 32, (4, 16, 8)),        # 3 chunks of 32
    (1, 192, 64, (4, 16, 16)),      # decoder shape class 192 -> 64 (3 chunks of 64)
    (1, 64, 128, (4, 16, 8)),
    (1, 128, 256, (2, 16, 8)),      # Ntile = 256: single accumulator stage
    (1, 64, 512, (2, 8, 8)),        # Nout split into two N tiles
    (3, 64, 48, (6, 18, 10)),       # N not a power of two
]


@pytest.mark.parametrize("n,k,nout,shape", CASES)
def test_conv3_tcgen05_matches_conv3d(n, k, nout, shape):
    torch.manual_seed(k + nout)
    x = torch.randn(n, k, *shape).bfloat16()
    w = (torch.randn(nout, k, 3, 3, 3) / (27 * k) ** 0.5).bfloat16()
    b = torch.randn(nout)
    ref = F.relu(F.conv3d(x.float(), w.float(), b, padding=1))
    xg = x.permute(0, 2, 3, 4, 1).contiguous().to(DEV)
    impl = ops.conv_select_impl(xg.shape, shape, k, nout, torch.bfloat16, 0, "tcgen05", xg.data_ptr())
    assert impl == 2
    y = ops.Conv3x3Fn.apply(xg, w.float().to(DEV), b.to(DEV), None, 1, "tcgen05")
    torch.cuda.synchronize()
    got = y.float().permute(0, 4, 1, 2, 3).cpu()
    assert relerr(got, ref) < 5e-3
    assert (got - ref).abs().max() < 0.05


def test_conv3_tcgen05_dgrad_and_epilogue_addend():
    torch.manual_seed(7)
    n, cin, cout, shape = 2, 64, 128, (4, 16, 8)
    x = torch.randn(n, cin, *shape).bfloat16().float().requires_grad_()
    w = (torch.randn(cout, cin, 3, 3, 3) / (27 * cin) ** 0.5).bfloat16().float().requires_grad_()
    add = torch.randn(n, cout, *shape).bfloat16().float()
    ref = F.elu(F.conv3d(x, w, None, padding=1) + add)
    g = torch.randn_like(ref).bfloat16().float()
    ref.backward(g)
    xg = x.detach().permute(0, 2, 3, 4, 1).contiguous().to(DEV, torch.bfloat16).requires_grad_()
    wg = w.detach().to(DEV).requires_grad_()
    ag = add.permute(0, 2, 3, 4, 1).contiguous().to(DEV, torch.bfloat16)
    y = ops.Conv3x3Fn.apply(xg, wg, None, ag, 3, "tcgen05")
    y.backward(g.permute(0, 2, 3, 4, 1).contiguous().to(DEV, torch.bfloat16))
    assert relerr(y.detach().float().permute(0, 4, 1, 2, 3).cpu(), ref.detach()) < 5e-3
    assert relerr(xg.grad.float().permute(0, 4, 1, 2, 3).cpu(), x.grad) < 1e-2
    cos = F.cosine_similarity(wg.grad.flatten().cpu(), w.grad.flatten(), dim=0).item()
    assert cos > 0.999


def test_auto_selects_tensor_cores_for_bf16_and_simt_for_fp32():
    x = torch.zeros(1, 4, 16, 8, 64, device=DEV, dtype=torch.bfloat16)
    assert ops.conv_select_impl(x.shape, (4, 16, 8), 64, 64, torch.bfloat16, 0, "auto", x.data_ptr()) == 2
    assert ops.conv_select_impl(x.shape, (4, 16, 8), 64, 64, torch.float32, 0, "auto", x.data_ptr()) == 1
    assert ops.conv_select_impl((1, 4, 16, 8, 1), (4, 16, 8), 1, 32, torch.bfloat16, 0, "auto") == 1   # Cin = 1


# ---------------------------------------------------------------------------------------------------------
# tensor-core weight gradient (csrc/wgrad_tcgen05.cu): MN-major operands, kw-chained taps, split-K partials
# ---------------------------------------------------------------------------------------------------------
WGRAD_CASES = [
    # N, Cin, Cout, (D,H,W)
    (1, 32, 64, (2, 16, 8)),        # U = dY (64 ch, second M atom zero-filled), one brick
    (1, 64, 64, (4, 16, 16)),
    (2, 64, 128, (5, 20, 12)),      # ragged bricks in all axes, odd depth; U = dY 128
    (1, 192, 64, (4, 16, 8)),       # U = X (192 = 128 + 64): mirrored taps, partially filled second U tile
    (1, 128, 128, (3, 9, 9)),
    (1, 384, 128, (2, 8, 8)),       # U = X, 3 U tiles, 4 S chunks
    (2, 256, 512, (1, 5, 7)),       # depth 1 (TD = 1), U = dY 512
    (1, 16, 32, (4, 16, 8)),        # cfg-5 first level: U = dY 32 (3/4 of M zero-filled), S = X 16 (half a chunk)
    (2, 32, 32, (3, 10, 9)),
    (1, 96, 32, (2, 16, 8)),        # U = X 96 = 64 + 32
    (1, 24, 40, (2, 8, 8)),         # multiples of 8 only
    (1, 64, 96, (2, 16, 8)),        # U = dY 96 = 64 + 32 (second atom half filled)
    (1, 20, 36, (2, 8, 8)),         # not multiples of 8: must be refused (CUDA-core kernel)
]


def _wgrad_ref(x, dy):
    """dW of y = conv3d(x, w, padding=1) for upstream gradient dy (fp64 on the CPU)."""
    cin, cout = x.shape[1], dy.shape[1]
    w = torch.zeros(cout, cin, 3, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv3d(x.double(), w, None, padding=1).backward(dy.double())
    return w.grad


@pytest.mark.parametrize("n,cin,cout,shape", WGRAD_CASES)
def test_wgrad_tcgen05_exact_on_integer_data(n, cin, cout, shape):
    """Small-integer operands: every product and every fp32 partial sum is exact, so the tensor-core result must
    equal the reference BIT FOR BIT whatever the summation order (catches any mis-addressed tap/voxel/channel)."""
    torch.manual_seed(cin * 7 + cout)
    x = torch.randint(-3, 4, (n, cin) + shape).float()
    dy = torch.randint(-3, 4, (n, cout) + shape).float()
    xg = x.permute(0, 2, 3, 4, 1).contiguous().to(DEV, torch.bfloat16)
    dg = dy.permute(0, 2, 3, 4, 1).contiguous().to(DEV, torch.bfloat16)
    p = ops.wgrad_params(dg, xg, 0, "auto")
    impl = ops.lib().mednet_conv3d_wgrad_select_impl(ops._abi.C.byref(p))
    cu, cs = (cin, cout) if cin > cout else (cout, cin)       # U = the operand with more channels, S = the other
    if cu % 8 != 0 or cs % 8 != 0 or cs <= 4:
        assert impl == 1                                      # refused by the tensor-core planner -> CUDA-core kernel
        return
    assert impl == 2, "bf16 wgrad with 8-aligned channels must run on the tensor cores"
    dw, _ = ops.k_wgrad(dg, xg, 0, "tcgen05")
    torch.cuda.synchronize()
    ref = _wgrad_ref(x, dy)
    assert torch.equal(dw.cpu().double(), ref)


def test_wgrad_tcgen05_random_data_and_bias():
    torch.manual_seed(3)
    n, cin, cout, shape = 2, 64, 128, (6, 18, 10)
    x = torch.randn(n, cin, *shape).bfloat16().float()
    dy = torch.randn(n, cout, *shape).bfloat16().float()
    xg = x.permute(0, 2, 3, 4, 1).contiguous().to(DEV, torch.bfloat16)
    dg = dy.permute(0, 2, 3, 4, 1).contiguous().to(DEV, torch.bfloat16)
    dw, db = ops.k_wgrad(dg, xg, 0, "tcgen05", want_bias=True)
    ref = _wgrad_ref(x, dy).float()
    assert relerr(dw.cpu(), ref) < 1e-5          # bf16 products are exact in fp32; only the summation order differs
    assert relerr(db.cpu(), dy.sum(dim=(0, 2, 3, 4))) < 1e-5
    dw_simt, _ = ops.k_wgrad(dg, xg, 0, "simt")
    assert relerr(dw.cpu(), dw_simt.cpu()) < 1e-5


def test_wgrad_tcgen05_is_deterministic():
    torch.manual_seed(5)
    xg = torch.randn(2, 8, 16, 16, 64, device=DEV).to(torch.bfloat16)
    dg = torch.randn(2, 8, 16, 16, 64, device=DEV).to(torch.bfloat16)
    a, _ = ops.k_wgrad(dg, xg, 0, "tcgen05")
    b, _ = ops.k_wgrad(dg, xg, 0, "tcgen05")
    assert torch.equal(a, b)


@pytest.mark.parametrize("cin,cout,shape", [(32, 16, (4, 16, 8)), (64, 32, (3, 9, 11)), (128, 64, (2, 5, 6)), (16, 16, (5, 4, 3)),
                                            (16, 32, (3, 4, 5))])
def test_conv_transpose_tcgen05_exact_on_integer_data(cin, cout, shape):
    """ConvTranspose3d(k3, s2, p1, op1) + bias + skip sum on the tensor cores (8 parity classes of 1/2/4/8 taps;
    dgrad through a stride-2 TMA map) -- small-integer operands make every product and fp32 sum exact, so forward and
    input gradient must equal F.conv_transpose3d bit for bit after the bf16 rounding of the OUTPUT.
    ref: midasmednet/unet/components.py:259-264, 283-284."""
    torch.manual_seed(cin + cout)
    x = torch.randint(-2, 3, (2, cin) + shape).float().requires_grad_()
    w = torch.randint(-1, 2, (cin, cout, 3, 3, 3)).float().requires_grad_()
    b = torch.randint(-2, 3, (cout,)).float()
    big = tuple(2 * s for s in shape)
    skip = torch.randint(-2, 3, (2, cout) + big).float()
    ref = F.conv_transpose3d(x, w, b, stride=2, padding=1, output_padding=1) + skip
    g = torch.randint(-1, 2, ref.shape).float()
    ref.backward(g)
    nd = lambda t: t.permute(0, 2, 3, 4, 1).contiguous().to(DEV, torch.bfloat16)
    xg = nd(x.detach()).requires_grad_()
    wg, bg = w.detach().to(DEV).requires_grad_(), b.to(DEV).requires_grad_()
    assert ops.conv_select_impl(xg.shape, big, cin, cout, xg.dtype, 1, "auto", xg.data_ptr()) == 2
    y = ops.ConvTranspose3x3Fn.apply(xg, wg, bg, nd(skip), "auto")
    y.backward(nd(g))
    back = lambda t: t.float().permute(0, 4, 1, 2, 3).cpu()
    assert torch.equal(back(y.detach()), ref.detach().bfloat16().float())
    assert torch.equal(back(xg.grad), x.grad.bfloat16().float())
    p = ops.wgrad_params(xg.detach(), nd(g), 2, "auto")
    assert ops.lib().mednet_conv3d_wgrad_select_impl(ops._abi.C.byref(p)) == 2      # weight gradient on the tensor cores too
    assert torch.equal(wg.grad.cpu(), w.grad)                      # integer data: exact whatever the summation order
    assert relerr(bg.grad.cpu(), g.sum(dim=(0, 2, 3, 4))) < 1e-6


def _upconv_ref(xl, w):
    return F.conv3d(F.interpolate(xl, scale_factor=2, mode="nearest"), w, None, padding=1)


@pytest.mark.parametrize("cl,cout,shape", [(32, 64, (3, 16, 8)), (128, 64, (2, 8, 8)), (64, 32, (5, 9, 7)),
                                           (128, 64, (4, 8, 8)), (64, 256, (5, 8, 8))])
def test_upsample_conv_tcgen05_exact_on_integer_data(cl, cout, shape):
    """MEDNET_GATHER_UPCONV_F / _B and the UPCONV_B weight gradient: conv3(nearest_up2(x)) evaluated on the coarse grid
    (8 summed taps per output parity class) against F.interpolate + F.conv3d on small-integer data -- every product and
    sum is exact in fp32, so the three kernels must agree with torch bit for bit (after the bf16 output rounding)."""
    torch.manual_seed(cl + cout)
    ops.calibrate_tcgen05()
    n = 2
    xl = torch.randint(-2, 3, (n, cl) + shape).float().requires_grad_()
    w = torch.randint(-1, 2, (cout, cl, 3, 3, 3)).float().requires_grad_()
    fine = tuple(2 * v for v in shape)
    ref = _upconv_ref(xl, w)
    g = torch.randint(-2, 3, ref.shape).float()
    ref.backward(g)
    xd = xl.detach().permute(0, 2, 3, 4, 1).contiguous().to(DEV, torch.bfloat16)
    gd = g.permute(0, 2, 3, 4, 1).contiguous().to(DEV, torch.bfloat16)
    wd = w.detach().to(DEV)
    y = ops.k_conv3(xd, ops.k_pack_weights(wd, cl, cout, torch.bfloat16, 6), cout, fine, 3, 2)
    assert torch.equal(y.float().cpu().permute(0, 4, 1, 2, 3), ref.detach().bfloat16().float())
    dx = ops.k_conv3(gd, ops.k_pack_weights(wd, cl, cout, torch.bfloat16, 7), cl, shape, 4, 2)
    assert torch.equal(dx.float().cpu().permute(0, 4, 1, 2, 3), xl.grad.bfloat16().float())
    dw, _ = ops.k_wgrad(xd, gd, 4, "tcgen05")                       # (Cl, Cout, 3, 3, 3)
    assert torch.equal(dw.cpu().permute(1, 0, 2, 3, 4), w.grad)


@pytest.mark.parametrize("cs,cl,cout,shape", [(64, 128, 64, (4, 8, 4)), (32, 64, 32, (3, 8, 8))])
def test_upsample_aware_decoder_join_against_torch(cs, cl, cout, shape):
    """GroupNorm(cat(skip, up(low))) -> conv -> ReLU through the split GroupNorm + conv3(skip) + coarse-grid conv(low)
    against the reference ops (components.py:277-280, :57, :8-9) with bf16 storage, forward and all five gradients."""
    torch.manual_seed(cs + cl)
    fine = tuple(2 * v for v in shape)
    q = lambda t: t.bfloat16().float()
    skip = q(torch.randn(2, cs, *fine)).requires_grad_()
    low = q(torch.randn(2, cl, *shape) * 1.5 + 0.3).requires_grad_()
    gamma = (torch.rand(cs + cl) + 0.5).requires_grad_()
    beta = (torch.randn(cs + cl) * 0.3).requires_grad_()
    w = (torch.randn(cout, cs + cl, 3, 3, 3) / (27 * (cs + cl)) ** 0.5).requires_grad_()
    cat = torch.cat((skip, F.interpolate(low, size=fine, mode="nearest")), dim=1)
    xn = F.group_norm(cat, 8, gamma, beta, eps=1e-5)
    xn = xn + (q(xn) - xn).detach()
    wq = w + (q(w) - w).detach()
    ref = F.relu(F.conv3d(xn, wq, None, padding=1))
    g = q(torch.randn_like(ref))
    ref.backward(g)
    nd = lambda t: t.detach().permute(0, 2, 3, 4, 1).contiguous().to(DEV, torch.bfloat16)
    sd, ld = nd(skip).requires_grad_(), nd(low).requires_grad_()
    gmd, btd, wd = (t.detach().to(DEV).requires_grad_() for t in (gamma, beta, w))
    assert ops.upconv_supported(sd, ld, wd)
    sn, ln = ops.UpcatGroupNormSplitFn.apply(sd, ld, gmd, btd, 8, 0, 0)
    y = ops.UpConvJoinFn.apply(sn, ln, wd, 1)
    y.backward(nd(g))
    back = lambda t: t.detach().float().cpu().permute(0, 4, 1, 2, 3)
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
    # small volumes in bf16: same tolerances as the materialised-concat path (tests/test_ops_gpu.py); the tight per-layer
    # gate (5e-3) is taken at 128^3 by tests/test_parity_gpu.py through the same code
    assert rel(back(y), ref.detach()) < 6e-3
    assert rel(back(sd.grad), skip.grad) < 4e-2
    assert rel(back(ld.grad), low.grad) < 4e-2
    assert rel(wd.grad.cpu(), w.grad) < 6e-2          # ReLU' flips where |pre-activation| is below the bf16 noise: ~sqrt(2e-3)
    assert rel(gmd.grad.cpu(), gamma.grad) < 5e-2
    assert rel(btd.grad.cpu(), beta.grad) < 5e-2
    # and equal (to bf16 rounding) to the materialised-concat path of the same library
    s2, l2 = nd(skip).requires_grad_(), nd(low).requires_grad_()
    g2, b2, w2 = (t.detach().to(DEV).requires_grad_() for t in (gamma, beta, w))
    y2 = ops.Conv3x3Fn.apply(ops.UpcatGroupNormFn.apply(s2, l2, g2, b2, 8, 0, 0), w2, None, None, 1, "auto")
    y2.backward(nd(g))
    assert rel(back(y), back(y2)) < 6e-3
    assert rel(back(sd.grad), back(s2.grad)) < 6e-2
    assert rel(back(ld.grad), back(l2.grad)) < 6e-2
    assert rel(wd.grad.cpu(), w2.grad.cpu()) < 6e-2  # the two paths round the effective weights differently: other flips

"""Parity at the FULL sizes of BASELINE.json (128^3 patches, f=64 channel widths), where the CPU oracle is too slow:

* the tensor-core convolutions (fprop, dgrad, wgrad) against the reference's own arithmetic engine -- torch's CUDA
  F.conv3d in fp32 on the same bf16-rounded operands, run here on the GPU purely as the checker;
* the whole UNet3D f=64 forward on one 128^3 patch against the oracle restatement evaluated ON the GPU with bf16
  storage (oracle.unet.Storage.bf16), north-star gate: logits relative error <= 1e-2;
* size-independent properties: impulse response at volume corners/faces across tile and CTA-wave boundaries,
  additivity of the weight gradient over a batch split, identity round trip of the 512x512x400 / 128^3 / overlap-16
  tile geometry of cfg-4.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from mednet_b200 import ops
from mednet_b200.predict import SlidingWindowPredictor
from mednet_b200.unet.model import UNet3D
from oracle import steps as osteps
from oracle import unet as ounet

pytestmark = pytest.mark.gpu
DEV = "cuda"


def relerr(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def nd(t):                                   # (N,C,D,H,W) -> NDHWC bf16 on the device
    return t.permute(0, 2, 3, 4, 1).contiguous().to(DEV, torch.bfloat16)


@pytest.mark.parametrize("cin,cout,n,edge", [(64, 64, 2, 128), (192, 64, 1, 128), (384, 128, 2, 64)])
def test_conv_fprop_dgrad_wgrad_full_size_against_torch_cuda(cin, cout, n, edge):
    torch.manual_seed(cin + cout)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    x = torch.randn(n, cin, edge, edge, edge, device=DEV).bfloat16().float().requires_grad_()
    w = (torch.randn(cout, cin, 3, 3, 3, device=DEV) / (27 * cin) ** 0.5).bfloat16().float().requires_grad_()
    ref = F.relu(F.conv3d(x, w, None, padding=1))
    g = torch.randn_like(ref).bfloat16().float()
    ref.backward(g)
    xg = x.detach().permute(0, 2, 3, 4, 1).contiguous().bfloat16().requires_grad_()
    wg = w.detach().clone().requires_grad_()
    assert ops.conv_select_impl(xg.shape, (edge,) * 3, cin, cout, xg.dtype, 0, "auto", xg.data_ptr()) == 2
    y = ops.Conv3x3Fn.apply(xg, wg, None, None, 1, "auto")
    y.backward(g.permute(0, 2, 3, 4, 1).contiguous().bfloat16())
    torch.cuda.synchronize()
    assert relerr(y.detach().float().permute(0, 4, 1, 2, 3), ref.detach()) < 5e-3
    assert relerr(xg.grad.float().permute(0, 4, 1, 2, 3), x.grad) < 5e-3
    # fp32 output; beyond the summation order, ReLU'(pre) differs from cuDNN's wherever |pre| is below the fp32 accumulation
    # noise (a ~1e-6 fraction of the voxels carrying O(1) gradients): ~1e-3 relative on a 27*Cin*Cout sum over 2M voxels
    assert relerr(wg.grad, w.grad) < 5e-3


def test_impulse_response_across_tiles_and_waves():
    """conv(delta at p) = the flipped kernel around p: exact (one product per output), for impulses at corners, faces
    and interior points of a 128^3 volume -> every tile / halo / zero-padding path of the persistent kernel."""
    torch.manual_seed(0)
    c, edge = 64, 128
    w = torch.randint(-32, 33, (c, c, 3, 3, 3)).float() / 64.0        # multiples of 1/64: overlapping responses sum exactly
    pts = [(0, 0, 0), (127, 127, 127), (0, 127, 64), (5, 16, 8), (63, 15, 7), (64, 16, 8), (126, 1, 120), (3, 127, 0)]
    x = torch.zeros(1, c, edge, edge, edge)
    for i, p in enumerate(pts):
        x[0, (7 * i) % c, p[0], p[1], p[2]] = 1.0 + i
    y = ops.Conv3x3Fn.apply(nd(x), w.to(DEV), None, None, 0, "tcgen05").float().cpu()      # (1, D, H, W, C)
    want = torch.zeros(edge, edge, edge, c)
    for i, p in enumerate(pts):
        ci = (7 * i) % c
        for kd in range(3):
            for kh in range(3):
                for kw in range(3):
                    q = (p[0] - kd + 1, p[1] - kh + 1, p[2] - kw + 1)
                    if all(0 <= v < edge for v in q):
                        want[q[0], q[1], q[2]] += (1.0 + i) * w[:, ci, kd, kh, kw]
    nz = want.abs().sum(-1) > 0
    assert torch.equal(y[0][nz], want[nz].bfloat16().float())
    assert float(y[0][~nz].abs().max()) == 0.0


def test_wgrad_is_additive_over_a_batch_split():
    """dW(batch A + batch B) == dW(A) + dW(B) (fp32, different split-K decompositions) at 64 -> 64 channels, 128^3."""
    torch.manual_seed(1)
    x = torch.randn(2, 128, 128, 128, 64, device=DEV).bfloat16()
    dy = torch.randn(2, 128, 128, 128, 64, device=DEV).bfloat16()
    full, _ = ops.k_wgrad(dy, x, 0, "tcgen05")
    a, _ = ops.k_wgrad(dy[:1].contiguous(), x[:1].contiguous(), 0, "tcgen05")
    b, _ = ops.k_wgrad(dy[1:].contiguous(), x[1:].contiguous(), 0, "tcgen05")
    e = relerr(full, a + b)
    print(f"wgrad additivity over a batch split, 4.2 M voxels per sample: relative difference {e:.2e}")
    assert e < 2e-4                                       # fp32 sums of 2 x 2.1 M products in different orders


def test_unet3d_f64_forward_one_128_patch_against_the_oracle_on_gpu():
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = UNet3D(1, 4, False).to(DEV)                      # f_maps = 64, 4 levels: the cfg-3 network
    x = torch.randn(1, 1, 128, 128, 128, device=DEV)
    with torch.no_grad():
        logits = net(x)
        sd = {k: v.detach().float() for k, v in net.state_dict().items()}
        ref16 = ounet.unet3d_forward(sd, x, f_maps=64, storage=ounet.Storage.bf16())
        ref32 = ounet.unet3d_forward(sd, x, f_maps=64)
    e_kernel, e_format = relerr(logits, ref16), relerr(ref16, ref32)
    print(f"128^3 f=64: vs bf16-storage oracle {e_kernel:.2e}; bf16-storage oracle vs fp32 oracle {e_format:.2e}")
    assert logits.shape == (1, 4, 128, 128, 128) and logits.dtype == torch.float32
    # 14 conv layers of 64..512 channels amplify every difference in fp32 summation order through ReLU'/pool decisions
    # (oracle/gates.py): the product must be closer to the bf16-storage reference than that reference is to fp32, and no
    # further from fp32 than twice the format's own distance
    assert e_kernel < max(1e-2, 0.75 * e_format)
    assert relerr(logits, ref32) < 2.0 * e_format + 5e-3
    flips = float((logits.argmax(1) != ref16.argmax(1)).float().mean())
    print(f"label-map disagreement with the bf16-storage oracle (near-ties of a random-init head): {flips:.4f}")
    assert flips < 0.25


def test_cfg4_tile_geometry_round_trip_identity():
    """512 x 512 x 400 volume, 128^3 tiles, overlap 16 (180 tiles): gather -> identity 'network' -> uint8 epilogue ->
    centre-crop scatter reproduces the volume (heatmap channel: clip/truncate to uint8 of integer data is lossless)."""
    class Identity(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))
            from mednet_b200.unet.components import ComputeConfig
            self.cfg = ComputeConfig(torch.float32)

        def forward(self, t):                                # (B,1,P,P,P) -> logits (B, 1 heatmap + 2 classes, ...)
            t = t.float()
            return torch.cat([t, t, -t], dim=1)

    rng = np.random.default_rng(0)
    vol = rng.integers(0, 256, size=(1, 512, 512, 400)).astype(np.float32)
    pred = SlidingWindowPredictor(Identity().to(DEV), [128] * 3, [16] * 3, 1, batch_size=6)
    out = pred(vol).cpu().numpy()
    assert pred.tiles_done == 180
    assert np.array_equal(out[0], vol[0].astype(np.uint8))              # heatmap channel = the volume itself
    assert np.array_equal(out[1], (vol[0] < 0).astype(np.uint8))        # argmax(softmax([t, -t])) = 0 for t >= 0

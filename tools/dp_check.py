#!/usr/bin/env python
"""Data-parallel equivalence on real GPUs (run under torchrun, >= 2 ranks):
the production path -- asynchronous weight gradients on the side stream + bucketed all-reduce launched behind it,
overlapped with backward -- must give the same summed gradient as the plain path (synchronous gradients, one
all-reduce of the whole flat buffer after backward).  Prints one JSON line on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-mednet_b200"))
from mednet_b200.optim import FusedAdam  # noqa: E402
from mednet_b200.parallel import BucketedAllReduce, init_distributed  # noqa: E402
from mednet_b200.unet.loss import DiceLoss  # noqa: E402
from mednet_b200.unet.model import UNet3D  # noqa: E402


def main():
    rank, local, world = init_distributed()
    dev = torch.device("cuda", local)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(2, 1, 32, 32, 32, generator=g).to(dev)
    y = torch.randint(0, 3, (2, 32, 32, 32), generator=g).to(dev)
    flats = []
    names = None
    for overlapped, use_reducer in ((True, True), (False, False), (True, False), (False, True)):
        torch.manual_seed(0)                                   # identical replicas
        net = UNet3D(1, 3, False, f_maps=[16, 32, 64]).to(dev)
        opt = FusedAdam(net.parameters(), lr=1e-3, async_wgrad=overlapped)
        reducer = BucketedAllReduce(opt.grad_slices(), opt.flat_grad, bucket_bytes=256 << 10) if use_reducer else None
        names = [(k, p.numel()) for k, p in net.named_parameters()]
        opt.zero_grad()
        for _ in range(2):                                     # two steps: the bucket counters must re-arm
            opt.zero_grad()
            DiceLoss()(net(x), y).backward()
            if reducer is not None:
                assert all(b[2] == 0 for b in reducer.buckets), [b[2] for b in reducer.buckets]   # every parameter reported once
                scale = reducer.finish()
                assert abs(scale - 1.0 / world) < 1e-12
            else:
                opt.sync_gradients()
                dist.all_reduce(opt.flat_grad, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
        flats.append(opt.flat_grad.clone())
        if reducer is not None:
            nb = len(reducer.buckets)
            reducer.remove()
    a, b = flats[0], flats[1]
    rel = ((a - b).norm() / b.norm()).item()
    if rank == 0:
        for tag, f in (("async+reducer", flats[0]), ("async only", flats[2]), ("reducer only", flats[3])):
            print(tag, "rel diff vs plain", ((f - b).norm() / b.norm()).item(), flush=True)
            off, shown = 0, 0
            for k, n in names:
                d = ((f[off:off + n] - b[off:off + n]).norm() / (b[off:off + n].norm() + 1e-30)).item()
                if d > 1e-5 and shown < 6:
                    print("    ", k, n, d, "norms", f[off:off + n].norm().item(), b[off:off + n].norm().item(), flush=True)
                    shown += 1
                off += n
    ok = rel < 1e-6 and float(b.abs().sum()) > 0
    # tile-sharded sliding-window inference (dataset.py:349-389, 444-474): every rank ends with the volume a single rank
    # produces, bit for bit (tiles are independent; only each rank's own centre crops are exchanged)
    import numpy as np
    from mednet_b200.predict import SlidingWindowPredictor
    torch.manual_seed(3)
    pnet = UNet3D(1, 2 + 3, False, f_maps=[16, 32]).to(dev)
    with torch.no_grad():
        pnet.final_conv.weight.mul_(40.0)
    pnet.eval()
    vol = np.random.default_rng(0).standard_normal((1, 70, 45, 52)).astype(np.float32)
    single = SlidingWindowPredictor(pnet, [32] * 3, [4] * 3, 2, batch_size=1)(vol)
    sharded = SlidingWindowPredictor(pnet, [32] * 3, [4] * 3, 2, batch_size=1, rank=rank, world=world)(vol)
    predict_ok = bool(torch.equal(single, sharded))
    ok = ok and predict_ok
    out = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(out, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"dp_check": "ok" if out.item() == 1.0 else "MISMATCH", "world": world, "buckets": nb,
                          "rel_diff_overlapped_vs_plain": rel, "sharded_predict_equals_single_rank": predict_ok}))
    dist.destroy_process_group()
    sys.exit(0 if out.item() == 1.0 else 1)


if __name__ == "__main__":
    main()

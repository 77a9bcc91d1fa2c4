"""world_size-2 gloo test of the bucketed gradient all-reduce (host-side logic of the data-parallel path)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "torch-mednet_b200"))
    from mednet_b200.parallel import BucketedAllReduce, FlatGradients
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                                torch.nn.Linear(16, 4))
    flat = FlatGradients(model.parameters())
    red = BucketedAllReduce(flat.slices, flat.flat, bucket_bytes=600)      # several small buckets
    torch.manual_seed(100 + rank)
    x = torch.randn(5, 8)
    for _ in range(2):                                                    # second round checks the re-armed counters
        flat.zero()
        model(x).square().sum().backward()
        order = list(red.launch_order)
        scale = red.finish()
    res = {"flat": flat.flat.clone() * scale, "order": order, "nb": len(red.buckets),
           "ranges": [(b[0], b[1]) for b in red.buckets]}
    # single-process reference: mean over both ranks' gradients
    ref = torch.zeros_like(flat.flat)
    for r in range(world):
        torch.manual_seed(0)
        m2 = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                                 torch.nn.Linear(16, 4))
        torch.manual_seed(100 + r)
        m2(torch.randn(5, 8)).square().sum().backward()
        ref += torch.cat([p.grad.flatten() for p in m2.parameters()])
    res["ref"] = ref / world
    if rank == 0:
        torch.save(res, out)
    dist.destroy_process_group()


def test_bucketed_allreduce_two_ranks(tmp_path):
    out = str(tmp_path / "res.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    torch.testing.assert_close(res["flat"], res["ref"], rtol=1e-5, atol=1e-6)
    assert res["nb"] >= 3
    assert res["order"] == list(range(res["nb"]))            # buckets fire in backward order: last layers first
    ranges = res["ranges"]
    assert ranges[0][1] == res["flat"].numel() and ranges[-1][0] == 0
    assert all(ranges[i][0] == ranges[i + 1][1] for i in range(len(ranges) - 1))   # contiguous cover, end to start


def test_async_reported_gradients_are_counted_once():
    """A conv weight whose gradient was enqueued on the side stream reports through the listener (`_on_async`);
    autograd still fires its post-accumulate hook with an undefined gradient (torch >= 2.1 behaviour, measured on two
    GPUs as every bucket being all-reduced twice).  The hook must not count it again, and must still count parameters
    that took the synchronous path."""
    import torch
    from mednet_b200.parallel import BucketedAllReduce, FlatGradients

    params = [torch.nn.Parameter(torch.randn(4, 3)) for _ in range(3)]
    fg = FlatGradients(params)
    red = BucketedAllReduce(fg.slices, fg.flat, bucket_bytes=1 << 20)
    assert len(red.buckets) == 1 and red.buckets[0][2] == 3
    red._on_async(params[2])             # async conv weight: listener first ...
    red._on_hook(params[2])              # ... then autograd's hook with an undefined gradient: ignored
    red._on_hook(params[1])
    assert red.buckets[0][2] == 1 and red.launch_order == []
    red._on_hook(params[0])
    assert red.buckets[0][2] == 0 and red.launch_order == [0]
    assert red.finish() == 1.0
    assert red.buckets[0][2] == 3 and not red._async_reported        # re-armed
    # the hook really fires for an undefined gradient
    class F(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w):
            return x * 2 + 0 * w.sum()

        @staticmethod
        def backward(ctx, g):
            return g * 2, None
    x = torch.ones(4, 3, requires_grad=True)
    F.apply(x, params[0]).sum().backward()
    assert red.buckets[0][2] == 2
    red.remove()


def _loader_worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "torch-mednet_b200"))
    from mednet_b200.dataset import SyntheticSegmentationDataset
    from mednet_b200.parallel import sharded_loader
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ds = SyntheticSegmentationDataset(11, (4, 4, 4))
    seen = []
    for epoch in range(2):
        batches = [b["data"] for b in sharded_loader(ds, 2, shuffle=True, epoch=epoch)]      # rank/world from the group
        seen.append(torch.cat(batches))
    torch.save(seen, f"{out}.{rank}")
    dist.destroy_process_group()


def test_each_rank_gets_its_own_share_of_an_epoch(tmp_path):
    """Patch batches are the unit of data parallelism: with the same seed on every rank the hooks must still hand
    different samples to different ranks, the same number of them, and reshuffle per epoch."""
    from mednet_b200.dataset import SyntheticSegmentationDataset
    from mednet_b200.parallel import epoch_order
    out = str(tmp_path / "seen")
    mp.spawn(_loader_worker, args=(2, 29500 + (os.getpid() % 2000) + 1, out), nprocs=2, join=True)
    a, b = torch.load(out + ".0"), torch.load(out + ".1")
    ds = SyntheticSegmentationDataset(11, (4, 4, 4))
    every = torch.stack([ds[i]["data"] for i in range(11)])

    def ids(t):
        return [int((every == p).flatten(1).all(1).nonzero()) for p in t]
    for epoch in range(2):
        ia, ib = ids(a[epoch]), ids(b[epoch])
        assert len(ia) == len(ib) == 5 and not set(ia) & set(ib)              # 11 -> 10 samples, 5 each, disjoint
        assert ia == epoch_order(11, True, epoch, 0, 0, 2).tolist()
    assert ids(a[0]) != ids(a[1])
    assert epoch_order(7, False, 0).tolist() == list(range(7))
    assert sorted(epoch_order(9, True, 3, rank=0, world=1).tolist()) == list(range(9))

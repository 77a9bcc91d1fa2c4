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

"""World-size-2 gloo test (CPU) of the multi-rank plumbing bench.py uses: the batch
shard plan (rank r owns [r*B/G, (r+1)*B/G), no data-path collective),
the max-over-ranks job time and the all-samples-processed-once check."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import bench
    w, r, _ = bench.dist_setup("gloo")
    assert (w, r) == (world, rank)
    mine = bench.shard_samples(batch, w, r)
    # each rank "processes" its shard: a per-sample checksum stands in for the kernels
    local = sum(float(s + 1) for s in mine)
    total = bench.sum_over_ranks(local, w)
    count = bench.sum_over_ranks(len(mine), w)
    slowest = bench.max_over_ranks(10.0 + rank, w)          # rank 1 is the slow one
    gathered = [None] * w
    dist.all_gather_object(gathered, mine)
    if r == 0:
        out.put((total, count, slowest, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_plan_and_timing_reduction():
    world, batch = 2, 8
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    total, count, slowest, gathered = q.get()
    assert count == batch
    assert total == sum(range(1, batch + 1))                # every sample exactly once
    assert slowest == 11.0                                  # max over ranks, not rank 0's clock
    assert gathered == [[0, 1, 2, 3], [4, 5, 6, 7]]


def test_shard_plan_uneven():
    import bench
    for world in (1, 2, 3, 4, 8):
        seen = sorted(s for r in range(world) for s in bench.shard_samples(8, world, r))
        assert seen == list(range(8))
        sizes = [len(bench.shard_samples(8, world, r)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def _ddp_worker(rank, world, port, out):
    """DDP over gloo on CPU with the libs/sync_bn shim classes in the model and a layer the forward never
    calls (what the reference models contain): gradients must come out averaged over the ranks."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.nn as nn
    from libs.sync_bn.modules.sync_bn import BatchNorm2d
    dist.init_process_group("gloo")

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = nn.Conv2d(2, 3, 3, padding=1)
            self.bn = BatchNorm2d(3)
            self.unused = nn.Conv2d(3, 3, 1)          # like GANet_deep.py:305 deconv0b

        def forward(self, x):
            return self.bn(self.conv(x)).mean()

    torch.manual_seed(0)
    net = Net()
    net.bn.eval()                                      # frozen statistics: nothing couples the samples
    ddp = torch.nn.parallel.DistributedDataParallel(net, find_unused_parameters=True)
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(world, 2, 5, 5, generator=gen)
    ddp(x[rank:rank + 1]).backward()
    g = net.conv.weight.grad.clone()
    if rank == 0:
        torch.manual_seed(0)
        ref = Net()
        ref.bn.eval()
        sum(ref(x[i:i + 1]) for i in range(world)).div(world).backward()
        out.put((float((g - ref.conv.weight.grad).abs().max()), net.unused.weight.grad is None
                 or float(net.unused.weight.grad.abs().max()) == 0.0))
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_gradient_averaging_with_shim_batchnorm_and_unused_layer():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(180) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    err, unused_ok = q.get()
    assert err < 1e-6 and unused_ok

"""Host-side logic of the one parallel strategy on this path: rays sharded across ranks, one flat
gradient bucket all-reduced (AVG) per step.  world_size 2 over gloo on CPU; the NCCL path differs only in the backend."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from articulated_point_nerf_b200.train import GradBucket, shard_rays


def test_shard_rays_partitions_exactly():
    for n, w in [(8192, 8), (160000, 3), (7, 8), (0, 2), (1048576, 8)]:
        spans = [shard_rays(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    extra = torch.nn.Parameter(torch.randn(11))
    opt = torch.optim.SGD([{"params": list(lin.parameters())}, {"params": [extra]}], lr=0.1)
    bucket = GradBucket(opt)
    assert bucket.numel == sum(p.numel() for p in lin.parameters()) + 11
    # the full batch is 8 rows; each rank differentiates the mean over its contiguous shard
    x = torch.arange(40, dtype=torch.float32).reshape(8, 5) / 10
    lo, hi = shard_rays(8, rank, world)
    bucket.zero()
    loss = (lin(x[lo:hi]).pow(2).mean(dim=1) + (extra * x[lo:hi, :1]).sum(dim=1)).mean()
    loss.backward()
    for p in bucket.params:                     # autograd accumulated in place into the bucket views
        assert p.grad.data_ptr() >= bucket.flat.data_ptr()
    bucket.all_reduce_avg()
    q.put((rank, bucket.flat.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_reproduces_single_process_gradient():
    world, port = 2, 29500 + (os.getpid() % 500)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == got[1]
    # single-process reference: equal shards + AVG == gradient of the mean over the whole batch
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    extra = torch.nn.Parameter(torch.randn(11))
    x = torch.arange(40, dtype=torch.float32).reshape(8, 5) / 10
    (lin(x).pow(2).mean(dim=1) + (extra * x[:, :1]).sum(dim=1)).mean().backward()
    ref = torch.cat([p.grad.reshape(-1) for p in list(lin.parameters()) + [extra]])
    flat = torch.tensor(got[0])
    # strip the 64-element alignment padding between slices
    vals, o = [], 0
    for p in list(lin.parameters()) + [extra]:
        vals.append(flat[o:o + p.numel()])
        o += (p.numel() + 63) // 64 * 64
    assert torch.allclose(torch.cat(vals), ref, rtol=1e-5, atol=1e-6)

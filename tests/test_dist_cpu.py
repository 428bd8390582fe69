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


def _worker_parts(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from articulated_point_nerf_b200.scene import build_model, make_scene
    from articulated_point_nerf_b200.train import create_optimizer, make_bucket
    model = build_model(make_scene("tiny"))
    opt = create_optimizer(model)
    bucket = make_bucket(model, opt)                              # the pipelined layout
    gen = torch.Generator().manual_seed(100 + rank)
    bucket.flat[:bucket.total].copy_(torch.randn(bucket.total, generator=gen))
    bucket.status[:1].fill_(float(rank))                          # a flag raised on rank 1 only
    before = bucket.flat.clone()
    bucket.all_reduce_avg(part=1)                                 # warp slice + status: the exchange inside the step
    mid = bucket.flat.clone()
    bucket.all_reduce_avg(part=0)                                 # decoder slice: the exchange beside the next step
    q.put((rank, bucket.split, bucket.total, before.tolist(), mid.tolist(), bucket.flat.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_pipelined_bucket_layout_and_slice_exchange():
    """make_bucket(model, opt) = "pipeline": [canonical_feat | other non-warp parameters | warp parameters | status]; the two
    slices are exchanged separately (part=1 inside the step, part=0 beside the next one) and together cover the bucket; the
    status words travel with the warp slice, so a flag raised on one rank is non-zero on every rank after it."""
    from articulated_point_nerf_b200.scene import build_model, make_scene
    from articulated_point_nerf_b200.train import GradBucket, create_optimizer, decoder_parameters, make_bucket, warp_parameters
    model = build_model(make_scene("tiny"))
    opt = create_optimizer(model)
    b = make_bucket(model, opt)
    warp = {id(p) for p in warp_parameters(model)}
    assert b.params[0] is model.canonical_feat and b.offsets[0] == 0
    assert 0 < b.first_split < b.split < b.total
    for p, o in zip(b.params, b.offsets):
        assert (id(p) in warp) == (o >= b.split)
        assert o % 64 == 0 and p.grad.data_ptr() == b.flat.data_ptr() + 4 * o
    assert {id(p) for p in decoder_parameters(model)} <= {id(p) for p, o in zip(b.params, b.offsets) if o < b.split}
    assert b.split > 0.5 * b.total                                  # (82 % at c2 / c4 sizes: canonical_feat grows with the cloud)
    assert sum(p.numel() for p in b.params) == b.numel == sum(p.numel() for g in opt.param_groups for p in g["params"])
    # the other layouts
    b3 = make_bucket(model, opt, True)
    assert b3.params[0] is model.canonical_feat and 0 < b3.first_split < b3.split
    assert {id(p) for p, o in zip(b3.params, b3.offsets) if o < b3.split} == {id(p) for p in decoder_parameters(model)}
    b1 = make_bucket(model, opt, False)
    assert b1.split == 0 and isinstance(b1, GradBucket)
    assert make_bucket(model, opt, "auto").split == 0               # tiny scene: far below 32 MiB of gradients
    # two ranks
    world, port = 2, 29100 + (os.getpid() % 400)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_parts, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        r, split, total, before, mid, after = q.get(timeout=300)
        got[r] = (split, total, torch.tensor(before), torch.tensor(mid), torch.tensor(after))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    split, total = got[0][0], got[0][1]
    assert (split, total) == (b.split, b.total)
    avg = 0.5 * (got[0][2] + got[1][2])
    for r in range(world):
        _, _, before, mid, after = got[r]
        assert torch.equal(mid[:split], before[:split])             # part=1 leaves the decoder slice alone
        assert torch.allclose(mid[split:], avg[split:], rtol=0, atol=1e-6)
        assert torch.equal(after[split:], mid[split:])              # part=0 leaves the warp slice alone
        assert torch.allclose(after[:split], avg[:split], rtol=0, atol=1e-6)
        assert float(after[total]) == 0.5                           # the flag of rank 1, averaged: non-zero everywhere


def test_tile_shares_of_a_frame_partition_its_pixels():
    """render.tile_pixels (16x16 tiles dealt round-robin): the shares of all ranks are disjoint, sorted,
    cover the frame — also when the frame is not a multiple of the tile — and are balanced to within one tile row."""
    from articulated_point_nerf_b200.render import tile_pixels
    for H, W, world, tile in [(400, 400, 8, 16), (37, 53, 3, 16), (16, 16, 4, 16), (2048, 2048, 8, 16), (5, 7, 2, 4)]:
        shares = [tile_pixels(H, W, r, world, tile) for r in range(world)]
        allpix = torch.cat(shares).long()
        assert len(allpix) == H * W and torch.equal(allpix.sort()[0], torch.arange(H * W))
        for s in shares:
            assert s.dtype == torch.int32 and bool((s[1:] > s[:-1]).all())
        n_tiles = ((H + tile - 1) // tile) * ((W + tile - 1) // tile)
        if n_tiles >= world:
            sizes = [len(s) for s in shares]
            assert max(sizes) - min(sizes) <= tile * tile * (1 + (n_tiles % world != 0)) + tile * max(H, W)
        # a pixel's owner is its tile's raster index modulo the world size
        y, x = 3 % H, 5 % W
        owner = ((y // tile) * ((W + tile - 1) // tile) + x // tile) % world
        assert (y * W + x) in set(shares[owner].tolist())

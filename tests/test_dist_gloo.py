"""CPU, world_size 2 over gloo: the data-parallel host logic of the ray-sharded step.
Each rank takes a contiguous shard of the ray batch, evaluates the product's loss heads
(morpheus_b200.render.get_sdf_loss with the global normaliser, mean-type image losses) on its shard,
scales by 1/world_size and all-reduces the flat gradient: the result must equal the single-process gradient."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    g = torch.Generator().manual_seed(0)
    N, S = 8, 6
    theta = torch.randn(5, generator=g)
    ray_depth = torch.tensor([1.0, 0.0, 1.2, 0.9, 0.0, 0.0, 1.1, 1.0])       # uneven number of rays with depth per shard
    z = torch.rand(N * S, 1, generator=g) * 2
    feats = torch.randn(N * S, 5, generator=g)
    rgb_gt = torch.rand(N, 3, generator=g)
    return N, S, theta, ray_depth, z, feats, rgb_gt


def _shard_loss(theta, lo, hi, world, N, S, ray_depth, z, feats, rgb_gt, shipped=False):
    from morpheus_b200.render import get_sdf_loss, global_count
    ri = torch.arange(lo, hi).repeat_interleave(S)
    sl = slice(lo * S, hi * S)
    sdf = feats[sl] @ theta
    t_gt = ray_depth[ri][:, None]
    if shipped and world > 1:
        # bench.py / train.GraphedStep: whoever shards the batch ships the global count with the shard ('n_depth'), no collective
        cnt = torch.count_nonzero(ray_depth).float() * S / world
    else:
        cnt = global_count(torch.count_nonzero(t_gt), world) if world > 1 else None
    _, sdf_loss = get_sdf_loss(z[sl], t_gt, sdf, 0.5, mask=torch.ones_like(t_gt), rays_w_depth=cnt)
    img = torch.sigmoid(sdf.view(hi - lo, S).mean(1, keepdim=True) * torch.ones(1, 3))
    rgb_loss = torch.nn.functional.mse_loss(img, rgb_gt[lo:hi])
    return 5.0 * rgb_loss + 10.0 * sdf_loss


def _worker(rank, world, port, q, shipped=False):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    N, S, theta, *rest = _problem()
    th = theta.clone().requires_grad_(True)
    n_local = N // world
    loss = _shard_loss(th, rank * n_local, (rank + 1) * n_local, world, N, S, *rest, shipped=shipped)
    (loss / world).backward()
    flat = th.grad.clone()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put(flat)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize('shipped', [False, True])
def test_sharded_gradient_equals_single_process(shipped):
    N, S, theta, *rest = _problem()
    th = theta.clone().requires_grad_(True)
    _shard_loss(th, 0, N, 1, N, S, *rest).backward()
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, shipped)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    torch.testing.assert_close(got, th.grad, rtol=1e-5, atol=1e-6)


# ---- round 2: the other sharding rules of the step (masked means with global counts, sum-type terms, all-gathered image) ----
def _problem2():
    g = torch.Generator().manual_seed(1)
    N = 8
    theta = torch.randn(4, generator=g)
    feats = torch.randn(N, 4, generator=g)
    valid = torch.tensor([1.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, 1.0])           # uneven number of valid points per shard
    target = torch.rand(N, 3, generator=g)
    return N, theta, feats, valid, target


def _loss2(theta, lo, hi, world, rank, N, feats, valid, target):
    """a shard's loss: masked mean over valid points (denominator = GLOBAL count), a SUM-type term (loss_orient, morpheus.py:712), and a
    replicated loss on the all-gathered per-ray image (the SDS chain of a ray-sharded novel view)"""
    from morpheus_b200.render import global_count
    from morpheus_b200.train import _AllGatherRows, weighted_sum
    y = feats[lo:hi] @ theta                                                   # per-ray quantity of this shard
    n_valid = global_count(valid[lo:hi].sum(), world) if world > 1 else valid[lo:hi].sum()
    masked_mean = (y.square() * valid[lo:hi]).sum() / n_valid.clamp(min=1.0)
    sum_term = torch.relu(y).square().sum() * world                            # world x local sum: (loss / world) summed over ranks = global sum
    img = torch.sigmoid(y)[:, None] * torch.ones(1, 3)
    if world > 1:
        img = _AllGatherRows.apply(img, world, rank)
    replicated = (img - target).square().sum()                                 # identical on every rank; its gradient reaches each shard unscaled
    return weighted_sum([masked_mean, sum_term], [10.0 / world, 0.01 / world]) + replicated


def _worker2(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    N, theta, *rest = _problem2()
    th = theta.clone().requires_grad_(True)
    n_local = N // world
    _loss2(th, rank * n_local, (rank + 1) * n_local, world, rank, N, *rest).backward()
    flat = th.grad.clone()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put(flat)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_round2_sharding_rules_equal_single_process():
    N, theta, *rest = _problem2()
    th = theta.clone().requires_grad_(True)
    _loss2(th, 0, N, 1, 0, N, *rest).backward()
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker2, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    torch.testing.assert_close(got, th.grad, rtol=1e-5, atol=1e-6)

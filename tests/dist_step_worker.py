"""Worker of tests/test_step_parity_gpu.py::test_two_rank_sharded_gradient_equals_single_gpu (launched with torchrun, one rank per GPU).

Every rank computes the single-GPU gradient of the full ray batch, then the ray-sharded step (its contiguous half of the rays, the same
injected RNG draws) followed by the NCCL all-reduce of the flat gradient buffer; the two flat gradients must agree per parameter group.
Also replays the sharded step through train.GraphedStep (NCCL captured inside the CUDA graph) and checks the loss it reports.
"""
import datetime
import os
import sys
import traceback

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    from test_step_parity_gpu import build_ours, group_grads, rel_l2, trained_like_state
    from morpheus_b200 import train as mtrain
    from morpheus_b200.rays import synthetic_real_view_batch
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
    dev = torch.device('cuda', torch.cuda.current_device())
    dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=90))

    def mark(msg):
        print(f'[rank {rank}] {msg}', flush=True)
    N, S = 512, 64
    sd = trained_like_state(31)
    tr = dict(mtrain.DEFAULT_TRAIN_CFG, surf_sdf_weight=10.0, surf_color_weight=5.0)        # + surface-point terms (data-dependent count)
    batch = {k: v.to(dev) for k, v in synthetic_real_view_batch(N, seed=8, frame=12).items()}
    g = torch.Generator().manual_seed(4)
    jitter = torch.rand(N, generator=g).to(dev)
    noise = torch.randn(N * S, 3, generator=g).to(dev)

    m, R, opt, _ = build_ours(sd, dev, S, 0.9, tr)
    loss_full = mtrain.train_step_compute(R, opt, batch, tr, 1, jitter=jitter, perturb_noise=noise)
    mark('single-GPU gradient done')
    g_full = opt.grad.clone()

    m2, R2, opt2, _ = build_ours(sd, dev, S, 0.9, tr)
    R2.world_size = world
    n = N // world
    shard = {k: v[rank * n:(rank + 1) * n].contiguous() for k, v in batch.items()}
    inj = {'jitter': jitter[rank * n:(rank + 1) * n].contiguous(), 'perturb_noise': noise[rank * n * S:(rank + 1) * n * S].contiguous()}
    loss_shard = mtrain.train_step_compute(R2, opt2, shard, tr, world, **inj)
    opt2.all_reduce()
    torch.cuda.synchronize()
    mark('sharded gradient + all-reduce done')
    errs = {}
    ours = group_grads({nme: p.grad for nme, p in m2.named_parameters() if p.grad is not None})
    ref_named = {}
    for (nme, p), (_, q) in zip(m2.named_parameters(), m.named_parameters()):
        ref_named[nme] = g_full[(q.grad.data_ptr() - opt.grad.data_ptr()) // 4:][:q.numel()].view(q.shape)
    ref = group_grads(ref_named)
    for k in ref:
        if float(ref[k].abs().max()) > 0:
            errs[k] = rel_l2(ours[k], ref[k])
    bad = {k: e for k, e in errs.items() if e > 2e-4}
    mark(f'errors {errs}')
    assert not bad, f'rank {rank}: sharded gradient differs from the single-GPU gradient: {bad}\nall: {errs}'
    # mean over ranks of the shard losses of the mean-type terms ~ full loss (sanity, loose: the shard losses are per-shard means)
    t = loss_shard.clone().reshape(1)
    dist.all_reduce(t)
    # ---- graph replay with the collective captured inside ----
    opt2.step()
    mark('capturing the graphed step')
    gs = mtrain.GraphedStep(R2, opt2, shard, tr, world, inject=inj)
    mark(f'captured, nccl_in_graph={gs.nccl_in_graph}')
    l = gs.step(shard)
    torch.cuda.synchronize()
    assert torch.isfinite(l).all()
    if rank == 0:
        print('groups', {k: f'{v:.1e}' for k, v in errs.items()}, 'nccl_in_graph', gs.nccl_in_graph, 'loss full', float(loss_full), 'mean shard', float(t) / world)
        print('SHARDED_OK', flush=True)
    # (no destroy_process_group: tearing the communicator down while CUDA graphs that captured its collectives are alive can hang)
    del gs
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)


if __name__ == '__main__':
    try:
        main()
    except BaseException:      # noqa: BLE001  -- a rank that fails must not leave its peer waiting in a collective
        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)

"""Oracle: one real-view training step on the CPU (TEST INFRASTRUCTURE / bench cpu_baseline).

Plain-torch restatement of the same step morpheus_b200.train.train_step runs on the GPU:
pose correction -> fixed-S sampling -> scene forward (albedo_normal) -> compositing -> perturbed
normal query -> losses -> autograd backward -> Adam.  Used by bench.py's `cpu_baseline` leg and by
`bench.py --impl reference` (the reference's own CPU path cannot travel to the GPU box, and its
hash-grid kernel is CUDA-only, so the CPU arm is this port; kind = "port").
"""
import torch
import torch.nn.functional as F

from . import render as orr
from .fields import SceneOracle

TRAIN_CFG = {'rgb_weight': 5.0, 'mask_weight': 0.5, 'depth_weight': 0.1, 'sdf_weight': 10.0, 'fs_weight': 0.0,
             'normal_smooth_3d': 0.1, 'smoothness_std': 0.005, 'code_reg': 0.5, 'beta_weight': 0.1, 'trunc': 0.1, 'lr': 5e-4}


def make_params(sd):
    return {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}


def step_loss(params, batch, n_samples, max_level, tr=TRAIN_CFG, num_frames=200, jitter=None, perturb_noise=None):
    """Forward + loss (differentiable).  batch: rays_o/rays_d [N,3], rays_t [N,1], rays_id [N,1], rgb [N,3], depth [N], mask [N], bg [N,3]."""
    sc = SceneOracle(params, 1.01, num_frames, max_level)
    o, d = sc.pose_optimisation(batch['rays_o'], batch['rays_d'], batch['rays_id'])
    N = o.shape[0]
    aabb = torch.tensor([-1.01, -1.01, -1.01, 1.01, 1.01, 1.01])
    if jitter is None:
        jitter = torch.rand(N)
    with torch.no_grad():
        samples = orr.sample_uniform(o.detach(), d.detach(), aabb, n_samples, jitter)
    M = samples[0].shape[0]
    if perturb_noise is None:
        perturb_noise = torch.randn(M, 3)
    light = F.normalize(o.detach() + torch.randn(3), dim=-1)
    out = orr.render_rays(sc, o, d, batch['rays_t'], batch['rays_id'], samples, bg_color=batch['bg'], ambient_ratio=1.0,
                          light_d=light, shading='albedo_normal', rays_depth=batch['depth'], rays_mask=batch['mask'],
                          perturb_noise=perturb_noise, trunc=tr['trunc'], smoothness_std=tr['smoothness_std'], training=True, real_view=True)
    gt_depth, gt_mask = batch['depth'].reshape(-1), batch['mask'].reshape(-1)
    loss = tr['rgb_weight'] * F.mse_loss(out['image'], batch['rgb'])
    loss = loss + tr['mask_weight'] * F.binary_cross_entropy(out['weights_sum'].reshape(-1).clip(1e-5, 1 - 1e-5), gt_mask.to(out['weights_sum'].dtype))
    xyz = batch['rays_o'] + gt_depth[:, None] * batch['rays_d']
    dm = ((gt_depth > 0) & (xyz.norm(dim=-1) <= 1.1) & (gt_mask > 0.5)).float()
    loss = loss + tr['depth_weight'] * F.mse_loss(out['depth'] * dm, gt_depth * dm)
    loss = loss + tr['sdf_weight'] * out['sdf_loss'] + tr['fs_weight'] * out['fs_loss']
    loss = loss + tr['normal_smooth_3d'] * out['loss_normal_perturb']
    ts = batch['rays_t'][:1]
    c0, cm, cp = sc.code(ts), sc.code(ts - 1 / num_frames), sc.code(ts + 1 / num_frames)
    loss = loss + tr['code_reg'] * torch.square(2 * c0 - cm - cp).mean()
    loss = loss + tr['beta_weight'] * (params['sdf2density.beta'].abs() + 1e-4)
    return loss, out

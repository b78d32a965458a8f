"""The render-and-loss hot path: B200-native mirror of `MorpheuS.render_rays`
(/root/reference/morpheus.py:558-794) plus the per-step pieces around it that the bench / smoke
need (ray generation for a pinhole camera, datasets/utils.py:28-65; `get_sdf_loss`, utils.py:91-113;
`update_occ_grid`, morpheus.py:905-913).

Launch count per real-view training step: sampler (2) + field forward (1) + composite (1) + the
perturbed-normal query (1) forward, and as many backward -- versus ~300 eager kernels in the
reference.
"""
import torch

from . import nerfacc_compat as nerfacc
from .model import safe_normalize


def get_camera_rays(H, W, fx, fy=None, cx=None, cy=None, device='cuda'):
    """datasets/utils.py:28-65, OpenGL convention, un-normalised directions [H,W,3]."""
    i, j = torch.meshgrid(torch.arange(W, dtype=torch.float32, device=device), torch.arange(H, dtype=torch.float32, device=device), indexing='xy')
    if cx is None:
        cx, cy = 0.5 * W, 0.5 * H
    if fy is None:
        fy = fx
    return torch.stack([(i + 0.5 - cx) / fx, -(j + 0.5 - cy) / fy, -torch.ones_like(i)], -1)


def get_sdf_loss(z_vals, target_d, predicted_sdf, truncation, mask=None, rays_w_depth=None):
    """utils.py:91-113 (including the per-sample normalisation quirk of sum(dim=-1) on [M,1] tensors).
    `rays_w_depth` overrides the local count_nonzero(target_d) normaliser: under ray sharding it must be the
    GLOBAL count (see global_count), otherwise the summed shard gradients differ from the single-GPU gradient."""
    s = predicted_sdf[..., None]
    depth_mask = target_d > 0.
    front_mask = (z_vals < (target_d - truncation)) | ((target_d < 0.) & (z_vals < 3.5))
    bound = torch.where(target_d < 0., torch.full_like(z_vals, 10.), target_d - z_vals)
    sdf_mask = (bound.abs() <= truncation) & depth_mask
    if mask is not None:
        sdf_mask = sdf_mask & (mask > 0.5)
    n = front_mask.sum(dim=-1) + sdf_mask.sum(dim=-1) + 1e-8
    if rays_w_depth is None:
        rays_w_depth = torch.count_nonzero(target_d)
    fs = torch.max(torch.exp(-5. * s) - 1., s - bound).clamp(min=0.) * front_mask
    fs_loss = (fs.sum(dim=-1) / n).sum() / rays_w_depth
    sdf_loss = ((torch.abs(s - bound) * sdf_mask).sum(dim=-1) / n).sum() / rays_w_depth
    return fs_loss, sdf_loss


def global_count(local_count, world_size):
    """mean-over-ranks of a per-shard count: dividing a shard's SUM by it and the loss by world_size (train_step) yields
    sum / GLOBAL count after the gradient all-reduce.  One extra 4-byte all-reduce per step."""
    c = local_count.to(torch.float32).reshape(1)
    if world_size > 1:
        import torch.distributed as dist
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        c = c / world_size
    return c[0]


class Renderer:
    """Holds what `MorpheuS` holds for rendering: the scene model, the occupancy estimator and the config dict
    (morpheus.py:131-140,196-202).  `render_rays` keeps the reference signature; the extra keyword-only arguments
    (`samples`, `perturb_noise`, `jitter`) let tests inject the RNG draws of SURVEY.md Appendix C."""

    def __init__(self, model, occupancy_grid, config, num_frames, uniform_samples=None):
        self.model, self.occupancy_grid, self.config, self.num_frames = model, occupancy_grid, config, num_frames
        # BASELINE cfg-1/2/4 use a fixed number of samples per ray (synthetic stand-in): when set, the
        # occupancy-grid march is replaced by the fixed-S lattice over the AABB chord (csrc/sampler.cu:uniform_kernel)
        self.uniform_samples = uniform_samples
        self.world_size = 1   # set by the data-parallel driver (bench.py / train.py)
        self.sdf_count_override = None   # 0-dim tensor: pre-reduced global normaliser (lets the step be graph-captured without NCCL inside)

    @torch.no_grad()
    def sample_uniform(self, rays_o, rays_d, S, jitter=None):
        import ctypes as C
        from . import _lib
        N = rays_o.shape[0]
        dev = rays_o.device
        rays_o, rays_d = rays_o.detach().contiguous().float(), rays_d.detach().contiguous().float()
        if jitter is None:
            jitter = torch.rand(N, device=dev)
        jitter = jitter.contiguous().float()
        ri = torch.empty(N * S, dtype=torch.int64, device=dev)
        t0 = torch.empty(N * S, device=dev)
        t1 = torch.empty(N * S, device=dev)
        aabb = (C.c_float * 6)(*[float(v) for v in self.occupancy_grid.aabbs[0].tolist()]) if not hasattr(self, '_aabb_c') else self._aabb_c
        self._aabb_c = aabb
        with _lib.timed('sample_uniform'):
          _lib.check(_lib.lib().mb_sample_rays_uniform(_lib.ptr(rays_o), _lib.ptr(rays_d), N, S, aabb, _lib.ptr(jitter), _lib.ptr(ri),
                                                     _lib.ptr(t0), _lib.ptr(t1), _lib.stream()), 'sample_rays_uniform')
        return ri, t0, t1

    def update_occ_grid(self, rays_t, step, cano=False):
        """morpheus.py:905-913"""
        def occ_eval_fn(x):
            return self.model.density(x, rays_t, allow_shape=True, cano=cano, return_color=False)['sigma'] * self.config['render']['step_size']
        self.occupancy_grid.update_every_n_steps(step=step, occ_eval_fn=occ_eval_fn)

    def render_rays(self, rays_o, rays_d, rays_t, rays_id, H=None, W=None, perturb=True, bg_color=None, ambient_ratio=1.0,
                    light_d=None, shading='albedo', real_view=True, cano=False, rays_depth=None, rays_mask=None,
                    optimize_pose=False, *, samples=None, perturb_noise=None, jitter=None):
        cfg = self.config
        model = self.model
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        rays_t = rays_t.contiguous().view(-1, 1)
        rays_id = rays_id.contiguous().view(-1, 1)
        if not cano and optimize_pose:
            rays_o, rays_d = model.pose_optimisation(rays_o, rays_d, rays_id)
        if rays_depth is not None:
            rays_depth = rays_depth.contiguous().view(-1, 1)
        if rays_mask is not None:
            rays_mask = rays_mask.contiguous().view(-1, 1)
        N = rays_o.shape[0]
        results = {}
        if samples is None and self.uniform_samples:
            samples = self.sample_uniform(rays_o, rays_d, self.uniform_samples, jitter)
        if samples is None:
            with torch.no_grad():
                samples = self.occupancy_grid.sampling(rays_o, rays_d, sigma_fn=None, render_step_size=cfg['render']['step_size'],
                                                       alpha_thre=0, stratified=True, cone_angle=0.0, early_stop_eps=0, jitter=jitter)
        ray_indices, t_starts, t_ends = samples
        ray_indices = ray_indices.long()
        if light_d is None:
            light_d = safe_normalize(rays_o + torch.randn(3, device=rays_o.device))
        t_positions = ((t_starts + t_ends) / 2.0)[..., None]
        xyzs = rays_o[ray_indices] + rays_d[ray_indices] * t_positions
        time_step = rays_t[ray_indices]
        sdf = None
        if xyzs.shape[0] == 0:  # morpheus.py:663-670 (sdf defined here; the reference would raise NameError at :701)
            image = torch.ones([*prefix, 3], device=rays_o.device)
            depth = torch.zeros([*prefix], device=rays_o.device)
            weights = opacity = normals = deform = normal_raw = None
        else:
            light = light_d[ray_indices] if shading != 'albedo' else None
            sdf, sigmas, rgbs, normals, deform, normal_raw = model(xyzs, time_step, light, ratio=ambient_ratio, shading=shading, cano=cano)
            weights, opacity, depth, rgb = nerfacc.composite(sigmas, rgbs, t_starts, t_ends, ray_indices, N)
            opacity = opacity[:, None]
            if bg_color is None:
                if cfg['model']['bg_radius'] > 0 and cano and (not real_view):
                    bg_color = model.background(rays_d, rays_t)
                else:
                    bg_color = 1
            image = (rgb + (1 - opacity) * bg_color).view(*prefix, 3)
            depth = depth.view(*prefix)
        results.update(image=image, depth=depth, sdf=sdf, weights=weights, weights_sum=opacity, normal=normals, deform=deform,
                       normal_raw=normal_raw)
        if model.training and xyzs.shape[0] > 0:
            tr = cfg['train']
            if tr['ori_weight'] > 0 and normals is not None and (not real_view):
                t_dirs = safe_normalize(rays_d[ray_indices])
                results['loss_orient'] = (weights.detach() * (normals * t_dirs).sum(-1).clamp(min=0) ** 2).sum(-1).mean()
            if tr['normal_smooth_3d'] > 0 and normals is not None:
                if tr.get('normal_dir', False) or not tr.get('topo_none', True):
                    raise NotImplementedError('normal_dir / topo_none=False branches are disabled in every shipped config')
                if perturb_noise is None:
                    perturb_noise = torch.randn_like(xyzs)
                xyzs_perturb = xyzs + perturb_noise * tr['smoothness_std']
                normals_perturb, _ = model.normal(xyzs_perturb, topo=None, cano=cano)
                results['loss_normal_perturb'] = (normals - normals_perturb).abs().mean()
            if tr['code_reg'] > 0 and not cano:
                ts = time_step[:1]
                code = model.get_deform_code(ts)
                code_prev = model.get_deform_code(ts - 1 / self.num_frames)
                code_next = model.get_deform_code(ts + 1 / self.num_frames)
                results['loss_code'] = torch.square(2 * code - code_prev - code_next).mean()
            if rays_depth is not None:
                t_gt = rays_depth[ray_indices]
                t_mask = rays_mask[ray_indices] if rays_mask is not None else None
                if self.sdf_count_override is not None:
                    cnt = self.sdf_count_override
                else:
                    cnt = global_count(torch.count_nonzero(t_gt), self.world_size) if self.world_size > 1 else None
                fs_loss, sdf_loss = get_sdf_loss(t_positions, t_gt, sdf, tr['trunc'], mask=t_mask, rays_w_depth=cnt)
                results['sdf_loss'], results['fs_loss'] = sdf_loss, fs_loss
        return results

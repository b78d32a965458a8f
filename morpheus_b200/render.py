"""The render-and-loss hot path: B200-native mirror of `MorpheuS.render_rays`
(/root/reference/morpheus.py:558-794) plus the per-step pieces around it that the bench / smoke
need (ray generation for a pinhole camera, datasets/utils.py:28-65; `get_sdf_loss`, utils.py:91-113;
`update_occ_grid`, morpheus.py:905-913).

Launch count per real-view training step: sampler (2) + field forward (1) + composite (1) + the
perturbed-normal query (1) forward, and as many backward -- versus ~300 eager kernels in the
reference.
"""
import math

import torch
import torch.nn.functional as F

from . import _lib
from . import nerfacc_compat as nerfacc
from ._lib import check, ptr, stream
from .model import safe_normalize


class _RayPoints(torch.autograd.Function):
    """xyzs = rays_o[ray_indices] + rays_d[ray_indices] * (t_starts + t_ends) / 2  (morpheus.py:645-646) in one launch; the
    backward is a warp-per-ray segmented sum (mb_ray_points_backward) instead of two index_put(accumulate) kernels."""

    @staticmethod
    def forward(ctx, rays_o, rays_d, ray_indices, t_starts, t_ends, seg):
        rays_o, rays_d = rays_o.contiguous().float(), rays_d.contiguous().float()
        M = ray_indices.shape[0]
        xyz = torch.empty(M, 3, device=rays_o.device, dtype=torch.float32)
        with _lib.timed('ray_points_fwd'):
            check(_lib.lib().mb_ray_points_forward(ptr(rays_o), ptr(rays_d), ptr(ray_indices), ptr(t_starts), ptr(t_ends), M, ptr(xyz), stream()),
                  'ray_points_forward')
        ctx.save_for_backward(t_starts, t_ends, seg)
        ctx.n_rays = rays_o.shape[0]
        return xyz

    @staticmethod
    def backward(ctx, g_xyz):
        t_starts, t_ends, seg = ctx.saved_tensors
        N = ctx.n_rays
        g_o = torch.empty(N, 3, device=g_xyz.device, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        g_d = torch.empty(N, 3, device=g_xyz.device, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        if g_o is not None or g_d is not None:
            with _lib.timed('ray_points_bwd'):
                check(_lib.lib().mb_ray_points_backward(ptr(seg), N, ptr(t_starts), ptr(t_ends), ptr(g_xyz.contiguous().float()), ptr(g_o), ptr(g_d),
                                                        stream()), 'ray_points_backward')
        return g_o, g_d, None, None, None, None


class _SdfLoss(torch.autograd.Function):
    """utils.get_sdf_loss (utils.py:91-113) over the packed samples in one launch: -> [2] = (sum fs_i/n_i, sum sdf_i/n_i)."""

    @staticmethod
    def forward(ctx, sdf, t_starts, t_ends, ray_indices, depth, mask, trunc):
        sdf = sdf.contiguous().float()
        out = torch.zeros(2, device=sdf.device, dtype=torch.float32)
        with _lib.timed('sdf_loss_fwd'):
            check(_lib.lib().mb_sdf_loss_forward(ptr(t_starts), ptr(t_ends), ptr(ray_indices), ptr(depth), ptr(mask), ptr(sdf), sdf.shape[0],
                                                 _lib.C.c_float(trunc), ptr(out), stream()), 'sdf_loss_forward')
        ctx.save_for_backward(sdf, t_starts, t_ends, ray_indices, depth, mask)
        ctx.trunc = trunc
        return out

    @staticmethod
    def backward(ctx, g_out):
        sdf, t_starts, t_ends, ray_indices, depth, mask = ctx.saved_tensors
        g_sdf = torch.empty_like(sdf)
        with _lib.timed('sdf_loss_bwd'):
            check(_lib.lib().mb_sdf_loss_backward(ptr(t_starts), ptr(t_ends), ptr(ray_indices), ptr(depth), ptr(mask), ptr(sdf), sdf.shape[0],
                                                  _lib.C.c_float(ctx.trunc), ptr(g_out.contiguous().float()), ptr(g_sdf), stream()), 'sdf_loss_backward')
        return g_sdf, None, None, None, None, None, None


def packed_sdf_loss(sdf, t_starts, t_ends, ray_indices, rays_depth, rays_mask, truncation, rays_w_depth=None):
    """get_sdf_loss(z_vals=(t0+t1)/2, target_d=depth[ray], sdf, truncation, mask=mask[ray]) without materialising any
    per-sample intermediate.  rays_w_depth: see get_sdf_loss."""
    depth = rays_depth.reshape(-1).contiguous().float()
    mask = rays_mask.reshape(-1).contiguous().float() if rays_mask is not None else None
    out = _SdfLoss.apply(sdf, t_starts, t_ends, ray_indices, depth, mask, float(truncation))
    if rays_w_depth is None:
        # count_nonzero(target_d) over SAMPLES (utils.py:107): every sample of a ray with depth != 0 counts
        seg_len = torch.bincount(ray_indices, minlength=depth.shape[0])
        rays_w_depth = (seg_len * (depth != 0)).sum()
    out = out / rays_w_depth
    return out[0], out[1]


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

    # -- forward-only consumers -------------------------------------------------------------------
    @torch.no_grad()
    def sdf_volume(self, resolution=128, S=128, t=None, cano=False):
        """The dense SDF grid `MorpheuS.export_mesh` feeds to marching cubes (morpheus.py:384-396): linspace(-1, 1, resolution)^3,
        queried in S^3 blocks through the fused density launch (SDF only, no colour net).  -> [resolution]^3 fp32 on the device."""
        dev = self.model.encoder.embeddings.device
        vol = torch.empty(resolution, resolution, resolution, device=dev)
        axis = torch.linspace(-1, 1, resolution, device=dev).split(S)
        for xi, xs in enumerate(axis):
            for yi, ys in enumerate(axis):
                for zi, zs in enumerate(axis):
                    xx, yy, zz = torch.meshgrid(xs, ys, zs, indexing='ij')
                    pts = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=-1)
                    val = self.model.density(pts, t=t, cano=cano, return_color=False)
                    vol[xi * S: xi * S + len(xs), yi * S: yi * S + len(ys), zi * S: zi * S + len(zs)] = val['sdf'].reshape(len(xs), len(ys), len(zs))
        return vol

    @torch.no_grad()
    def render_image(self, rays_o, rays_d, rays_t, rays_id, chunk=16384, **kw):
        """eval-time render of a full view in ray chunks (morpheus.py:1238-1269 eval_step -> render_rays with perturb=False):
        -> dict(image [N,3], depth [N], weights_sum [N,1])."""
        was_training = self.model.training
        self.model.eval()
        try:
            o, d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
            t, i = rays_t.reshape(-1, 1), rays_id.reshape(-1, 1)
            outs = {'image': [], 'depth': [], 'weights_sum': []}
            for a in range(0, o.shape[0], chunk):
                r = self.render_rays(o[a:a + chunk], d[a:a + chunk], t[a:a + chunk], i[a:a + chunk], perturb=False, **kw)
                n = r['image'].reshape(-1, 3).shape[0]
                outs['image'].append(r['image'].reshape(-1, 3))
                outs['depth'].append(r['depth'].reshape(-1))
                ws = r['weights_sum']
                outs['weights_sum'].append(ws.reshape(-1, 1) if ws is not None else torch.zeros(n, 1, device=o.device))
            return {k: torch.cat(v, 0) for k, v in outs.items()}
        finally:
            self.model.train(was_training)

    @staticmethod
    def get_ortho_normal_dir(normals, phi=None):
        """morpheus.py:518-528: a random unit direction orthogonal to the normal (`phi` [.., 1] in [0, 2 pi) injects the draw)"""
        n = F.normalize(normals, dim=-1)
        u = F.normalize(torch.stack([n[..., 1], -n[..., 0], torch.zeros_like(n[..., 2])], dim=-1), dim=-1)   # n[..., [1,0,2]] * (1,-1,0), capture-safe
        v = torch.cross(n, u, dim=-1)
        if phi is None:
            phi = torch.rand(list(normals.shape[:-1]) + [1], device=normals.device) * 2. * math.pi
        return torch.cos(phi) * u + torch.sin(phi) * v

    def get_normal_smoothness_loss(self, rays_o, rays_d, rays_t, depth, *, trunc_noise=None, phi=None):
        """morpheus.py:530-556 (L_smooth in observation space): 11 points per ray in a band around the rendered depth, the
        fused `normal(x, t)` query (deform + topology nets + 6 warped SDF queries, ONE launch) at each point and at a point
        displaced by smoothness_std along a random tangent, mean squared normal difference.  The reference drops the points
        with |x| >= 1.1 by boolean indexing (a data-dependent shape: host sync, not graph-capturable); here every point is
        evaluated and the dropped ones get weight 0 -- the same mean over the kept points."""
        tr = self.config['train']
        n_pts = int(tr['trunc'] * 100 + 1)
        dev = rays_o.device
        trunc_normal = torch.linspace(-0.5 * tr['trunc'], 0.5 * tr['trunc'], n_pts, device=dev)
        if trunc_noise is None:
            trunc_noise = torch.rand_like(trunc_normal)
        trunc_normal = trunc_normal + 0.01 * trunc_noise
        depth = depth.reshape(1, -1)
        surf_pts = ((depth + trunc_normal[:, None])[..., None] * rays_d[None, ...] + rays_o[None, ...]).reshape(-1, 3)
        surf_t = rays_t.reshape(1, -1, 1).repeat(n_pts, 1, 1).reshape(-1, 1)
        keep = (torch.linalg.norm(surf_pts.detach(), ord=2, dim=-1) < 1.1).to(surf_pts.dtype)
        n1, _ = self.model.normal(surf_pts, t=surf_t)
        w = self.get_ortho_normal_dir(n1, phi)
        n2, _ = self.model.normal(surf_pts + w * tr['smoothness_std'], t=surf_t)
        return (torch.square(n1 - n2) * keep[:, None]).sum() / (3.0 * self._gcount(keep.sum()).clamp(min=1.0))

    def _gcount(self, local_count):
        """data-dependent normaliser under ray sharding: mean over ranks (global_count), so that loss / world_size summed over the
        ranks is sum / GLOBAL count; the identity on one GPU"""
        return global_count(local_count, self.world_size) if self.world_size > 1 else local_count

    def _mean_over_samples(self, per_sample_sum, M, fixed):
        """mean over the packed samples of this shard; with the occupancy sampler M differs per rank, so the divisor is the
        rank-mean sample count (fixed-S sampling: M is the same on every rank and no collective is needed)"""
        if self.world_size > 1 and not fixed:
            return per_sample_sum / self._gcount(torch.tensor(float(M), device=per_sample_sum.device))
        return per_sample_sum / M

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
        used_uniform = samples is None and bool(self.uniform_samples)      # exactly S samples per ray (the count shortcut below relies on it)
        if used_uniform:
            samples = self.sample_uniform(rays_o, rays_d, self.uniform_samples, jitter)
        if samples is None:
            with torch.no_grad():
                samples = self.occupancy_grid.sampling(rays_o, rays_d, sigma_fn=None, render_step_size=cfg['render']['step_size'],
                                                       alpha_thre=0, stratified=True, cone_angle=0.0, early_stop_eps=0, jitter=jitter)
        ray_indices, t_starts, t_ends = samples
        ray_indices = ray_indices.long()
        # the light direction only matters when the colour depends on the normal (morpheus.py:641; model.py:523-529): with 'albedo' and with
        # 'albedo_normal' at ratio 1 (every real view) the lambertian factor is exactly 1, so the draw and its gather are skipped
        needs_light = shading not in ('albedo',) and not (shading == 'albedo_normal' and float(ambient_ratio) == 1.0)
        if light_d is None and needs_light:
            light_d = safe_normalize(rays_o + torch.randn(3, device=rays_o.device))
        t_starts, t_ends = t_starts.contiguous().float(), t_ends.contiguous().float()
        ray_indices = ray_indices.contiguous()
        seg = nerfacc.ray_segments(ray_indices, N) if ray_indices.shape[0] > 0 else None
        xyzs = (_RayPoints.apply(rays_o, rays_d, ray_indices, t_starts, t_ends, seg) if ray_indices.shape[0] > 0
                else torch.zeros(0, 3, device=rays_o.device))
        time_step = rays_t[ray_indices]
        sdf = None
        if xyzs.shape[0] == 0:  # morpheus.py:663-670 (sdf defined here; the reference would raise NameError at :701)
            image = torch.ones([*prefix, 3], device=rays_o.device)
            depth = torch.zeros([*prefix], device=rays_o.device)
            weights = opacity = normals = deform = normal_raw = None
        else:
            light = light_d[ray_indices] if (shading != 'albedo' and light_d is not None) else None
            tr_ = cfg.get('train', {}) if model.training else {}
            # real-view training step: with 'albedo_normal' (ratio 1) the colour does not depend on the normal, so the two FD-normal
            # sets of a sample (at x, and at the perturbed point) only feed loss_normal_perturb: ONE fused forward+backward launch
            fused_fd = (_lib.USE_TC and _lib.USE_FD_REG and model.training and torch.is_grad_enabled() and real_view and not cano
                        and shading == 'albedo_normal' and float(ambient_ratio) == 1.0 and tr_.get('normal_smooth_3d', 0) > 0
                        and not tr_.get('normal_dir', False) and tr_.get('topo_none', True))
            if fused_fd:
                sdf, sigmas, rgbs, deform, topo_w = model.forward_with_topo(xyzs, time_step)
                if perturb_noise is None:
                    perturb_noise = torch.randn_like(xyzs)
                M_ = xyzs.shape[0]
                l_np, normals, normal_raw = model.fd_regulariser(xyzs, topo_w, perturb_noise, tr_['smoothness_std'], 1.0 / (3.0 * M_))
                if self.world_size > 1 and not used_uniform:      # ragged shards: mean over the GLOBAL sample count
                    l_np = l_np * (float(M_) / self._gcount(torch.tensor(float(M_), device=xyzs.device)))
                results['loss_normal_perturb'] = l_np
            else:
                sdf, sigmas, rgbs, normals, deform, normal_raw = model(xyzs, time_step, light, ratio=ambient_ratio, shading=shading, cano=cano)
            weights, opacity, depth, rgb = nerfacc.composite(sigmas, rgbs, t_starts, t_ends, ray_indices, N, seg=seg)
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
                # `.sum(-1).mean()` on the [M] tensor (morpheus.py:712) is the SUM over all samples: a shard contributes
                # world_size x its local sum, so that (loss / world_size) summed over the ranks is the global sum
                results['loss_orient'] = (weights.detach() * (normals * t_dirs).sum(-1).clamp(min=0) ** 2).sum(-1).mean() * self.world_size
            if tr['normal_smooth_3d'] > 0 and normals is not None and 'loss_normal_perturb' not in results:
                if tr.get('normal_dir', False) or not tr.get('topo_none', True):
                    raise NotImplementedError('normal_dir / topo_none=False branches are disabled in every shipped config')
                if perturb_noise is None:
                    perturb_noise = torch.randn_like(xyzs)
                xyzs_perturb = xyzs + perturb_noise * tr['smoothness_std']
                normals_perturb, _ = model.normal(xyzs_perturb, topo=None, cano=cano)
                results['loss_normal_perturb'] = self._mean_over_samples((normals - normals_perturb).abs().sum(), 3 * xyzs.shape[0], used_uniform)
            if tr['code_reg'] > 0 and not cano:
                results['loss_code'] = model.code_regulariser(time_step[:1], self.num_frames)      # morpheus.py:762-771, one launch
            if tr.get('normal_smoothness', 0) > 0:
                results['normal_reg'] = self.get_normal_smoothness_loss(rays_o, rays_d, rays_t, depth)      # morpheus.py:778-785
            if rays_depth is not None:
                if self.sdf_count_override is not None:
                    cnt = self.sdf_count_override
                elif used_uniform:
                    cnt = torch.count_nonzero(rays_depth) * self.uniform_samples      # == count_nonzero(depth[ray_indices])
                    if self.world_size > 1:
                        cnt = global_count(cnt, self.world_size)
                else:
                    cnt = global_count(torch.count_nonzero(rays_depth[ray_indices]), self.world_size) if self.world_size > 1 else None
                fs_loss, sdf_loss = packed_sdf_loss(sdf, t_starts, t_ends, ray_indices, rays_depth, rays_mask, tr['trunc'], rays_w_depth=cnt)
                results['sdf_loss'], results['fs_loss'] = sdf_loss, fs_loss
        return results

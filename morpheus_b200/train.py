"""Per-iteration training step around render_rays: loss heads that consume the render outputs
(morpheus.py:946-983 get_real_view_render_loss, :985-999 sdf part of get_real_view_point_loss,
:1090-1145 get_regularization_loss -- the terms active with the shipped weights),
a flat parameter/gradient arena with one fused Adam launch (morpheus.py:154-155: Adam(betas=(.9,.99),
eps=1e-15) over the named groups of models/model.py:313-324), and the data-parallel exchange: one
NCCL all-reduce of the flat gradient arena per step (the reference is single-GPU).
"""
import ctypes as C
import math
import random

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import _lib
from ._lib import check, ptr, stream

DEFAULT_TRAIN_CFG = {  # configs/snoopy.yaml:40-94 (only what the step reads)
    'rgb_weight': 5.0, 'mask_weight': 0.5, 'depth_weight': 0.1, 'sdf_weight': 10.0, 'fs_weight': 0.0,
    'normal_smooth_3d': 0.1, 'smoothness_std': 0.005, 'topo_none': True, 'normal_dir': False, 'code_reg': 0.5,
    'beta_weight': 0.1, 'ori_weight': 0.01, 'trunc': 0.1, 'lr': 5e-4,
    # SURVEY 8f rank 1 (off in the BASELINE cfg-2 'mode B' step; FULL_TRAIN_CFG switches them on with the shipped weights)
    'normal_smoothness': 0.0, 'surf_sdf_weight': 0.0, 'surf_color_weight': 0.0,
    'albedo_iter_ratio': 0.1, 'min_ambient_ratio': 0.1, 'textureless_ratio': 0.2, 'warm_up_end': 200, 'n_epochs': 2000, 'ema_decay': 0.95,
}
FULL_TRAIN_CFG = dict(DEFAULT_TRAIN_CFG, normal_smoothness=0.4, surf_sdf_weight=10.0, surf_color_weight=5.0)   # configs/snoopy.yaml:70-74


def surface_point_loss(model, batch, tr, world_size=1):
    """get_real_view_point_loss, surface terms (morpheus.py:1005-1029): one fused density query (warp + SDF + colour) at the
    back-projected depth points; SDF^2 averaged over the valid points, colour MSE over all points with invalid ones zeroed.
    Under ray sharding the valid-point count is the mean over ranks (render.global_count), so that the summed shard gradients
    equal the single-GPU gradient."""
    from .render import global_count
    gt_depth, gt_mask = batch['depth'].reshape(-1), batch['mask'].reshape(-1)
    xyz = batch['rays_o'].reshape(-1, 3) + gt_depth[:, None] * batch['rays_d'].reshape(-1, 3)
    dm = ((gt_depth > 0) & (xyz.norm(dim=-1) <= 1.1) & (gt_mask > 0.5)).float()
    res = model.density(xyz, t=batch['rays_t'].reshape(-1, 1))
    n_valid = global_count(dm.sum(), world_size) if world_size > 1 else dm.sum()
    surf_sdf = tr['surf_sdf_weight'] * (res['sdf'].square() * dm).sum() / n_valid.clamp(min=1.0)
    surf_col = tr['surf_color_weight'] * F.mse_loss(res['albedo'] * dm[:, None], batch['rgb'] * dm[:, None])
    return surf_sdf + surf_col


class _RayLoss(torch.autograd.Function):
    """rgb MSE + mask BCE + masked depth MSE of get_real_view_render_loss (morpheus.py:946-983) in one launch; the kernel also
    emits the per-ray gradients, so the backward is three scalings."""

    @staticmethod
    def forward(ctx, image, opacity, depth, gt_rgb, gt_depth, gt_mask, rays_o, rays_d, w_rgb, w_mask, w_depth):
        image, opacity, depth = image.contiguous().float(), opacity.contiguous().float(), depth.contiguous().float()
        N = opacity.shape[0]
        out = torch.zeros(1, device=image.device, dtype=torch.float32)
        g_i, g_o, g_d = torch.empty_like(image), torch.empty_like(opacity), torch.empty_like(depth)
        with _lib.timed('ray_loss'):
            check(_lib.lib().mb_ray_loss(ptr(image), ptr(opacity), ptr(depth), ptr(gt_rgb.contiguous().float()), ptr(gt_depth.contiguous().float()),
                                         ptr(gt_mask.contiguous().float()), ptr(rays_o.contiguous().float()), ptr(rays_d.contiguous().float()), N,
                                         C.c_float(w_rgb), C.c_float(w_mask), C.c_float(w_depth), ptr(out), ptr(g_i), ptr(g_o), ptr(g_d), stream()),
                  'ray_loss')
        ctx.save_for_backward(g_i, g_o, g_d)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        g_i, g_o, g_d = ctx.saved_tensors
        return (g_i * g, g_o * g, g_d * g) + (None,) * 8


class _WeightedSum(torch.autograd.Function):
    """sum_i w_i * term_i over 0-dim tensors with host-constant weights in 3 launches (cat, mul, sum) and ONE backward launch, instead of
    two eager launches per term and direction: the loss assembly of a step was ~35 of its ~85 tiny launches, a fixed cost that does not
    shrink with the number of GPUs (strong scaling)."""

    @staticmethod
    def forward(ctx, wvec, *terms):
        ctx.save_for_backward(wvec)
        ctx.n = len(terms)
        return (torch.stack([t.reshape(()) for t in terms]) * wvec).sum()

    @staticmethod
    def backward(ctx, g):
        wvec, = ctx.saved_tensors
        gw = g * wvec
        return (None,) + tuple(gw[i] for i in range(ctx.n))


_WVEC_CACHE = {}


def weighted_sum(terms, weights):
    """terms: 0-dim tensors, weights: python floats (zero-weight terms are dropped)"""
    keep = [(t, float(w)) for t, w in zip(terms, weights) if float(w) != 0.0]
    dev = keep[0][0].device
    key = (str(dev), tuple(w for _, w in keep))
    if key not in _WVEC_CACHE:
        _WVEC_CACHE[key] = torch.tensor([w for _, w in keep], dtype=torch.float32, device=dev)
    return _WeightedSum.apply(_WVEC_CACHE[key], *[t for t, _ in keep])


def real_view_loss(out, batch, model, tr, world_size=1, scale=1.0):
    """rgb MSE x5 + mask BCE x.5 + masked depth MSE x.1 (morpheus.py:946-983, one fused launch) + sdf band loss x10 (:991-992)
    + normal_smooth_3d x.1 + code_reg x.5 + beta x.1 (:1116-1142), assembled by one weighted-sum op.  `scale` multiplies the total
    (1 / world_size under ray sharding) without an extra launch."""
    pred_rgb = out['image'].reshape(-1, 3)
    pred_depth = out['depth'].reshape(-1)
    pred_mask = out['weights_sum'].reshape(-1)
    gt_depth = batch['depth'].reshape(-1)
    terms = [_RayLoss.apply(pred_rgb, pred_mask, pred_depth, batch['rgb'], gt_depth, batch['mask'].reshape(-1), batch['rays_o'].reshape(-1, 3),
                            batch['rays_d'].reshape(-1, 3), float(tr['rgb_weight']), float(tr['mask_weight']), float(tr['depth_weight']))]
    weights = [1.0]
    if 'sdf_loss' in out:
        terms += [out['sdf_loss'], out['fs_loss']]
        weights += [tr['sdf_weight'], tr['fs_weight']]
    for key, w in (('loss_normal_perturb', 'normal_smooth_3d'), ('loss_code', 'code_reg'), ('normal_reg', 'normal_smoothness')):
        if key in out:
            terms.append(out[key])
            weights.append(tr[w])
    if tr.get('surf_sdf_weight', 0) > 0:
        terms.append(surface_point_loss(model, batch, tr, world_size))
        weights.append(1.0)
    terms.append(model.sdf2density.get_beta())          # mean of a 0-dim tensor is the tensor itself (morpheus.py:1124-1125)
    weights.append(tr['beta_weight'])
    return weighted_sum(terms, [w * scale for w in weights])


def real_view_loss_torch(out, batch, model, tr, world_size=1):
    """the same loss as eager torch ops (reference formulation; used by the parity tests)"""
    pred_rgb = out['image'].reshape(-1, 3)
    pred_depth = out['depth'].reshape(-1)
    pred_mask = out['weights_sum'].reshape(-1)
    gt_rgb, gt_depth, gt_mask = batch['rgb'], batch['depth'].reshape(-1), batch['mask'].reshape(-1)
    loss = tr['rgb_weight'] * F.mse_loss(pred_rgb, gt_rgb)
    loss = loss + tr['mask_weight'] * F.binary_cross_entropy(pred_mask.clip(1e-5, 1.0 - 1e-5), gt_mask.float())
    xyz = batch['rays_o'].reshape(-1, 3) + gt_depth[:, None] * batch['rays_d'].reshape(-1, 3)
    depth_mask = ((gt_depth > 0) & (xyz.norm(dim=-1) <= 1.1) & (gt_mask > 0.5)).float()
    loss = loss + tr['depth_weight'] * F.mse_loss(pred_depth * depth_mask, gt_depth * depth_mask)
    if 'sdf_loss' in out:
        loss = loss + tr['sdf_weight'] * out['sdf_loss'] + tr['fs_weight'] * out['fs_loss']
    if 'loss_normal_perturb' in out:
        loss = loss + tr['normal_smooth_3d'] * out['loss_normal_perturb']
    if 'loss_code' in out:
        loss = loss + tr['code_reg'] * out['loss_code']
    if 'normal_reg' in out:
        loss = loss + tr['normal_smoothness'] * out['normal_reg']
    if tr.get('surf_sdf_weight', 0) > 0:
        loss = loss + surface_point_loss(model, batch, tr, world_size)
    loss = loss + tr['beta_weight'] * torch.mean(model.sdf2density.get_beta())
    return loss


def progressive_level(exp_iter_ratio, enabled=True):
    """morpheus.py:808-813: coarse-to-fine level of the hash grids / frequency bands, exp_iter_ratio = epoch / n_epochs"""
    return min(1.0, 0.5 + 0.5 * exp_iter_ratio) if enabled else None


def get_shading(exp_iter_ratio, real_view, tr, rng=random):
    """morpheus.py:865-888 -> (ambient_ratio, shading).  Real views render 'albedo_normal' with ratio 1; virtual views use albedo
    during the first `albedo_iter_ratio` of training, then a random ambient ratio and lambertian / textureless shading (same two
    random.random() draws, in the same order, as the reference)."""
    if real_view:
        return 1.0, 'albedo_normal'
    if exp_iter_ratio <= tr['albedo_iter_ratio']:
        return 1.0, 'albedo'
    ambient_ratio = tr['min_ambient_ratio'] + (1.0 - tr['min_ambient_ratio']) * rng.random()
    shading = 'textureless' if rng.random() >= (1.0 - tr['textureless_ratio']) else 'lambertian'
    return ambient_ratio, shading


def get_bg_color(real_view, B, N, device, bg_radius=1.4, rng=random, generator=None):
    """morpheus.py:890-903: per-ray random background on real views (also blended into the ground-truth image outside the mask,
    :942); on virtual views None (= background network, only with cano) with probability 1/2, else one random colour"""
    if real_view:
        return torch.rand((B * N, 3), generator=generator).to(device)
    if bg_radius > 0 and rng.random() > 0.5:
        return None
    return torch.rand(3, generator=generator).to(device)


def blend_gt_background(gt_rgb, gt_mask, bg_color):
    """get_gt_from_data, morpheus.py:939-942 on per-ray tensors: mask binarised at 0.5, gt = rgb * mask + bg * (1 - mask)"""
    m = (gt_mask > 0.5).to(gt_rgb.dtype).reshape(-1, 1)
    return gt_rgb * m + bg_color * (1 - m), m.reshape(-1)


def learning_factor(epoch, warm_up_end, n_epochs, scale_factor=1.0):
    """morpheus.py:476-486"""
    if epoch < warm_up_end:
        f = 0.01 if epoch < 100 else 0.01 + (epoch - 100) / (warm_up_end - 100) * 0.99
    else:
        alpha = 0.05
        progress = (epoch - warm_up_end) / (n_epochs - warm_up_end)
        f = (math.cos(math.pi * progress) + 1.0) * 0.5 * (1 - alpha) + alpha
    return f * scale_factor


class FlatEMA:
    """torch_ema.ExponentialMovingAverage (requirements.txt:24; morpheus.py:160-162, :1299-1301, :1368-1369, :1432-1433) over the
    ONE flat parameter buffer of FlatAdam: update() is a single lerp launch instead of one sub_ per parameter tensor."""

    def __init__(self, flat, decay, use_num_updates=True, model=None):
        # `model`: the scene_representation whose parameters are views of `flat`; copy_to / restore / load_state_dict rewrite the
        # buffer through raw storage (no version-counter bump), so its cached packed arena must be dropped explicitly
        self.flat, self.decay, self.model = flat, float(decay), model
        self.num_updates = 0 if use_num_updates else None
        self.shadow = flat.detach().clone()
        self.collected = None

    def update(self):
        decay = self.decay
        if self.num_updates is not None:
            self.num_updates += 1
            decay = min(decay, (1 + self.num_updates) / (10 + self.num_updates))
        self.shadow.lerp_(self.flat, 1.0 - decay)          # shadow -= (1 - decay) * (shadow - param)

    def store(self):
        self.collected = self.flat.detach().clone()

    def _changed(self):
        if self.model is not None:
            self.model.invalidate()

    def copy_to(self):
        self.flat.copy_(self.shadow)
        self._changed()

    def restore(self):
        self.flat.copy_(self.collected)
        self.collected = None
        self._changed()

    def state_dict(self):
        return {'decay': self.decay, 'num_updates': self.num_updates, 'shadow': self.shadow, 'collected': self.collected}

    def load_state_dict(self, sd):
        self.decay, self.num_updates = sd['decay'], sd['num_updates']
        self.shadow.copy_(sd['shadow'])
        self.collected = sd['collected']


class FlatAdam:
    """All trainable parameters re-homed into ONE flat fp32 buffer (and their .grad into one flat gradient
    buffer), so a step is: one all-reduce, one fused Adam launch (csrc/sampler.cu:adam_groups_kernel) that also clears the
    gradient buffer for the next step.  Per-group learning rates follow get_params_all() (models/model.py:313-324).

    torch.optim.Adam semantics per group (the reference runs torch 2.0: zero_grad() sets .grad to None and Adam SKIPS parameters
    without a gradient -- no moment decay, no step increment, no move): `set_active()` marks the groups that receive a gradient in
    the coming step; inactive groups are left untouched and every group carries its own step count for the bias corrections."""

    DEFORM_GROUPS = ('code_deform', 'decoder_deform', 'decoder_topo')
    # groups that never see a gradient with the shipped flags: bg_net is only reached with cano=True
    NEVER = ('decoder_bg',)

    def __init__(self, model, lr, betas=(0.9, 0.99), eps=1e-15):
        groups = model.get_params_all(lr)
        plist, gid, lrs = [], [], []
        for gi, g in enumerate(groups):
            lrs.append(g['lr'])
            for p in g['params']:
                plist.append(p)
                gid.append(gi)
        n = sum(p.numel() for p in plist)
        dev = plist[0].device
        self.flat = torch.empty(n, device=dev)
        self.grad = torch.zeros(n, device=dev)
        self.m = torch.zeros(n, device=dev)
        self.v = torch.zeros(n, device=dev)
        self.group_id = torch.empty(n, dtype=torch.uint8, device=dev)
        off = 0
        self.group_slices = {}
        for p, gi in zip(plist, gid):
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view(p.shape)
            p.grad = self.grad[off:off + k].view(p.shape)
            self.group_id[off:off + k] = gi
            a, b = self.group_slices.get(groups[gi]['name'], (off, off))
            self.group_slices[groups[gi]['name']] = (min(a, off), off + k)
            off += k
        self.group_names = [g['name'] for g in groups]
        self.group_lr = torch.tensor(lrs, device=dev, dtype=torch.float32)
        self.group_active = torch.ones(len(groups), dtype=torch.uint8, device=dev)
        self.group_step = torch.zeros(len(groups), dtype=torch.int32, device=dev)     # device-side per-group step counts (graph-replayable)
        self._active_host = None
        self.betas, self.eps, self.t, self.n = betas, eps, 0, n
        self.params = plist
        self.model = model
        model.grad_sink = True      # every .grad is a view of self.grad: the kernels accumulate into it directly
        self.current_learning_rate = lr
        self._clean = True          # the gradient buffer is all zeros (fresh, or cleared by the last fused step)
        self.set_active(real_view=True)

    def set_group_lr(self, name, lr):
        self.group_lr[self.group_names.index(name)] = lr

    def set_active(self, real_view=True, shading='albedo_normal', optimize_pose=None):
        """Which groups receive a gradient in the coming step (everything else keeps .grad = None in the reference and is skipped by
        torch.optim.Adam): the pose correction only with optimize_pose (real views, morpheus.py:1418 vs :1399), the colour grid /
        decoder not under 'textureless' shading (models/model.py:527-529: the albedo does not reach the output), bg_net never."""
        if optimize_pose is None:
            optimize_pose = real_view
        off = set(self.NEVER)
        if not optimize_pose:
            off.add('pose')
        if shading == 'textureless':
            off.update(('encoder_color', 'decoder_color'))
        host = tuple(0 if n in off else 1 for n in self.group_names)
        if host != self._active_host:
            self.group_active.copy_(torch.tensor(host, dtype=torch.uint8))
            self._active_host = host

    # -- learning-rate schedule of the reference trainer (morpheus.py:471-516), on the device-resident per-group table ------
    def update_learning_rate(self, epoch, tr, scale_factor=1.0):
        """morpheus.py:471-502: 0.01 -> 1 warm-up, then cosine to 0.05; 'pose' runs at lr/10, every other group at lr"""
        self.current_learning_rate = tr['lr'] * learning_factor(epoch, tr['warm_up_end'], tr['n_epochs'], scale_factor)
        lrs = [self.current_learning_rate * (0.1 if n == 'pose' else 1.0) for n in self.group_names]
        self.group_lr.copy_(torch.tensor(lrs, dtype=torch.float32))
        return self.current_learning_rate

    def freeze_lr_deform(self):
        """morpheus.py:504-511 (virtual steps while epoch <= 400 do not move the deformation field)"""
        for n in self.DEFORM_GROUPS:
            self.set_group_lr(n, 0.0)

    def reset_lr_deform(self):
        """morpheus.py:513-516"""
        for n in self.DEFORM_GROUPS:
            self.set_group_lr(n, self.current_learning_rate)

    def zero_grad(self):
        """no launch when the previous fused step already cleared the buffer"""
        if not self._clean:
            self.grad.zero_()
        self._clean = False

    def all_reduce(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM)

    def step(self):
        self.t += 1
        with _lib.timed('adam'):
            check(_lib.lib().mb_adam_step_groups(ptr(self.flat), ptr(self.grad), ptr(self.m), ptr(self.v), ptr(self.group_id), ptr(self.group_lr),
                                                 ptr(self.group_active), ptr(self.group_step), len(self.group_names), C.c_uint64(self.n),
                                                 C.c_float(self.betas[0]), C.c_float(self.betas[1]), C.c_float(self.eps), 1, stream()), 'adam_step_groups')
        self._clean = True
        self.model.invalidate()    # parameters changed through raw pointers: version counters cannot see it


def train_step(renderer, opt, batch, tr, world_size=1, shading='albedo_normal', samples=None, **inject):
    """One real-view optimiser step (morpheus.py:1147-1236 + :1415-1424) on this rank's shard of the ray batch.
    Loss terms are means over the GLOBAL batch (fixed-size terms divide by world_size, data-dependent counts go through
    render.global_count) so that the summed all-reduce equals the single-GPU gradient.  `inject`: RNG draws for parity tests
    (jitter, perturb_noise, light_d)."""
    loss = train_step_compute(renderer, opt, batch, tr, world_size, shading, samples=samples, **inject)
    opt.all_reduce()
    opt.step()
    return loss


def train_step_compute(renderer, opt, batch, tr, world_size=1, shading='albedo_normal', samples=None, **inject):
    """zero-grad + render + loss + backward (no collective, no optimiser)"""
    opt.set_active(real_view=True, shading=shading)
    opt.zero_grad()
    out = renderer.render_rays(batch['rays_o'], batch['rays_d'], batch['rays_t'], batch['rays_id'], bg_color=batch['bg'],
                               shading=shading, real_view=True, rays_depth=batch['depth'], rays_mask=batch['mask'], optimize_pose=True,
                               samples=samples, **inject)
    loss = real_view_loss(out, batch, renderer.model, tr, world_size, scale=1.0 / world_size)
    loss.backward()
    return loss.detach() * world_size if world_size > 1 else loss.detach()


def virtual_view_loss_terms(out, model, tr):
    """get_regularization_loss on a virtual view (morpheus.py:1090-1145 with the shipped weights): orientation
    x ori_weight (:1119-1120), normal_smooth_3d (:1116-1117), normal_smoothness x normal_reg (:1127-1128 -- NOT gated on real_view),
    code_reg (:1139-1140), beta (:1124-1125)."""
    loss = tr['beta_weight'] * torch.mean(model.sdf2density.get_beta())
    if 'loss_orient' in out:
        loss = loss + tr['ori_weight'] * out['loss_orient']
    if 'loss_normal_perturb' in out:
        loss = loss + tr['normal_smooth_3d'] * out['loss_normal_perturb']
    if 'normal_reg' in out:
        loss = loss + tr['normal_smoothness'] * out['normal_reg']
    if 'loss_code' in out:
        loss = loss + tr['code_reg'] * out['loss_code']
    return loss


class _AllGatherRows(torch.autograd.Function):
    """[n, C] shard of every rank -> [world * n, C] on every rank (NCCL all-gather); backward hands each rank the rows it contributed.
    Used for the novel-view image: the rays of the view are rendered ray-sharded, the 62 KB image is gathered, and the (replicated)
    SDS chain runs on the full image."""

    @staticmethod
    def forward(ctx, x, world, rank):
        x = x.contiguous()
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)      # concatenated form: NCCL and gloo
        dist.all_gather_into_tensor(out, x)
        ctx.rank, ctx.n = rank, x.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        return g[ctx.rank * ctx.n:(ctx.rank + 1) * ctx.n].contiguous(), None, None


def virtual_view_step(renderer, guidance, opt, view, embeddings, tr, shading='lambertian', ambient_ratio=0.5, bg_color=None,
                      guidance_scale=5.0, grad_weight=0.01, t=None, noise=None, vae_noise=None, light_d=None, world_size=1):
    """One novel-view (SDS) optimiser step: morpheus.py:1147-1236 with real_view=False + get_virtual_view_loss (:1044-1088,
    single reference view) + the regularisers that are live on virtual views (virtual_view_loss_terms).  `view` comes from
    rays.virtual_view_rays.  world_size > 1: every rank renders a contiguous slice of the view's rays, the image is all-gathered and
    the replicated SDS chain runs on the full image (its gradient reaches each rank's slice unscaled; the regularisers are per-shard
    means and are divided by world_size), then one all-reduce of the flat gradient buffer."""
    model = renderer.model
    opt.set_active(real_view=False, shading=shading)
    opt.zero_grad()
    H, W = view['H'], view['W']
    rays = {k: view[k].reshape(1, H * W, -1) for k in ('rays_o', 'rays_d', 'rays_t', 'rays_id')}
    rank = 0
    if world_size > 1:
        rank = dist.get_rank()
        n = (H * W) // world_size
        if n * world_size != H * W:
            raise RuntimeError(f'virtual_view_step: {H}x{W} rays do not split evenly over {world_size} ranks')
        rays = {k: v[:, rank * n:(rank + 1) * n] for k, v in rays.items()}
        if t is None:       # the replicated SDS chain must draw the SAME timestep / noise on every rank
            g = torch.Generator(device=rays['rays_o'].device).manual_seed(int(opt.t) + 12345)
            t = torch.randint(guidance.min_step, guidance.max_step + 1, (1,), dtype=torch.long, device=rays['rays_o'].device, generator=g)
            noise = torch.randn(1, 4, 32, 32, device=rays['rays_o'].device, generator=g) if noise is None else noise
            vae_noise = torch.randn(1, 4, 32, 32, device=rays['rays_o'].device, generator=g) if vae_noise is None else vae_noise
    out = renderer.render_rays(rays['rays_o'], rays['rays_d'], rays['rays_t'], rays['rays_id'], H, W, bg_color=bg_color,
                               ambient_ratio=ambient_ratio, light_d=light_d, shading=shading, real_view=False, optimize_pose=False)
    image = out['image'].reshape(-1, 3)
    if world_size > 1:
        image = _AllGatherRows.apply(image, world_size, rank)
    pred_rgb = image.reshape(1, H, W, 3).permute(0, 3, 1, 2).contiguous()
    loss_sds, t, grad_scale, noise = guidance.train_step(embeddings, pred_rgb, view['polar'], view['azimuth'], view['radius'],
                                                         guidance_scale=guidance_scale, grad_scale=grad_weight, t=t, noise=noise, vae_noise=vae_noise)
    reg = virtual_view_loss_terms(out, model, tr)
    (loss_sds + reg / world_size).backward()
    opt.all_reduce()
    opt.step()
    return (loss_sds + reg).detach(), out


class GraphedStep:
    """The real-view optimiser step captured as ONE CUDA graph and replayed per iteration: render + losses + backward, the NCCL
    all-reduce of the flat gradient buffer (NCCL collectives are stream-ordered and capturable) and the fused Adam (which also
    clears the gradient buffer for the next replay).  Inputs live in static device buffers (`self.batch`); shapes are fixed
    (fixed-S sampler), RNG draws inside the step use torch's graph-safe Philox offsets, the Adam step counts are device-resident.

    Constraint: kernel arguments passed BY VALUE are baked into the graph.  The coarse-to-fine level (`model.max_level` ->
    n_levels / n_freq, morpheus.py:808-813 rewrites it every step) is such an argument: `step()` compares `model._levels()` with
    the captured value and RE-CAPTURES when it changed (8 -> 16 grid levels / 3 -> 6 bands: at most 11 re-captures per training).
    MORPHEUS_B200_GRAPH_NCCL=0 keeps the collective outside (two graphs around an eager all-reduce)."""

    def __init__(self, renderer, opt, example_batch, tr, world_size=1, warmup=3, inject=None):
        import os
        self.renderer, self.opt, self.tr, self.world = renderer, opt, tr, world_size
        self.batch = {k: v.clone() for k, v in example_batch.items()}
        # `inject`: static device tensors for the RNG draws of the step (jitter [N], perturb_noise [M,3]; parity tests refill them
        # in place between replays); None = the step draws them itself (graph-safe Philox offsets)
        self.inject = inject or {}
        dev = self.batch['rays_o'].device
        if world_size > 1:
            renderer.sdf_count_override = torch.ones((), device=dev)
        self.nccl_in_graph = world_size > 1 and os.environ.get('MORPHEUS_B200_GRAPH_NCCL', '1') != '0'
        self.captures = 0
        self._prepare()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                train_step_compute(renderer, opt, self.batch, tr, world_size, **self.inject)
                opt.all_reduce()
                opt.step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._capture()

    def _capture(self):
        renderer, opt, tr, world = self.renderer, self.opt, self.tr, self.world
        self.levels = renderer.model._levels()
        self.graph_b = None
        if world == 1 or self.nccl_in_graph:
            try:
                self.graph_a = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_a):
                    self.loss = train_step_compute(renderer, opt, self.batch, tr, world, **self.inject)
                    opt.all_reduce()
                    opt.step()
                self.captures += 1
                return
            except Exception:
                if world == 1:
                    raise
                self.nccl_in_graph = False       # this NCCL / torch build cannot capture the collective: split the step
                torch.cuda.synchronize()
                opt._clean = False
        self.graph_a = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_a):
            self.loss = train_step_compute(renderer, opt, self.batch, tr, world, **self.inject)
        self.graph_b = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_b):
            opt.step()
        self.captures += 1

    def _prepare(self):
        """global count of samples with a depth observation (utils.py:107 normaliser) for the fixed-S sampler:
        (#rays with depth > 0) * S, averaged over ranks (see render.global_count)"""
        if self.world > 1:
            if 'n_depth' in self.batch:
                # the data pipeline that shards the ray batch knows the GLOBAL number of samples with a depth observation: it ships
                # it with the shard (4 bytes), so the step needs no extra collective and no eager launches before the graph
                self.renderer.sdf_count_override.copy_(self.batch['n_depth'].reshape(()))
                return
            from .render import global_count
            S = self.renderer.uniform_samples
            cnt = global_count(torch.count_nonzero(self.batch['depth']) * S, self.world)
            self.renderer.sdf_count_override.copy_(cnt)

    def load(self, batch, non_blocking=True):
        for k, v in batch.items():
            self.batch[k].copy_(v, non_blocking=non_blocking)

    def step(self, batch=None):
        if batch is not None:
            self.load(batch)
        if self.renderer.model._levels() != self.levels:
            self._capture()       # coarse-to-fine level changed: the captured kernels carry the old n_levels / n_freq by value
        self._prepare()
        self.graph_a.replay()
        if self.graph_b is not None:
            self.opt.all_reduce()
            self.graph_b.replay()
        return self.loss

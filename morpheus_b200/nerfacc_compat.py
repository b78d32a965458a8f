"""nerfacc look-alikes for the three calls MorpheuS makes (morpheus.py:200-202, :629-638,
:675-685, :913): `OccGridEstimator` (`sampling`, `update_every_n_steps`, state_dict),
`render_weight_from_density`, `accumulate_along_rays` -- same argument names and returns as the
nerfacc 0.5.x public API, compute in csrc/sampler.cu and csrc/composite.cu.

nerfacc itself is an un-vendored, unpinned dependency of the reference (docs/INSTALL.md:22-23), so
the *sampler's* tie-breaking is parity-unpinned; the compositing math is closed form.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, ptr, stream


def ray_segments(ray_indices, n_rays):
    """[n_rays+1] int32 first-sample offsets of a sorted packed ray_indices."""
    return torch.searchsorted(ray_indices.contiguous(), torch.arange(n_rays + 1, device=ray_indices.device, dtype=ray_indices.dtype)).to(torch.int32)


class _Composite(torch.autograd.Function):
    """(sigmas [M], rgbs [M,3] | None) -> weights [M], opacity [N], depth [N], rgb [N,3] in one launch."""

    @staticmethod
    def forward(ctx, sigmas, rgbs, t_starts, t_ends, seg, n_rays):
        M = sigmas.shape[0]
        dev = sigmas.device
        sigmas, t_starts, t_ends = sigmas.contiguous().float(), t_starts.contiguous().float(), t_ends.contiguous().float()
        rgbs = rgbs.contiguous().float() if rgbs is not None else None
        weights = torch.empty(M, device=dev)
        trans = torch.empty(M, device=dev)
        alphas = torch.empty(M, device=dev)
        opacity = torch.empty(n_rays, device=dev)
        depth = torch.empty(n_rays, device=dev)
        rgb = torch.empty(n_rays, 3, device=dev) if rgbs is not None else None
        with _lib.timed('composite_fwd'):
          check(_lib.lib().mb_composite_forward(ptr(seg), n_rays, M, ptr(sigmas), ptr(t_starts), ptr(t_ends), ptr(rgbs), ptr(weights),
                                              ptr(trans), ptr(alphas), ptr(opacity), ptr(depth), ptr(rgb), stream()), 'composite_forward')
        ctx.save_for_backward(sigmas, rgbs, t_starts, t_ends, seg)
        ctx.n_rays = n_rays
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(trans, alphas)
        return weights, opacity, depth, rgb, trans, alphas

    @staticmethod
    def backward(ctx, g_w, g_o, g_d, g_rgb, _gt, _ga):
        sigmas, rgbs, t0, t1, seg = ctx.saved_tensors
        M = sigmas.shape[0]

        def cg(g):
            return g.contiguous().float() if g is not None else None
        g_w, g_o, g_d, g_rgb = map(cg, (g_w, g_o, g_d, g_rgb))
        g_sig = torch.empty_like(sigmas)
        g_rgbs = torch.empty_like(rgbs) if rgbs is not None else None
        with _lib.timed('composite_bwd'):
          check(_lib.lib().mb_composite_backward(ptr(seg), ctx.n_rays, M, ptr(sigmas), ptr(t0), ptr(t1), ptr(rgbs), ptr(g_w), ptr(g_o),
                                               ptr(g_d), ptr(g_rgb), ptr(g_sig), ptr(g_rgbs), stream()), 'composite_backward')
        return g_sig, g_rgbs, None, None, None, None


def composite(sigmas, rgbs, t_starts, t_ends, ray_indices, n_rays, seg=None):
    """fused weights + opacity + depth + rgb (what morpheus.py:675-685 does in four nerfacc calls)"""
    if seg is None:
        seg = ray_segments(ray_indices, n_rays)
    w, o, d, rgb, _, _ = _Composite.apply(sigmas, rgbs, t_starts, t_ends, seg, n_rays)
    return w, o, d, rgb


def render_weight_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None, n_rays=None, prefix_trans=None):
    """nerfacc.render_weight_from_density -> (weights, trans, alphas)"""
    if prefix_trans is not None:
        raise NotImplementedError('prefix_trans is never used by MorpheuS')
    if ray_indices is None:
        raise NotImplementedError('packed samples are addressed by ray_indices (morpheus.py:679)')
    seg = ray_segments(ray_indices, n_rays)
    w, _, _, _, trans, alphas = _Composite.apply(sigmas, None, t_starts, t_ends, seg, n_rays)
    return w, trans, alphas


def accumulate_along_rays(weights, values=None, ray_indices=None, n_rays=None):
    """nerfacc.accumulate_along_rays -> [n_rays, D].  Kept for API parity with the reference's three separate
    calls; the product path (morpheus_b200.render) gets all three sums from the single fused `composite` launch."""
    src = weights[..., None] if values is None else weights[..., None] * values
    out = torch.zeros((n_rays, src.shape[-1]), device=src.device, dtype=src.dtype)
    return out.index_add_(0, ray_indices, src)


class OccGridEstimator(nn.Module):
    """nerfacc.OccGridEstimator(roi_aabb, resolution, levels=1) as used at morpheus.py:200-202."""

    DIM = 3

    def __init__(self, roi_aabb, resolution=128, levels=1):
        super().__init__()
        if levels != 1:
            raise NotImplementedError('MorpheuS uses a single-level grid (morpheus.py:200)')
        aabb = torch.as_tensor(roi_aabb, dtype=torch.float32).reshape(1, 6)
        self.resolution = int(resolution)
        self.levels = 1
        self.cells_per_lvl = self.resolution ** 3
        self.register_buffer('aabbs', aabb)
        self.register_buffer('occs', torch.zeros(self.cells_per_lvl))
        self.register_buffer('binaries', torch.zeros((1, self.resolution, self.resolution, self.resolution), dtype=torch.bool))
        r = torch.arange(self.resolution)
        coords = torch.stack(torch.meshgrid(r, r, r, indexing='ij'), dim=-1).reshape(-1, 3)
        self.register_buffer('grid_coords', coords, persistent=False)
        self.register_buffer('grid_indices', torch.arange(self.cells_per_lvl), persistent=False)

    def _aabb_host(self):
        return (C.c_float * 6)(*[float(v) for v in self.aabbs[0].tolist()])

    @torch.no_grad()
    def sampling(self, rays_o, rays_d, sigma_fn=None, alpha_fn=None, near_plane=0.0, far_plane=1e10, t_min=None, t_max=None,
                 render_step_size=1e-3, early_stop_eps=1e-4, alpha_thre=0.0, stratified=False, cone_angle=0.0, jitter=None):
        """-> (ray_indices [M] int64, t_starts [M], t_ends [M]) packed and sorted by ray.  `jitter` ([N] in [0,1))
        lets tests inject the stratified draw; by default it is torch.rand (morpheus.py:635 stratified=True)."""
        if sigma_fn is not None or alpha_fn is not None or cone_angle != 0.0 or t_min is not None or t_max is not None:
            raise NotImplementedError('MorpheuS samples with sigma_fn=None, alpha_thre=0, cone_angle=0 (morpheus.py:629-638)')
        rays_o, rays_d = rays_o.contiguous().float(), rays_d.contiguous().float()
        N = rays_o.shape[0]
        dev = rays_o.device
        if stratified and jitter is None:
            jitter = torch.rand(N, device=dev)
        jitter = jitter.contiguous().float() if jitter is not None else None
        binaries = self.binaries.view(torch.uint8).contiguous()
        aabb = self._aabb_host()
        counts = torch.empty(N, dtype=torch.int32, device=dev)
        L = _lib.lib()
        check(L.mb_sample_rays_count(ptr(rays_o), ptr(rays_d), N, ptr(binaries), self.resolution, aabb, C.c_float(render_step_size),
                                     C.c_float(near_plane), C.c_float(far_plane), ptr(jitter), ptr(counts), stream()), 'sample_rays_count')
        csum = torch.cumsum(counts, 0, dtype=torch.int32)
        M = int(csum[-1].item()) if N > 0 else 0   # the packed length must reach the host (nerfacc does the same)
        offsets = (csum - counts).contiguous()
        ray_indices = torch.empty(M, dtype=torch.int64, device=dev)
        t_starts = torch.empty(M, device=dev)
        t_ends = torch.empty(M, device=dev)
        if M > 0:
            check(L.mb_sample_rays_write(ptr(rays_o), ptr(rays_d), N, ptr(binaries), self.resolution, aabb, C.c_float(render_step_size),
                                         C.c_float(near_plane), C.c_float(far_plane), ptr(jitter), ptr(offsets), ptr(ray_indices),
                                         ptr(t_starts), ptr(t_ends), stream()), 'sample_rays_write')
        return ray_indices, t_starts, t_ends

    @torch.no_grad()
    def update_every_n_steps(self, step, occ_eval_fn, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16):
        if not self.training:
            raise RuntimeError('update_every_n_steps() is only for training (call .train())')
        if step % n == 0 and self.training:
            self._update(step, occ_eval_fn, occ_thre, ema_decay, warmup_steps)

    @torch.no_grad()
    def _update(self, step, occ_eval_fn, occ_thre, ema_decay, warmup_steps, cell_idx=None, cell_jitter=None):
        """nerfacc 0.5.x OccGridEstimator._update: cells = all (step < warmup_steps) or N/4 uniform + <= N/4 occupied; x = (ijk + rand) /
        resolution mapped into the AABB; occs[c] = max(ema_decay * occs[c], occ(x)); binaries = occs > min(mean(occs[occs >= 0]), occ_thre).
        `cell_idx` / `cell_jitter` inject the two random draws (parity tests).  During warm-up (every cell, the 2 097 152-point refresh)
        nothing here synchronises with the host: the threshold stays on the device."""
        dev = self.occs.device
        if cell_idx is not None:
            idx = cell_idx.to(dev).long()
        elif step < warmup_steps:
            idx = self.grid_indices
        else:
            n = self.cells_per_lvl // 4
            uni = torch.randint(self.cells_per_lvl, (n,), device=dev)
            occ_idx = torch.nonzero(self.binaries.flatten())[:, 0]
            if occ_idx.shape[0] > n:
                occ_idx = occ_idx[torch.randint(occ_idx.shape[0], (n,), device=dev)]
            idx = torch.cat([uni, occ_idx])
        coords = self.grid_coords[idx]
        if cell_jitter is None:
            cell_jitter = torch.rand_like(coords, dtype=torch.float32)
        x = (coords + cell_jitter.to(dev)) / self.resolution
        lo, hi = self.aabbs[0, :3], self.aabbs[0, 3:]
        x = lo + x * (hi - lo)
        occ = occ_eval_fn(x).reshape(-1).contiguous().float()   # the reference passes sigma * step_size (morpheus.py:911)
        idx = idx.contiguous()
        check(_lib.lib().mb_occ_update(ptr(self.occs), ptr(idx), ptr(occ), idx.shape[0], C.c_float(ema_decay), C.c_float(1.0), stream()),
              'occ_update')
        valid = self.occs >= 0
        thre = torch.clamp((self.occs * valid).sum() / valid.sum().clamp(min=1), max=occ_thre).reshape(1).float().contiguous()
        bin8 = torch.empty(self.cells_per_lvl, dtype=torch.uint8, device=dev)
        check(_lib.lib().mb_occ_binarize_dev(ptr(self.occs), self.cells_per_lvl, ptr(thre), ptr(bin8), stream()), 'occ_binarize_dev')
        self.binaries = bin8.view(torch.bool).view(self.binaries.shape)

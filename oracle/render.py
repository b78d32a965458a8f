"""Oracle: ray generation, sampling, compositing, losses, render_rays (TEST INFRASTRUCTURE, CPU).

nerfacc is NOT in /root/reference (un-vendored, unpinned pip dependency, docs/INSTALL.md:22-23).
**Parity unpinned at that boundary.**  The functions below restate the published nerfacc 0.5.x
semantics that the reference call sites rely on (morpheus.py:200-202, :629-638, :675-685, :913);
the compositing math is closed form, the sampler is checked through invariants.
"""
import math

import numpy as np
import torch

from .fields import safe_normalize


# ------------------------------------------------------------------------------------------------
# rays  (datasets/utils.py:28-65, datasets/dataset.py:336-433)
# ------------------------------------------------------------------------------------------------
def camera_dirs(H, W, fx, fy, cx, cy):
    """OpenGL pinhole directions ((i+.5-cx)/fx, -(j+.5-cy)/fy, -1), un-normalised, [H,W,3]."""
    i, j = torch.meshgrid(torch.arange(W, dtype=torch.float32), torch.arange(H, dtype=torch.float32), indexing='xy')
    return torch.stack([(i + 0.5 - cx) / fx, -(j + 0.5 - cy) / fy, -torch.ones_like(i)], -1)


def rays_from_pose(dirs_cam, c2w):
    """d_w = R d_c ; o = c2w[:3,3]  (datasets/dataset.py:363-396)"""
    d = (dirs_cam[..., None, :] * c2w[:3, :3]).sum(-1)
    o = c2w[:3, 3].expand_as(d)
    return o, d


def look_at_pose(theta_deg, phi_deg, radius):
    """Camera on a sphere looking at the origin, OpenGL convention (-z forward, +y up).
    Synthetic 'snoopy-shaped' camera used by bench/tests."""
    th, ph = math.radians(theta_deg), math.radians(phi_deg)
    c = torch.tensor([radius * math.sin(th) * math.sin(ph), radius * math.cos(th), radius * math.sin(th) * math.cos(ph)])
    fwd = -c / c.norm()
    up = torch.tensor([0.0, 1.0, 0.0])
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm()
    upv = torch.linalg.cross(right, fwd)
    c2w = torch.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, upv, -fwd, c
    return c2w


# ------------------------------------------------------------------------------------------------
# sampling
# ------------------------------------------------------------------------------------------------
def ray_aabb(o, d, aabb):
    """slab test; returns tmin,tmax (tmax<tmin => miss); near plane 0."""
    inv = 1.0 / d
    t0 = (aabb[:3] - o) * inv
    t1 = (aabb[3:] - o) * inv
    tmin = torch.minimum(t0, t1).amax(-1).clamp(min=0.0)
    tmax = torch.maximum(t0, t1).amin(-1)
    return tmin, tmax


def sample_uniform(o, d, aabb, n_samples, jitter):
    """Fixed-S synthetic sampler of BASELINE cfg-1/2/4: S equal intervals spanning the
    AABB chord of each ray, whole lattice shifted by jitter[r]*delta_r (stratified)."""
    tmin, tmax = ray_aabb(o, d, aabb)
    hit = tmax > tmin
    tmax = torch.where(hit, tmax, tmin + 1.0)
    delta = (tmax - tmin) / (n_samples + 1)
    k = torch.arange(n_samples, dtype=o.dtype)
    t0 = tmin[:, None] + (k[None, :] + jitter[:, None]) * delta[:, None]
    t1 = t0 + delta[:, None]
    N = o.shape[0]
    ray_indices = torch.arange(N).repeat_interleave(n_samples)
    return ray_indices, t0.reshape(-1), t1.reshape(-1)


def sample_occgrid(o, d, binaries, aabb, step, jitter, near=0.0, far=1e10):
    """OccGridEstimator.sampling(sigma_fn=None, alpha_thre=0, stratified=True, cone_angle=0)
    as called at morpheus.py:629-638 -- restated (nerfacc source absent, see module docstring):
    per ray, clip to the AABB, start a lattice of step `step` (in units of |d|, d un-normalised)
    at t = max(tmin,near) + jitter*step, keep every interval [t,t+step) whose midpoint lies in an
    occupied cell.  Python loop: small cases only."""
    res = binaries.shape[-1]
    ri, ts, te = [], [], []
    tmin, tmax = ray_aabb(o, d, aabb)
    lo, ext = aabb[:3], aabb[3:] - aabb[:3]
    for r in range(o.shape[0]):
        a, b = max(float(tmin[r]), near), min(float(tmax[r]), far)
        if not b > a:
            continue
        t = np.float32(a) + np.float32(jitter[r]) * np.float32(step)
        while np.float32(t + np.float32(0.5) * np.float32(step)) < np.float32(b):
            mid = np.float32(t + np.float32(0.5) * np.float32(step))
            p = o[r] + d[r] * float(mid)
            c = torch.floor((p - lo) / ext * res).long().clamp(0, res - 1)
            if bool(binaries[c[0], c[1], c[2]]):
                ri.append(r), ts.append(float(t)), te.append(float(np.float32(t + np.float32(step))))
            t = np.float32(t + np.float32(step))
    return (torch.tensor(ri, dtype=torch.long), torch.tensor(ts, dtype=torch.float32), torch.tensor(te, dtype=torch.float32))


# ------------------------------------------------------------------------------------------------
# compositing  (nerfacc.render_weight_from_density / accumulate_along_rays; morpheus.py:675-685)
# ------------------------------------------------------------------------------------------------
def render_weight_from_density(t_starts, t_ends, sigmas, ray_indices, n_rays):
    """alpha_i = 1-exp(-sigma_i*dt_i); T_i = exp(-sum_{j<i in ray} sigma_j dt_j); w = T*alpha.
    Differentiable torch (segmented exclusive cumsum built from a global cumsum)."""
    sd = sigmas * (t_ends - t_starts)
    alphas = 1.0 - torch.exp(-sd)
    cs = torch.cumsum(sd, 0)
    excl = cs - sd
    # subtract the running total at each ray's first sample
    first = torch.ones_like(ray_indices, dtype=torch.bool)
    first[1:] = ray_indices[1:] != ray_indices[:-1]
    start_val = torch.zeros(n_rays, dtype=sd.dtype)
    start_val[ray_indices[first]] = excl[first]
    trans = torch.exp(-(excl - start_val[ray_indices]))
    return trans * alphas, trans, alphas


def accumulate_along_rays(weights, values, ray_indices, n_rays):
    src = weights[:, None] if values is None else weights[:, None] * values
    out = torch.zeros(n_rays, src.shape[-1], dtype=src.dtype)
    return out.index_add(0, ray_indices, src)


# ------------------------------------------------------------------------------------------------
# losses on the path  (utils.py:91-113)
# ------------------------------------------------------------------------------------------------
def get_sdf_loss(z_vals, target_d, predicted_sdf, truncation, mask=None):
    """utils.py:91-113 including its per-sample 'sum(dim=-1)' quirk on [M,1] tensors."""
    s = predicted_sdf[..., None]
    depth_mask = target_d > 0.0
    front = (z_vals < (target_d - truncation)) | ((target_d < 0.0) & (z_vals < 3.5))
    bound = torch.where(target_d < 0.0, torch.full_like(z_vals, 10.0), target_d - z_vals)
    sdf_mask = (bound.abs() <= truncation) & depth_mask
    if mask is not None:
        sdf_mask = sdf_mask & (mask > 0.5)
    n = front.sum(-1) + sdf_mask.sum(-1) + 1e-8
    rays_w_depth = torch.count_nonzero(target_d)
    fs = torch.max(torch.exp(-5.0 * s) - 1.0, s - bound).clamp(min=0.0) * front
    fs_loss = (fs.sum(-1) / n).sum() / rays_w_depth
    sdf_loss = ((torch.abs(s - bound) * sdf_mask).sum(-1) / n).sum() / rays_w_depth
    return fs_loss, sdf_loss


# ------------------------------------------------------------------------------------------------
# render_rays  (morpheus.py:558-794), deterministic: every RNG draw is an explicit argument
# ------------------------------------------------------------------------------------------------
def render_rays(scene, rays_o, rays_d, rays_t, rays_id, samples, bg_color=None, ambient_ratio=1.0,
                light_d=None, shading='albedo', optimize_pose=False, rays_depth=None, rays_mask=None,
                perturb_noise=None, trunc=0.1, smoothness_std=0.005, training=False, real_view=True):
    """samples = (ray_indices, t_starts, t_ends) from a sampler above.  Returns the `results` dict of
    morpheus.py:699-706 plus the per-sample aux losses that live inside render_rays
    (loss_orient :709-712, loss_normal_perturb :714-741, sdf_loss/fs_loss :787-790)."""
    o, d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
    t, rid = rays_t.reshape(-1, 1), rays_id.reshape(-1, 1)
    if optimize_pose:
        o, d = scene.pose_optimisation(o, d, rid)
    N = o.shape[0]
    ray_indices, t0, t1 = samples
    tm = ((t0 + t1) / 2.0)[:, None]
    xyz = o[ray_indices] + d[ray_indices] * tm
    res = {}
    if xyz.shape[0] == 0:  # morpheus.py:663-670
        res.update(image=torch.ones(N, 3), depth=torch.zeros(N), sdf=None, weights=None, weights_sum=None,
                   normal=None, deform=None, normal_raw=None)
        return res
    light = light_d[ray_indices] if light_d is not None else None
    sdf, sigma, rgb, normal, deform, raw = scene.forward(xyz, t[ray_indices], light, ratio=ambient_ratio, shading=shading)
    w, _, _ = render_weight_from_density(t0, t1, sigma, ray_indices, N)
    opacity = accumulate_along_rays(w, None, ray_indices, N)
    depth = accumulate_along_rays(w, tm, ray_indices, N)
    col = accumulate_along_rays(w, rgb, ray_indices, N)
    bg = 1 if bg_color is None else bg_color
    res.update(image=col + (1 - opacity) * bg, depth=depth[:, 0], sdf=sdf, weights=w, weights_sum=opacity,
               normal=normal, deform=deform, normal_raw=raw)
    if training:
        if normal is not None and not real_view:
            tdir = safe_normalize(d[ray_indices])
            res['loss_orient'] = (w.detach() * (normal * tdir).sum(-1).clamp(min=0) ** 2).sum(-1).mean()
        if normal is not None and perturb_noise is not None:
            n2, _ = scene.normal(xyz + perturb_noise * smoothness_std, topo=None)
            res['loss_normal_perturb'] = (normal - n2).abs().mean()
        if rays_depth is not None:
            fs, sl = get_sdf_loss(tm, rays_depth.reshape(-1, 1)[ray_indices], sdf, trunc,
                                  mask=rays_mask.reshape(-1, 1)[ray_indices] if rays_mask is not None else None)
            res['fs_loss'], res['sdf_loss'] = fs, sl
    return res


# ------------------------------------------------------------------------------------------------
# observation-space normal smoothness  (morpheus.py:518-556), RNG draws explicit
# ------------------------------------------------------------------------------------------------
def ortho_normal_dir(normals, phi):
    """morpheus.py:518-528 with the random angle `phi` [.., 1] in [0, 2 pi) passed in"""
    n = torch.nn.functional.normalize(normals, dim=-1)
    u = torch.nn.functional.normalize(n[..., [1, 0, 2]] * torch.tensor([1., -1., 0.]), dim=-1)
    v = torch.cross(n, u, dim=-1)
    return torch.cos(phi) * u + torch.sin(phi) * v


def normal_smoothness_loss(scene, rays_o, rays_d, rays_t, depth, trunc_noise, phi, trunc=0.1, smoothness_std=0.005):
    """morpheus.py:530-556: 11 points per ray in a band around `depth`, normal(x, t) there and at a point displaced along a random
    tangent, mean squared difference over the points inside the 1.1 sphere (boolean indexing, as the reference)."""
    n_pts = int(trunc * 100 + 1)
    tn = torch.linspace(-0.5 * trunc, 0.5 * trunc, n_pts) + 0.01 * trunc_noise
    pts = ((depth.reshape(1, -1) + tn[:, None])[..., None] * rays_d[None, ...] + rays_o[None, ...]).view(-1, 3)
    ts = rays_t[None, ...].repeat(n_pts, 1, 1).view(-1, 1)
    keep = torch.linalg.norm(pts, ord=2, dim=-1) < 1.1
    n1, _ = scene.normal(pts[keep], t=ts[keep])
    w = ortho_normal_dir(n1, phi[keep] if phi.shape[0] == keep.shape[0] else phi)
    n2, _ = scene.normal(pts[keep] + w * smoothness_std, t=ts[keep])
    return torch.mean(torch.square(n1 - n2))


# ------------------------------------------------------------------------------------------------
# occupancy refresh  (nerfacc OccGridEstimator.update_every_n_steps -> _update; call site morpheus.py:905-913)
# ------------------------------------------------------------------------------------------------
def occ_grid_update(occs, cell_idx, jitter, aabb, resolution, occ_eval_fn, ema_decay=0.95, occ_thre=0.01):
    """Published nerfacc 0.5.x semantics (source absent: parity unpinned against nerfacc itself, see the module docstring): the cells
    `cell_idx` (all of them while step < 256, else N/4 uniform + <= N/4 occupied) are probed at one jittered point each,
    x = aabb_lo + (ijk + jitter) / resolution * extent with ijk the x-major coordinates of the cell (idx = (ix * res + iy) * res + iz),
    occs[c] = max(ema_decay * occs[c], occ(x)), binaries = occs > min(mean(occs[occs >= 0]), occ_thre).  -> (occs, binaries)"""
    ix = cell_idx // (resolution * resolution)
    iy = (cell_idx // resolution) % resolution
    iz = cell_idx % resolution
    coords = torch.stack([ix, iy, iz], dim=-1).to(jitter.dtype)
    x = (coords + jitter) / resolution
    x = aabb[:3] + x * (aabb[3:] - aabb[:3])
    occ = occ_eval_fn(x).reshape(-1)
    occs = occs.clone()
    occs[cell_idx] = torch.maximum(occs[cell_idx] * ema_decay, occ)
    thre = torch.clamp(occs[occs >= 0].mean(), max=occ_thre)
    return occs, occs > thre

"""Ray generation for real (RGB-D) and virtual (novel) views, plus the synthetic 'snoopy-shaped' data the
bench and the tests use (no dataset can be downloaded here).

Mirrors datasets/utils.py:28-65 (get_camera_rays), datasets/dataset.py:225-266 (look-at pose, OpenGL,
keep_chirality), :363-396 (world-space rays of a frame), :398-433 (sample_real_view_rays: one random
frame, one shared set of random pixels) and :503-578 (get_virtual_view_rays: full low-res image from a
random camera on the sphere + delta polar / azimuth / radius w.r.t. the frame's real camera).
Everything is torch and device-agnostic: generated on the GPU it removes the per-step CPU gather + H2D
copy the reference pays (morpheus.py:841-850).
"""
import math

import torch


def safe_normalize(x, eps=1e-20):
    return x / torch.sqrt(torch.clamp(torch.sum(x * x, -1, keepdim=True), min=eps))


def get_camera_rays(H, W, fx, fy=None, cx=None, cy=None, device='cpu'):
    i, j = torch.meshgrid(torch.arange(W, dtype=torch.float32, device=device), torch.arange(H, dtype=torch.float32, device=device), indexing='xy')
    if cx is None:
        cx, cy = 0.5 * W, 0.5 * H
    if fy is None:
        fy = fx
    return torch.stack([(i + 0.5 - cx) / fx, -(j + 0.5 - cy) / fy, -torch.ones_like(i)], -1)


def c2w_from_cam_center(cam_centers, targets=0.0):
    """datasets/dataset.py:225-266 for camera_convention='OpenGL', x_axis=None, keep_chirality=True."""
    fwd = safe_normalize(cam_centers - targets)
    up = torch.tensor([0.0, 1.0, 0.0], device=cam_centers.device).expand_as(fwd)
    right = safe_normalize(torch.linalg.cross(up, fwd, dim=-1))
    up = safe_normalize(torch.linalg.cross(fwd, right, dim=-1))
    poses = torch.eye(4, device=cam_centers.device).unsqueeze(0).repeat(fwd.shape[0], 1, 1)
    poses[:, :3, :3] = torch.stack((right, up, fwd), dim=-1)
    poses[:, :3, 3] = cam_centers
    return poses


def cam_center_from_polar(theta_deg, phi_deg, radius):
    """camera centre on the sphere: y up, polar angle theta from +y, azimuth phi about y"""
    th, ph = torch.deg2rad(theta_deg), torch.deg2rad(phi_deg)
    return torch.stack([radius * torch.sin(th) * torch.sin(ph), radius * torch.cos(th), radius * torch.sin(th) * torch.cos(ph)], -1)


def world_rays(dirs_cam, c2w):
    """d_w = sum_j d_j R[:, j];  o = c2w[:3, 3]   (datasets/dataset.py:363-396)"""
    d = torch.sum(dirs_cam[..., None, :] * c2w[:3, :3], -1)
    return c2w[:3, 3].expand_as(d).contiguous(), d.contiguous()


def virtual_view_rays(frame, num_frames, H, W, focal, scale, theta_range=(45.0, 105.0), phi_range=(-180.0, 180.0), radius=2.5,
                      ref_theta=90.0, ref_phi=0.0, ref_radius=2.5, generator=None, device='cpu', theta_deg=None, phi_deg=None):
    """get_virtual_view_rays (datasets/dataset.py:503-578) for one random training view; the shipped configs use
    uniform_sphere_rate 0 (configs/snoopy.yaml:18), i.e. polar / azimuth uniform in their ranges (get_virtual_view_data :467-470).
    `theta_deg` / `phi_deg` inject the two draws (degrees; the reference wraps negative azimuths by +360 before use, :469)."""
    u = torch.rand(2, generator=generator)
    theta = torch.tensor([theta_range[0] + float(u[0]) * (theta_range[1] - theta_range[0]) if theta_deg is None else float(theta_deg)])
    phi = torch.tensor([phi_range[0] + float(u[1]) * (phi_range[1] - phi_range[0]) if phi_deg is None else float(phi_deg)])
    phi[phi < 0] += 360.0
    pose = c2w_from_cam_center(cam_center_from_polar(theta, phi, torch.tensor([radius])))[0].to(device)
    h, w = int(scale * H), int(scale * W)
    dirs = get_camera_rays(h, w, focal * scale, focal * scale, 0.5 * W * scale, 0.5 * H * scale, device=device).reshape(-1, 3)
    o, d = world_rays(dirs, pose)
    dphi = phi - ref_phi
    dphi[dphi > 180] -= 360
    return {'H': h, 'W': w, 'rays_o': o[None], 'rays_d': d[None], 'rays_t': torch.full((1, h * w, 1), frame / num_frames, device=device),
            'rays_id': torch.full((1, h * w, 1), frame, dtype=torch.long, device=device), 'polar': theta - ref_theta, 'azimuth': dphi,
            'radius': torch.tensor([radius - ref_radius])}


def synthetic_real_view_batch(n_rays, seed, frame=41, num_frames=200, H=360, W=360, focal=517.0, radius=2.5, device='cpu'):
    """sample_real_view_rays(ray_num=n_rays) on a synthetic snoopy-shaped scene: H=W=360, f=517, camera on the r=2.5
    sphere looking at the origin, theta in [45,105] deg; RGB-D 'observations' come from an analytic sphere of radius
    0.4 (depth = z-depth because directions have camera z = -1).  Returns CPU (or `device`)
    tensors: rays_o/rays_d [N,3], rays_t [N,1], rays_id [N,1] int64, rgb [N,3], depth [N], mask [N], bg [N,3]."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(2, generator=g)
    centre = cam_center_from_polar(torch.tensor([45.0 + 60.0 * float(u[0])]), torch.tensor([-180.0 + 360.0 * float(u[1])]), torch.tensor([radius]))
    c2w = c2w_from_cam_center(centre)[0]
    dirs = get_camera_rays(H, W, focal, focal, W / 2, H / 2).reshape(-1, 3)
    idx = torch.randint(0, dirs.shape[0], (n_rays,), generator=g)
    o, d = world_rays(dirs[idx], c2w)
    a = (d * d).sum(-1)
    b = 2 * (o * d).sum(-1)
    c = (o * o).sum(-1) - 0.16
    disc = b * b - 4 * a * c
    hit = disc > 0
    tt = (-b - torch.sqrt(disc.clamp(min=0))) / (2 * a)
    depth = torch.where(hit, tt, torch.zeros_like(tt))
    mask = hit.float()
    p = o + d * depth[:, None]
    bg = torch.rand(n_rays, 3, generator=g)
    rgb = (0.5 + 0.5 * torch.sin(p * 7.0)) * mask[:, None] + bg * (1 - mask[:, None])
    batch = {'rays_o': o, 'rays_d': d, 'rays_t': torch.full((n_rays, 1), frame / num_frames),
             'rays_id': torch.full((n_rays, 1), frame, dtype=torch.long), 'rgb': rgb, 'depth': depth, 'mask': mask, 'bg': bg}
    return {k: v.to(device) for k, v in batch.items()}


class RealViewData:
    """Device-resident RGB-D sequence with the sampling interface of the reference dataset:
    `get_real_view_rays` (datasets/dataset.py:336-396) + `sample_real_view_rays` (:398-433), known_view_scale 1.
    The reference materialises the full [F, H*W, 3] origin / direction tables on the CPU once, gathers on the CPU every step and
    copies the batch to the GPU (morpheus.py:841-850); here images / depths / masks / poses live on `device` and the rays of the
    selected pixels are computed on the fly (2 small launches), so a step needs no host work and no H2D copy.

    images [F,H,W,3] float, depths [F,H,W] float, masks [F,H,W], poses [F,4,4] (c2w, OpenGL), K [3,3] or [4,4]."""

    def __init__(self, images, depths, masks, poses, K, device='cpu', theta=None, phi=None, radius=None):
        self.device = torch.device(device)
        self.num_frames, self.H, self.W = int(images.shape[0]), int(images.shape[1]), int(images.shape[2])
        f32 = dict(device=self.device, dtype=torch.float32)
        self.image = torch.as_tensor(images).to(**f32).permute(0, 3, 1, 2).contiguous()       # [F,3,H,W] as the reference keeps it
        self.depth = torch.as_tensor(depths).to(**f32)
        self.mask = torch.as_tensor(masks).to(device=self.device, dtype=torch.int64)
        self.pose = torch.as_tensor(poses).to(**f32)
        K = torch.as_tensor(K).to(**f32)
        self.intri = K
        self.dirs = get_camera_rays(self.H, self.W, K[0, 0], K[1, 1], K[0, 2], K[1, 2], device=self.device).reshape(-1, 3)
        self.theta, self.phi, self.radius = theta, phi, radius

    def sample_real_view_rays(self, idx=None, bs=1, ray_num=None, index=None, generator=None):
        """same keys / shapes as the reference: rays_o / rays_d [bs, n, 3], rays_t [bs, n, 1], rays_id [bs, n, 1] int64, image
        [bs, 3, n, 1], mask / depth [bs, n, 1], H = n, W = 1 (n = ray_num, or H*W when ray_num is None).  `idx` / `index` inject the
        frame and pixel draws (torch.randint in the reference, :400,413)."""
        if idx is None:
            idx = torch.randint(0, self.num_frames, (bs,), generator=generator)
        elif isinstance(idx, int):
            idx = torch.tensor([idx])
        idx = torch.as_tensor(idx).to(self.device).long()
        bs = idx.shape[0]
        if ray_num is not None and index is None:
            index = torch.randint(0, self.H * self.W, (ray_num,), generator=generator)
        full_frame = index is None
        if full_frame:
            index = torch.arange(self.H * self.W)
        index = torch.as_tensor(index).to(self.device).long()
        n = index.shape[0]
        pose = self.pose[idx]                                                    # [bs,4,4]
        d_cam = self.dirs[index]                                                 # [n,3]
        rays_d = torch.sum(d_cam[None, :, None, :] * pose[:, None, :3, :3], -1)   # d_w = sum_j d_j R[:, j]
        rays_o = pose[:, None, :3, 3].expand(bs, n, 3).contiguous()
        out = {'rays_o': rays_o, 'rays_d': rays_d,
               'rays_t': (idx.float() / self.num_frames)[:, None, None].expand(bs, n, 1).contiguous(),
               'rays_id': idx[:, None, None].expand(bs, n, 1).contiguous(),
               'image': self.image[idx].reshape(bs, 3, -1)[..., index].reshape(bs, 3, n, 1),
               'mask': self.mask[idx].reshape(bs, -1)[..., index].reshape(bs, n, 1),
               'depth': self.depth[idx].reshape(bs, -1)[..., index].reshape(bs, n, 1),
               'H': n if not full_frame else self.H, 'W': 1 if not full_frame else self.W, 'intri': self.intri, 'pose': pose}
        if full_frame:      # the reference leaves image / mask / depth of a whole frame in image layout (:402-411)
            out['image'], out['mask'], out['depth'] = self.image[idx], self.mask[idx], self.depth[idx]
        for k in ('theta', 'phi', 'radius'):
            v = getattr(self, k)
            if v is not None:
                out[k] = torch.as_tensor(v).to(self.device)[idx]
        return out

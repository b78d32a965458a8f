"""Oracle: the scalar SDS chain of Zero123.train_step (TEST INFRASTRUCTURE, closed form).
Restates models/guidance/zero123_utils.py:75-87 (schedule), :147-152 (angle-based grad scale), :180 (add_noise, diffusers
DDIMScheduler semantics), :197-206 (pose token T, classifier-free guidance), :210-212 (w(t), grad, nan_to_num), :233-234 (loss)."""
import math

import torch


def alphas_cumprod(n=1000, beta_start=0.00085, beta_end=0.012):
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, 0)


def add_noise(z, eps, t, ac):
    a = ac[t] ** 0.5
    b = (1 - ac[t]) ** 0.5
    return a.view(-1, 1, 1, 1) * z + b.view(-1, 1, 1, 1) * eps


def angle_between_deg(sph1, sph2):
    """python double loop of zero123_utils.py:102-120"""
    def cart(s):
        r, th, ph = float(s[0]), float(s[1]), float(s[2])
        return torch.tensor([r * math.sin(th) * math.cos(ph), r * math.sin(th) * math.sin(ph), r * math.cos(th)])
    out = torch.empty(len(sph1), len(sph2))
    for i, a in enumerate(sph1):
        for j, b in enumerate(sph2):
            u, v = cart(a), cart(b)
            out[i, j] = torch.arccos(torch.clip(torch.dot(u / u.norm(), v / v.norm()), -1.0, 1.0))
    return torch.rad2deg(out)


def pose_token(polar_deg, azimuth_deg, radius):
    a = azimuth_deg.clone()
    a[a > 180] -= 360
    return torch.stack([torch.deg2rad(polar_deg), torch.sin(torch.deg2rad(a)), torch.cos(torch.deg2rad(a)), radius], dim=-1)[:, None, :]


def sds_grad(eps_uncond, eps_cond, noise, t, ac, guidance_scale, grad_scale):
    pred = eps_uncond + guidance_scale * (eps_cond - eps_uncond)
    w = 1 - ac[t]
    return torch.nan_to_num((grad_scale * w).view(-1, 1, 1, 1) * (pred - noise))


def sds_loss(latents, grad):
    target = (latents - grad).detach()
    return 0.5 * torch.nn.functional.mse_loss(latents.float(), target, reduction='sum') / latents.shape[0]

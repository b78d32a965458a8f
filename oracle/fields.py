"""Oracle: the scene representation (TEST INFRASTRUCTURE, torch on CPU, fp32 or fp64).

A functional restatement of /root/reference/models/{model,decoders,deform_code,density,encodings,pose}.py
that consumes a *reference-format state_dict* (key names of SURVEY.md Appendix B), so the same
seeded weights can be pushed through the unmodified reference classes (tests/golden/make_scene_golden.py)
and through this file, and through the CUDA product path.

Everything is plain differentiable torch, so `torch.autograd` on this oracle yields the reference
gradients for the backward-parity tests.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .grid import _GridEncodeOracle, make_offsets, per_level_scale

EPS_FD = 2e-3  # models/model.py:367


def freq_encode(x, n_freqs=6, max_level=None):
    """models/encodings.py:35-57: [x, sin(2^k x), cos(2^k x)]_k, bands >= int(max_level*n) zeroed."""
    n_on = n_freqs if max_level is None else int(max_level * n_freqs)
    parts = [x]
    for k in range(n_on):
        f = float(2 ** k)
        parts += [torch.sin(x * f), torch.cos(x * f)]
    if n_freqs - n_on > 0:
        parts.append(torch.zeros(*x.shape[:-1], (n_freqs - n_on) * 2 * x.shape[-1], dtype=x.dtype))
    return torch.cat(parts, dim=-1)


def multicode_sample(volumes, t):
    """models/deform_code.py:20-40: align_corners=True 1-D linear interpolation at t*(S-1), t clamped to [0,1].
    volumes: list of [1,C,S,1]; t: [M,1] -> [M, C*len(volumes)]"""
    t = t.clamp(0, 1).reshape(-1)
    feats = []
    for v in volumes:
        line = v[0, :, :, 0]  # [C,S]
        S = line.shape[1]
        # grid_sample maps coord g in [-1,1] to (g+1)/2*(S-1) for align_corners=True; g = 2t-1
        pos = ((t * 2 - 1) + 1) / 2 * (S - 1)
        i0 = torch.floor(pos).long().clamp(0, S - 1)
        i1 = (i0 + 1).clamp(0, S - 1)
        w1 = pos - i0.to(pos.dtype)
        w0 = 1 - w1
        # out-of-range tap (i0+1 == S) contributes zero (padding_mode='zeros'); its weight is 0 at t==1
        valid1 = ((i0 + 1) <= S - 1).to(pos.dtype)
        feats.append((line[:, i0] * w0 + line[:, i1] * w1 * valid1).t())
    return torch.cat(feats, dim=-1)


def effective_weight(sd, prefix):
    """old-style nn.utils.weight_norm (dim=0): W = g * v / ||v||_row (models/decoders.py:51-52)."""
    if prefix + '.weight' in sd:
        return sd[prefix + '.weight']
    g, v = sd[prefix + '.weight_g'], sd[prefix + '.weight_v']
    return v * (g / v.norm(dim=1, keepdim=True))


def mlp(sd, name, n_layers, x):
    """models/decoders.py:59-64: Linear, ReLU between, none after the last."""
    for l in range(n_layers):
        W = effective_weight(sd, f'{name}.net.{l}')
        x = F.linear(x, W, sd[f'{name}.net.{l}.bias'])
        if l != n_layers - 1:
            x = torch.relu(x)
    return x


def laplace_sigma(sdf, beta_param):
    """models/density.py:22-31."""
    beta = beta_param.abs() + 1e-4
    return (1.0 / beta) * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() / beta))


def safe_normalize(x, eps=1e-20):
    """utils.py:70-71"""
    return x / torch.sqrt(torch.clamp((x * x).sum(-1, keepdim=True), min=eps))


def rotation_matrices(rot):
    """models/pose.py:35-58 (Euler angles -> R, stacked column-wise)."""
    ca, cb, cg = torch.cos(rot[:, 0]), torch.cos(rot[:, 1]), torch.cos(rot[:, 2])
    sa, sb, sg = torch.sin(rot[:, 0]), torch.sin(rot[:, 1]), torch.sin(rot[:, 2])
    c1 = torch.stack([ca * cb, sa * cb, -sb], -1)
    c2 = torch.stack([ca * sb * sg - sa * cg, sa * sb * sg + ca * cg, cb * sg], -1)
    c3 = torch.stack([ca * sb * cg + sa * sg, sa * sb * cg - ca * sg, cb * cg], -1)
    return torch.stack([c1, c2, c3], -1)


class SceneOracle:
    """Functional mirror of models/model.py:31 `scene_representation` for the shipped flags
    (use_t False, use_app False, use_joint True, color_grid True, encode_topo False;
    configs/snoopy.yaml:96-108).  `sd` is a dict of tensors with the reference key names."""

    N_DEFORM_LAYERS = 6
    N_SDF_LAYERS = 3
    N_COLOR_LAYERS = 3

    def __init__(self, sd, bound=1.01, num_frames=200, max_level=None):
        self.sd = sd
        self.bound = bound
        self.num_frames = num_frames
        self.max_level = max_level
        self.S = float(np.log2(per_level_scale(16, 128, 16)))  # grid.py:40
        self.H = 16

    # -- encoders ------------------------------------------------------------------------------
    def grid(self, which, x):
        """GridEncoder.forward, grid.py:152-169; which in {'encoder','encoder_c'}"""
        u = (x + self.bound) / (2 * self.bound)
        return _GridEncodeOracle.apply(u, self.sd[which + '.embeddings'], self.sd[which + '.offsets'],
                                       self.S, self.H, u.requires_grad, 0, False, 0, self.max_level)

    def code(self, t):
        return multicode_sample([self.sd[f'deform_code.volumes.{i}'] for i in range(3)], t)

    # -- model.py:412-437 ------------------------------------------------------------------------
    def warp(self, x, t):
        z = torch.cat([freq_encode(x, 6, self.max_level), self.code(t)], dim=-1)
        return mlp(self.sd, 'deform_net', self.N_DEFORM_LAYERS, z), mlp(self.sd, 'topo_net', self.N_DEFORM_LAYERS, z)

    # -- model.py:273-307 ------------------------------------------------------------------------
    def sigma_albedo(self, x, topo=None, return_color=True):
        enc = self.grid('encoder', x)
        if topo is None:
            topo = torch.zeros(x.shape[0], 2, dtype=x.dtype)
        h = mlp(self.sd, 'sdf_net', self.N_SDF_LAYERS, torch.cat([freq_encode(x, 6, self.max_level), enc, topo], dim=-1))
        sdf = h[..., 0]
        sigma = laplace_sigma(sdf, self.sd['sdf2density.beta'])
        albedo = None
        if return_color:
            enc_c = self.grid('encoder_c', x)
            albedo = torch.sigmoid(mlp(self.sd, 'color_net', self.N_COLOR_LAYERS, torch.cat([enc_c, h[..., 1:]], dim=-1)))
        return sdf, sigma, albedo

    # -- model.py:367-398 ------------------------------------------------------------------------
    def fd_normal_raw(self, x, topo=None):
        cols = []
        for ax in range(3):
            e = torch.zeros(1, 3, dtype=x.dtype)
            e[0, ax] = EPS_FD
            sp, _, _ = self.sigma_albedo((x + e).clamp(-self.bound, self.bound), topo, False)
            sn, _, _ = self.sigma_albedo((x - e).clamp(-self.bound, self.bound), topo, False)
            cols.append(0.5 * (sp - sn) / EPS_FD)
        return torch.stack(cols, dim=-1)

    def normal(self, x, t=None, cano=False, topo=None):
        if t is not None and not cano:
            deform, topo = self.warp(x, t)
            x = x + deform
        raw = self.fd_normal_raw(x, topo)
        return torch.nan_to_num(safe_normalize(raw)), raw

    # -- model.py:439-481 ------------------------------------------------------------------------
    def density(self, x, t=None, cano=False, allow_shape=False, return_color=True):
        topo = None
        if not (cano or t is None):
            if isinstance(t, float):
                t = t * torch.ones(x.shape[0], 1, dtype=x.dtype)
            if x.shape[0] != t.shape[0]:
                if not allow_shape:
                    raise Exception('Shape inconsistent!!!')
                t = t[0, 0] * torch.ones(x.shape[0], 1, dtype=x.dtype)
            deform, topo = self.warp(x, t)
            x = x + deform
        sdf, sigma, albedo = self.sigma_albedo(x, topo, return_color)
        return {'sdf': sdf, 'sigma': sigma, 'albedo': albedo}

    # -- model.py:483-533 ------------------------------------------------------------------------
    def forward(self, x, t, light_dir=None, ratio=1, shading='albedo', cano=False, return_color=True):
        if cano:
            xw, deform, topo = x, None, None
        else:
            deform, topo = self.warp(x, t)
            xw = x + deform
        sdf, sigma, albedo = self.sigma_albedo(xw, topo, return_color)
        if shading == 'albedo':
            return sdf, sigma, albedo, None, deform, None
        normal, raw = self.normal(x, topo=topo)  # at observation-space x (model.py:521)
        lambertian = ratio + (1 - ratio) * (normal * light_dir).sum(-1).clamp(min=0)
        if shading == 'textureless':
            color = lambertian.unsqueeze(-1).repeat(1, 3)
        elif shading == 'normal':
            color = (normal + 1) / 2
        else:
            color = albedo * lambertian.unsqueeze(-1)
        return sdf, sigma, color, normal, deform, raw

    # -- model.py:400-410 ------------------------------------------------------------------------
    def background(self, d, t):
        h = torch.cat([freq_encode(d, 6, None), freq_encode(t, 6, self.max_level)], dim=-1)
        return torch.sigmoid(mlp(self.sd, 'bg_net', 2, h))

    # -- model.py:335-346 ------------------------------------------------------------------------
    def pose_optimisation(self, rays_o, rays_d, frame_ids):
        ids = frame_ids.reshape(-1)
        data = self.sd['pose_array.data']
        R = rotation_matrices(data[:, 0:3][ids])
        tr = data[:, 3:6][ids]
        return rays_o + tr, (rays_d[..., None, :] * R).sum(-1)


def init_reference_like_state(num_frames=200, seed=0, dtype=torch.float32, emb_scale=1e-4, randomize=False, sphere=False):
    """Seeded state_dict with the reference's shapes/names.  With
    randomize=False it follows the reference initialisers (geometric init for sdf_net,
    models/decoders.py:24-43; weight_norm g=||v||; U(-1e-4,1e-4) tables, grid.py:145-147);
    randomize=True perturbs everything (g, biases, beta, poses, big tables) so parity tests see
    a 'trained-like' generic model.  sphere=True (with randomize=True) keeps the geometric-init SDF net untouched
    (sdf ~ |x| - 0.4: rays cross a surface, so sigma / weights / depth are well-conditioned, SURVEY.md 8d "trained-like
    sphere SDF") and shrinks the deformation output so the sphere survives the warp."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(prefix, fan_in, fan_out, wn):
        bnd = 1 / math.sqrt(fan_in)
        v = (torch.rand(fan_out, fan_in, generator=g) * 2 - 1) * bnd  # kaiming_uniform(a=sqrt(5))
        b = (torch.rand(fan_out, generator=g) * 2 - 1) * bnd
        if wn:
            sd[prefix + '.weight_v'] = v
            gg = v.norm(dim=1, keepdim=True)
            if randomize:
                gg = gg * (0.5 + torch.rand(fan_out, 1, generator=g))
            sd[prefix + '.weight_g'] = gg
        else:
            sd[prefix + '.weight'] = v
        sd[prefix + '.bias'] = b

    for name, dout in (('deform_net', 3), ('topo_net', 2)):
        dims = [87, 128, 128, 128, 128, 128, dout]
        for l in range(6):
            lin(f'{name}.net.{l}', dims[l], dims[l + 1], True)
    # sdf_net: geometric init (decoders.py:24-43), no weight norm (model.py:169-171)
    dims = [73, 64, 64, 33]
    for l in range(3):
        W = torch.zeros(dims[l + 1], dims[l])
        b = torch.zeros(dims[l + 1])
        if l == 2:
            W = math.sqrt(math.pi) / math.sqrt(dims[l]) + 1e-4 * torch.randn(dims[l + 1], dims[l], generator=g)
            b = torch.full((dims[l + 1],), -0.4)
        elif l == 0:
            W[:, :3] = torch.randn(dims[l + 1], 3, generator=g) * (math.sqrt(2) / math.sqrt(dims[l + 1]))
        else:
            W = torch.randn(dims[l + 1], dims[l], generator=g) * (math.sqrt(2) / math.sqrt(dims[l + 1]))
        if randomize and not sphere:
            W = W + 0.05 * torch.randn(W.shape, generator=g)
            b = b + 0.05 * torch.randn(b.shape, generator=g)
        sd[f'sdf_net.net.{l}.weight'] = W
        sd[f'sdf_net.net.{l}.bias'] = b
    dims = [64, 64, 64, 3]
    for l in range(3):
        lin(f'color_net.net.{l}', dims[l], dims[l + 1], True)
    dims = [52, 32, 3]
    for l in range(2):
        lin(f'bg_net.net.{l}', dims[l], dims[l + 1], True)
    for i, s in enumerate([num_frames // 8, num_frames // 4, num_frames]):
        sd[f'deform_code.volumes.{i}'] = torch.randn(1, 16, s, 1, generator=g)
    offs = torch.from_numpy(make_offsets(3, 16, 16, per_level_scale(16, 128, 16), 15))
    for enc in ('encoder', 'encoder_c'):
        sd[enc + '.offsets'] = offs.clone()
        sd[enc + '.embeddings'] = (torch.rand(int(offs[-1]), 2, generator=g) * 2 - 1) * emb_scale
    sd['sdf2density.beta'] = torch.tensor(0.1 if not randomize else (0.05 if sphere else 0.037))
    if sphere:
        sd['deform_net.net.5.weight_g'] = sd['deform_net.net.5.weight_g'] * 0.2
    sd['pose_array.data'] = torch.zeros(num_frames, 6) if not randomize else 0.02 * torch.randn(num_frames, 6, generator=g)
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}

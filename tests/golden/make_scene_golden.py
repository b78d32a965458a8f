"""Generate tests/golden/scene_*.npz by running the UNMODIFIED reference Python
(/root/reference/models/model.py and friends) on CPU with seeded weights.

Run in the build container only (needs /root/reference):
    python tests/golden/make_scene_golden.py

Two shims are required to import the reference on a GPU-less box:
  * torch.Tensor.cuda -> no-op (models/density.py:20 calls .cuda() in __init__),
  * `external.encoders.gridencoder.grid` -> a module exposing oracle.grid.GridEncoderOracle as
    GridEncoder (the reference encoder is CUDA-only; its own numerics are pinned separately by
    tests/golden/gridencoder_ref_sm100.npz, generated on a B200 from the reference kernel).
Every other line executed is the reference's own.
"""
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = '/root/reference'
sys.path.insert(0, REPO)
from oracle.fields import init_reference_like_state  # noqa: E402
from oracle.grid import GridEncoderOracle  # noqa: E402

sys.path.insert(0, REF)
torch.Tensor.cuda = lambda self, *a, **k: self
stub = types.ModuleType('external.encoders.gridencoder.grid')
stub.GridEncoder = GridEncoderOracle
for name in ('external', 'external.encoders', 'external.encoders.gridencoder'):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules['external.encoders.gridencoder.grid'] = stub

from models.model import scene_representation  # noqa: E402  (reference, unmodified)

OUT = os.path.dirname(os.path.abspath(__file__))
CONFIG = {'model': {'bg_radius': 1.4, 'activation': 'exp'}}
F = 200
BOUND = 1.01


def build(seed, randomize, emb_scale):
    torch.manual_seed(seed)
    m = scene_representation(CONFIG, BOUND, num_frames=F, deform_dim=16, use_app=False, use_t=False,
                             amb_dim=2, color_grid=True, use_joint=True, encode_topo=False)
    sd = init_reference_like_state(F, seed=seed, randomize=randomize, emb_scale=emb_scale)
    missing = m.load_state_dict(sd, strict=True)
    return m, sd


def points(seed, M):
    g = torch.Generator().manual_seed(1000 + seed)
    x = (torch.rand(M, 3, generator=g) * 2 - 1) * 0.9
    x[:4] = torch.tensor([[BOUND, 0, 0], [-BOUND, BOUND, -BOUND], [0, 0, 0], [1.2, 0, 0]])  # boundary, OOB
    t = torch.full((M, 1), 37.0 / F)
    t[M // 2:] = 151.0 / F  # two frames in one batch (generic per-sample t)
    t[-1] = 1.0
    t[-2] = 0.0
    light = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1)
    return x, t, light


def _nz_min(o):
    a = o.abs()
    a = a[a > 0]   # exact zeros (e.g. the geo-init sdf_net at the origin) are deterministic in every implementation
    return a.min().item() if a.numel() else 1.0


def relu_margin(m, x, t, light):
    """min |pre-activation| over every hidden Linear output in one albedo_normal forward: a ReLU whose
    input is within fp32 rounding of 0 may flip between implementations and makes *gradient* goldens
    ambiguous, so the generator rejects point sets with a tiny margin."""
    margins = []
    hooks = []
    for net in (m.deform_net, m.topo_net, m.sdf_net, m.color_net):
        for lin in list(net.net)[:-1]:
            hooks.append(lin.register_forward_hook(lambda mod, i, o: margins.append(_nz_min(o.detach()))))
    with torch.no_grad():
        m(x, t, light, ratio=1.0, shading='albedo_normal')
    for h in hooks:
        h.remove()
    return min(margins)


def run_case(tag, seed, randomize, emb_scale, max_level, M=96):
    m, sd = build(seed, randomize, emb_scale)
    m.max_level = max_level
    pseed = seed
    for _attempt in range(40):
        x, t, light = points(pseed, M)
        mg = relu_margin(m, x, t, light)
        if mg > 2e-6:
            break
        print(f'{tag}: point seed {pseed} rejected (ReLU margin {mg:.2e})')
        pseed += 100
    out = {'x': x.numpy(), 't': t.numpy(), 'light': light.numpy(), 'max_level': np.array(-1.0 if max_level is None else max_level),
           'seed': np.array(seed), 'relu_margin': np.array(mg), 'randomize': np.array(int(randomize)), 'emb_scale': np.array(emb_scale)}
    with torch.no_grad():
        for shading, ratio in (('albedo', 1.0), ('albedo_normal', 1.0), ('lambertian', 0.3), ('textureless', 0.55), ('normal', 1.0)):
            sdf, sigma, color, normal, deform, raw = m(x, t, light, ratio=ratio, shading=shading)
            out[f'{shading}.sdf'], out[f'{shading}.sigma'], out[f'{shading}.color'] = sdf.numpy(), sigma.numpy(), color.numpy()
            out[f'{shading}.deform'] = deform.numpy()
            if normal is not None:
                out[f'{shading}.normal'], out[f'{shading}.normal_raw'] = normal.numpy(), raw.numpy()
        d = m.density(x, t)
        out['density.sdf'], out['density.sigma'], out['density.albedo'] = d['sdf'].numpy(), d['sigma'].numpy(), d['albedo'].numpy()
        d = m.density(x, None)
        out['density_cano.sdf'], out['density_cano.albedo'] = d['sdf'].numpy(), d['albedo'].numpy()
        d = m.density(x, t[:3], allow_shape=True, return_color=False)
        out['density_allow_shape.sigma'] = d['sigma'].numpy()
        n, raw = m.normal(x, t=t)
        out['normal_warped.normal'], out['normal_warped.raw'] = n.numpy(), raw.numpy()
        n, raw = m.normal(x, topo=None)
        out['normal_cano.normal'], out['normal_cano.raw'] = n.numpy(), raw.numpy()
        deform, topo, _ = m.warp(x, t)
        out['warp.deform'], out['warp.topo'] = deform.numpy(), topo.numpy()
        out['code'] = m.get_deform_code(t).numpy()
        out['background'] = m.background(light, t).numpy()
        ids = torch.randint(0, F, (M, 1), generator=torch.Generator().manual_seed(seed))
        o2, d2 = m.pose_optimisation(x, light, ids)
        out['pose.ids'], out['pose.o'], out['pose.d'] = ids.numpy(), o2.numpy(), d2.numpy()
    # gradients of a scalar through the full 'albedo_normal' forward (param-grad parity)
    m.zero_grad()
    xg = x.clone().requires_grad_(True)
    sdf, sigma, color, normal, deform, raw = m(xg, t, light, ratio=1.0, shading='albedo_normal')
    g = torch.Generator().manual_seed(77)
    loss = (sdf * torch.randn(M, generator=g)).sum() + (sigma * torch.randn(M, generator=g)).sum() * 1e-2 \
        + (color * torch.randn(M, 3, generator=g)).sum() + (normal * torch.randn(M, 3, generator=g)).sum() \
        + (deform * torch.randn(M, 3, generator=g)).sum()
    loss.backward()
    out['grad.x'] = xg.grad.numpy()
    for k, p in m.named_parameters():
        if p.grad is not None:
            out['grad.' + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(OUT, f'scene_{tag}.npz'), **out)
    print(tag, 'saved', len(out), 'arrays')


if __name__ == '__main__':
    run_case('init_full', seed=1, randomize=False, emb_scale=1e-4, max_level=None)
    run_case('rand_c2f', seed=2, randomize=True, emb_scale=0.5, max_level=0.53)
    run_case('rand_full', seed=3, randomize=True, emb_scale=0.5, max_level=1.0)

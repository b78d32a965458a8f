"""GPU: Zero123.train_step of morpheus_b200.guidance (C-ABI kernels mb_add_noise / mb_sds_grad, CUDA-graphed UNet, frozen VAE
with input-gradient backward) against the closed-form oracle chain (oracle/sds.py) evaluated on the CPU with the same
seeded weights and injected draws (t, noise, VAE noise) -- BASELINE cfg-3 with random weights (the Zero-1-to-3 checkpoint
is not downloadable here).  Bar: SDS gradient within 1e-3 (north_star)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ldm_util import load_key_table, seeded_state  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@pytest.mark.parametrize('graph', [False, True])
def test_train_step_vs_oracle(graph):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from morpheus_b200 import guidance
    from oracle import sds as osds
    dev = torch.device('cuda:0')
    table = load_key_table()
    sd = {}
    sd.update(seeded_state(table['unet'], 1, 'model.diffusion_model.'))
    sd.update(seeded_state(table['encoder'], 2, 'first_stage_model.encoder.'))
    sd.update(seeded_state(table['quant_conv'], 3, 'first_stage_model.quant_conv.'))
    sd.update(seeded_state(table['cc_projection'], 4, 'cc_projection.'))
    z123 = guidance.Zero123(dev, state_dict=sd, t_range=[0.02, 0.5], graph=graph)
    g = torch.Generator().manual_seed(3)
    emb = {'c_crossattn': [torch.randn(1, 1, 768, generator=g)], 'c_concat': [torch.randn(1, 4, 32, 32, generator=g)],
           'ref_radii': [2.5], 'ref_polars': [90.0], 'ref_azimuths': [0.0], 'zero123_ws': [1]}
    pred = torch.rand(1, 3, 72, 72, generator=g)
    polar, azimuth, radius = torch.tensor([10.0]), torch.tensor([200.0]), torch.tensor([0.1])
    t = torch.tensor([260])
    noise = torch.randn(1, 4, 32, 32, generator=g)
    vae_noise = torch.randn(1, 4, 32, 32, generator=g)
    # ---- ours, twice when graphed (capture, then replay) ----
    for rep in range(2 if graph else 1):
        pg = pred.clone().to(dev).requires_grad_(True)
        loss, t_out, gs, noise_out = z123.train_step(emb, pg, polar, azimuth, radius, guidance_scale=5, grad_scale=0.01, t=t.to(dev),
                                                     noise=noise.to(dev), vae_noise=vae_noise.to(dev))
        loss.backward()
    # ---- oracle chain on the CPU (same functional nets are validated against the reference classes in test_sds_cpu.py) ----
    unet = guidance._KeyIndex({k[len('model.diffusion_model.'):]: v for k, v in sd.items() if k.startswith('model.diffusion_model.')})
    vae = guidance._KeyIndex({k[len('first_stage_model.'):]: v for k, v in sd.items() if k.startswith('first_stage_model.')})
    ac = osds.alphas_cumprod()
    pc = pred.clone().requires_grad_(True)
    img = torch.nn.functional.interpolate(pc, (256, 256), mode='bilinear', align_corners=False)
    mean, logvar = guidance.vae_encode_moments(vae, img * 2 - 1).chunk(2, dim=1)
    lat = 0.18215 * (mean + torch.exp(0.5 * logvar.clamp(-30, 20)) * vae_noise)
    ang = osds.angle_between_deg(torch.stack([radius + 2.5, torch.deg2rad(polar + 90.0), torch.deg2rad(azimuth + 0.0)], -1),
                                 torch.tensor([[2.5, np.deg2rad(90.0), 0.0]]))
    grad_scale = (torch.exp(ang.min(dim=1)[0] / 180.0) - 1) * 0.01
    with torch.no_grad():
        T = osds.pose_token(polar, azimuth, radius)
        clip = torch.nn.functional.linear(torch.cat([emb['c_crossattn'][0], T], -1), sd['cc_projection.weight'], sd['cc_projection.bias'])
        x_in = torch.cat([torch.cat([osds.add_noise(lat.detach(), noise, t, ac)] * 2), torch.cat([torch.zeros(1, 4, 32, 32), emb['c_concat'][0]])], 1)
        eps = guidance.unet_forward(unet, x_in, torch.cat([t, t]), torch.cat([torch.zeros_like(clip), clip]))
        grad = osds.sds_grad(eps[0:1], eps[1:2], noise, t, ac, 5.0, grad_scale)
    lo = osds.sds_loss(lat, grad)
    lo.backward()
    assert abs(float(gs) - float(grad_scale)) < 1e-6 * max(1.0, abs(float(grad_scale)))
    assert rel(loss, lo) < 1e-3
    assert rel(pg.grad, pc.grad) < 1e-3, rel(pg.grad, pc.grad)     # d loss / d pred_rgb: SDS gradient pulled through the VAE encoder

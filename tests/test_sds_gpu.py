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
    emb_dev = dict(emb, c_crossattn=[c.to(dev) for c in emb['c_crossattn']], c_concat=[c.to(dev) for c in emb['c_concat']])
    t_dev, noise_dev, vae_noise_dev = t.to(dev), noise.to(dev), vae_noise.to(dev)
    for rep in range(3 if graph else 2):
        pg = pred.clone().to(dev).requires_grad_(True)
        torch.cuda.synchronize()
        # from the second call on (graphs captured, cuDNN plans cached) the step must not synchronise with the device at all:
        # abar_t, w(t) and the angle weights never come back to the host (zero123_utils.py:161-212 reads them with .item()-style ops)
        torch.cuda.set_sync_debug_mode('error' if rep > 0 else 'default')
        try:
            loss, t_out, gs, noise_out = z123.train_step(emb_dev, pg, polar, azimuth, radius, guidance_scale=5, grad_scale=0.01, t=t_dev,
                                                         noise=noise_dev, vae_noise=vae_noise_dev)
            with guidance._precision('fp32'):      # the VAE backward convolutions must not drop to TF32 either
                loss.backward()
        finally:
            torch.cuda.set_sync_debug_mode('default')
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


def test_virtual_view_step_end_to_end():
    """BASELINE cfg-3: 72x72 novel view -> occupancy-grid sampling -> fused render (lambertian) -> pred_rgb -> SDS -> backward
    through compositing and the field kernels -> fused Adam.  Checks that the SDS gradient reaches the scene parameters."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from morpheus_b200 import guidance, rays
    from morpheus_b200 import train as mtrain
    from morpheus_b200.model import scene_representation
    from morpheus_b200.nerfacc_compat import OccGridEstimator
    from morpheus_b200.render import Renderer
    dev = torch.device('cuda:0')
    table = load_key_table()
    sd = {}
    sd.update(seeded_state(table['unet'], 1, 'model.diffusion_model.'))
    sd.update(seeded_state(table['encoder'], 2, 'first_stage_model.encoder.'))
    sd.update(seeded_state(table['quant_conv'], 3, 'first_stage_model.quant_conv.'))
    sd.update(seeded_state(table['cc_projection'], 4, 'cc_projection.'))
    z123 = guidance.Zero123(dev, state_dict=sd, t_range=[0.02, 0.5], precision='tf32')
    torch.manual_seed(0)
    cfg = {'model': {'bg_radius': 1.4, 'activation': 'exp'}, 'render': {'step_size': 0.01}, 'train': dict(mtrain.DEFAULT_TRAIN_CFG)}
    model = scene_representation(cfg, 1.01, num_frames=200, deform_dim=16, use_app=False, use_t=False, amb_dim=2, color_grid=True,
                                 use_joint=True, encode_topo=False).to(dev).train()
    model.max_level = 0.75
    est = OccGridEstimator(torch.tensor([-1.01] * 3 + [1.01] * 3), 128).to(dev).train()
    R = Renderer(model, est, cfg, 200)
    opt = mtrain.FlatAdam(model, 5e-4)
    view = rays.virtual_view_rays(frame=12, num_frames=200, H=360, W=360, focal=517.0, scale=0.2, generator=torch.Generator().manual_seed(1), device=dev)
    R.update_occ_grid(view['rays_t'].reshape(-1, 1), step=0)                       # morpheus.py:1189 (first refresh covers all 128^3 cells)
    assert 0 < int(est.binaries.sum()) < 128 ** 3
    g = torch.Generator().manual_seed(2)
    emb = {'c_crossattn': [torch.randn(1, 1, 768, generator=g)], 'c_concat': [torch.randn(1, 4, 32, 32, generator=g)],
           'ref_radii': [2.5], 'ref_polars': [90.0], 'ref_azimuths': [0.0], 'zero123_ws': [1]}
    before = opt.flat.clone()
    loss, out = mtrain.virtual_view_step(R, z123, opt, view, emb, cfg['train'], shading='lambertian', ambient_ratio=0.4,
                                         bg_color=torch.rand(3, device=dev))
    torch.cuda.synchronize()
    assert out['image'].shape == (1, 72 * 72, 3) and torch.isfinite(loss)
    # (the fused Adam launch clears the gradient buffer: the SDS gradient is observed through the parameter updates)
    assert torch.isfinite(opt.flat).all() and float(opt.grad.abs().max()) == 0
    moved = {n: float((opt.flat[a:b] - before[a:b]).abs().max()) for n, (a, b) in opt.group_slices.items()}
    # (fresh geometric init: the SDF net's grid / topology columns are zero, so the SDF table and the topology net get exactly zero gradient)
    for n in ('encoder_color', 'decoder_sdf', 'decoder_color', 'decoder_deform', 'density'):
        assert moved[n] > 0, moved
    assert moved['pose'] == 0 and moved['decoder_bg'] == 0, moved      # no gradient on a virtual view: torch.optim.Adam skips them

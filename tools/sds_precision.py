"""Accuracy / speed of the arithmetic modes of the frozen Zero-1-to-3 networks in the SDS step (north_star bar: SDS gradient within 1e-3).
Truth = the same functional networks evaluated in FLOAT64 on the GPU (same seeded random weights, same injected t / noise / VAE noise).
For every mode: relative L2 error of d loss / d pred_rgb (the SDS gradient pulled through the VAE encoder) and of the loss, and the
CUDA-event time of train_step forward + backward (whole-chain CUDA graph, median of 10).
    python tools/sds_precision.py        (run under gpurun; prints one JSON document)"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from ldm_util import load_key_table, seeded_state  # noqa: E402
from morpheus_b200 import guidance  # noqa: E402

dev = torch.device('cuda:0')
table = load_key_table()
sd = {}
sd.update(seeded_state(table['unet'], 1, 'model.diffusion_model.'))
sd.update(seeded_state(table['encoder'], 2, 'first_stage_model.encoder.'))
sd.update(seeded_state(table['quant_conv'], 3, 'first_stage_model.quant_conv.'))
sd.update(seeded_state(table['cc_projection'], 4, 'cc_projection.'))
g = torch.Generator().manual_seed(3)
emb = {'c_crossattn': [torch.randn(1, 1, 768, generator=g).to(dev)], 'c_concat': [torch.randn(1, 4, 32, 32, generator=g).to(dev)],
       'ref_radii': [2.5], 'ref_polars': [90.0], 'ref_azimuths': [0.0], 'zero123_ws': [1]}
pred = torch.rand(1, 3, 72, 72, generator=g).to(dev)
polar, azimuth, radius = torch.tensor([10.0]), torch.tensor([200.0]), torch.tensor([0.1])
t = torch.tensor([260], device=dev)
noise = torch.randn(1, 4, 32, 32, generator=g).to(dev)
vae_noise = torch.randn(1, 4, 32, 32, generator=g).to(dev)


def rel(a, b):
    a, b = a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


# ---- truth: float64 ----
z64 = guidance.Zero123(dev, state_dict=sd, t_range=[0.02, 0.5], precision='fp32')
for part in (z64.unet, z64.vae, z64.cc):
    for k in list(part.keys()):
        part[k] = part[k].double()
z64.alphas = z64.alphas.double()
pc = pred.double().clone().requires_grad_(True)
img = F.interpolate(pc, (256, 256), mode='bilinear', align_corners=False)
mean, logvar = guidance.vae_encode_moments(z64.vae, img * 2 - 1).chunk(2, dim=1)
lat = 0.18215 * (mean + torch.exp(0.5 * logvar.clamp(-30, 20)) * vae_noise.double())
gs_host, views = z64._view_weights(emb, polar, azimuth, radius, 0.01)
with torch.no_grad():
    ab = z64.alphas[t]
    noisy = ab.sqrt() * lat + (1 - ab).sqrt() * noise.double()
    T = views[0][2].to(dev).double()
    clip = F.linear(torch.cat([emb['c_crossattn'][0].double(), T], -1), z64.cc['weight'], z64.cc['bias'])
    x_in = torch.cat([torch.cat([noisy] * 2), torch.cat([torch.zeros(1, 4, 32, 32, device=dev, dtype=torch.float64), emb['c_concat'][0].double()])], 1)
    eps = guidance.unet_forward(z64.unet, x_in, torch.cat([t, t]), torch.cat([torch.zeros_like(clip), clip]))
    grad = float(gs_host) * (1 - ab) * (eps[0:1] + 5.0 * (eps[1:2] - eps[0:1]) - noise.double())
loss64 = 0.5 * F.mse_loss(lat, (lat - grad).detach(), reduction='sum')
loss64.backward()
g64 = pc.grad.clone()
del z64
torch.cuda.empty_cache()

out = {'truth': 'float64 functional networks on the GPU', 'bar': 1e-3, 'modes': {}}
for mode in ('fp32', 'reference', 'tf32', 'bf16'):
    z = guidance.Zero123(dev, state_dict=sd, t_range=[0.02, 0.5], precision=mode, graph=True)
    ts = []
    for it in range(13):
        pg = pred.clone().requires_grad_(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        loss, _, _, _ = z.train_step(emb, pg, polar, azimuth, radius, guidance_scale=5, grad_scale=0.01, t=t, noise=noise, vae_noise=vae_noise)
        with guidance._precision(mode):
            loss.backward()
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(e0.elapsed_time(e1))
    out['modes'][mode] = {'grad_rel_err': rel(pg.grad, g64), 'loss_rel_err': abs(float(loss) - float(loss64)) / abs(float(loss64)),
                          'ms_fwd_bwd': float(np.median(ts)), 'passes_1e-3': rel(pg.grad, g64) < 1e-3,
                          'what': {'fp32': 'true fp32 convolutions and matmuls', 'reference': 'stock PyTorch defaults the reference runs with on this GPU: TF32 cuDNN convolutions, fp32 matmuls',
                                   'tf32': 'TF32 convolutions and matmuls', 'bf16': 'bf16 autocast UNet (no whole-chain graph), TF32 VAE'}[mode]}
    del z
    torch.cuda.empty_cache()
print(json.dumps(out))

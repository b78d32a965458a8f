"""CUDA-event timing of the virtual-view (SDS) step, BASELINE cfg-3 with seeded random Zero-1-to-3 weights (the checkpoint
is not downloadable offline):  python tools/time_virtual.py [fp32|tf32] [steps]      (run under gpurun)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from ldm_util import load_key_table, seeded_state  # noqa: E402
from morpheus_b200 import guidance, rays  # noqa: E402
from morpheus_b200 import train as mtrain  # noqa: E402
from morpheus_b200.model import scene_representation  # noqa: E402
from morpheus_b200.nerfacc_compat import OccGridEstimator  # noqa: E402
from morpheus_b200.render import Renderer  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device('cuda:0')
table = load_key_table()
sd = {}
sd.update(seeded_state(table['unet'], 1, 'model.diffusion_model.'))
sd.update(seeded_state(table['encoder'], 2, 'first_stage_model.encoder.'))
sd.update(seeded_state(table['quant_conv'], 3, 'first_stage_model.quant_conv.'))
sd.update(seeded_state(table['cc_projection'], 4, 'cc_projection.'))
z123 = guidance.Zero123(dev, state_dict=sd, t_range=[0.02, 0.5], precision=prec)
torch.manual_seed(0)
cfg = {'model': {'bg_radius': 1.4, 'activation': 'exp'}, 'render': {'step_size': 0.01}, 'train': dict(mtrain.DEFAULT_TRAIN_CFG)}
model = scene_representation(cfg, 1.01, num_frames=200, deform_dim=16, use_app=False, use_t=False, amb_dim=2, color_grid=True,
                             use_joint=True, encode_topo=False).to(dev).train()
model.max_level = 0.75
est = OccGridEstimator(torch.tensor([-1.01] * 3 + [1.01] * 3), 128).to(dev).train()
R = Renderer(model, est, cfg, 200)
opt = mtrain.FlatAdam(model, 5e-4)
g = torch.Generator().manual_seed(2)
emb = {'c_crossattn': [torch.randn(1, 1, 768, generator=g)], 'c_concat': [torch.randn(1, 4, 32, 32, generator=g)],
       'ref_radii': [2.5], 'ref_polars': [90.0], 'ref_azimuths': [0.0], 'zero123_ws': [1]}
times, n_samples = [], 0
for it in range(steps + 2):
    view = rays.virtual_view_rays(frame=(7 * it) % 200, num_frames=200, H=360, W=360, focal=517.0, scale=0.2,
                                  generator=torch.Generator().manual_seed(it), device=dev)
    if it == 0:
        R.update_occ_grid(view['rays_t'].reshape(-1, 1), step=0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    loss, out = mtrain.virtual_view_step(R, z123, opt, view, emb, cfg['train'], shading='lambertian', ambient_ratio=0.4, bg_color=torch.rand(3, device=dev))
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
    n_samples = int(out['weights'].shape[0])
ms = sorted(times[2:])[len(times[2:]) // 2]
print(json.dumps({'workload': 'virtual-view SDS step, 72x72 rays, occupancy-grid sampling, lambertian, random-weight Zero-1-to-3', 'precision': prec,
                  'ms_per_step': ms, 'rays_per_s': 72 * 72 / (ms * 1e-3), 'packed_samples': n_samples, 'loss': float(loss)}))

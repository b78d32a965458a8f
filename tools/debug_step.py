"""Per-term comparison of one real-view step (ours on the GPU vs the CPU oracle) -- diagnostic for tests/test_step_parity_gpu.py.
    python tools/debug_step.py [n_rays] [samples]"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_step_parity_gpu import build_ours, rel_l2, trained_like_state  # noqa: E402
from morpheus_b200 import train as mtrain  # noqa: E402
from morpheus_b200.rays import synthetic_real_view_batch  # noqa: E402
from oracle import train_step as ots  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ML = 0.8
dev = torch.device('cuda:0')
sd = trained_like_state(21)
batch = synthetic_real_view_batch(N, seed=5, frame=63)
g = torch.Generator().manual_seed(9)
jitter = torch.rand(N, generator=g)
noise = torch.randn(N * S, 3, generator=g)
params_o = ots.make_params(sd)
loss_o, out_o = ots.step_loss(params_o, batch, S, ML, jitter=jitter, perturb_noise=noise)
m, R, opt, tr = build_ours(sd, dev, S, ML)
b = {k: v.to(dev) for k, v in batch.items()}
out = R.render_rays(b['rays_o'], b['rays_d'], b['rays_t'], b['rays_id'], bg_color=b['bg'], shading='albedo_normal', real_view=True,
                    rays_depth=b['depth'], rays_mask=b['mask'], optimize_pose=True, jitter=jitter.to(dev), perturb_noise=noise.to(dev))
for k in ('image', 'depth', 'weights_sum', 'sdf', 'weights', 'normal', 'deform'):
    print(f'{k:12s} rel-L2 {rel_l2(out[k].reshape(-1), out_o[k].reshape(-1)):.2e}')
for k in ('sdf_loss', 'fs_loss', 'loss_normal_perturb'):
    print(f'{k:20s} ours {float(out[k]):.8f} oracle {float(out_o[k]):.8f} rel {abs(float(out[k]) - float(out_o[k])) / abs(float(out_o[k]) + 1e-30):.2e}')
gt_depth, gt_mask = batch['depth'].reshape(-1), batch['mask'].reshape(-1)
xyz = batch['rays_o'] + gt_depth[:, None] * batch['rays_d']
dm = ((gt_depth > 0) & (xyz.norm(dim=-1) <= 1.1) & (gt_mask > 0.5)).float()
for name, src in (('ours', {k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in out.items()}), ('oracle', out_o)):
    img, dep, ws = src['image'].reshape(-1, 3), src['depth'].reshape(-1), src['weights_sum'].reshape(-1)
    print(name, 'rgb', float(5 * F.mse_loss(img, batch['rgb'])), 'mask', float(0.5 * F.binary_cross_entropy(ws.clip(1e-5, 1 - 1e-5), gt_mask)),
          'depth', float(0.1 * F.mse_loss(dep * dm, gt_depth * dm)), 'code', float(src['loss_code']) if 'loss_code' in src else None)
print('total ours', float(mtrain.real_view_loss(out, b, m, tr)), 'oracle', float(loss_o))

""""Equal loss" evidence for the north star (>= 10x the reference's single-GPU rays/s AT EQUAL LOSS): the same N optimiser steps from the
same initial state on the same batches with the same injected RNG draws (stratified jitter, perturbation noise), run by
  (a) this library (train.train_step, CUDA kernels through the C ABI) and
  (b) the reference GPU path (oracle/ref_gpu_step.py: eager torch restatement + the UNMODIFIED reference gridencoder kernel + torch.optim.Adam),
and the two loss curves compared step by step.  Prints one JSON document (loss curves, max / mean relative difference, wall-clock of both).
    python tools/loss_curve.py [steps] [n_rays] [samples]      (run under gpurun)"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench  # noqa: E402
from morpheus_b200 import train as mtrain  # noqa: E402
from morpheus_b200.nerfacc_compat import OccGridEstimator  # noqa: E402
from morpheus_b200.rays import synthetic_real_view_batch  # noqa: E402
from morpheus_b200.render import Renderer  # noqa: E402
from oracle import fields as of  # noqa: E402
from oracle import ref_gpu_step as rgs  # noqa: E402
from oracle import train_step as ots  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
S = int(sys.argv[3]) if len(sys.argv) > 3 else 128
dev = torch.device('cuda:0')
with torch.device('cpu'):
    m0 = bench.make_state()
sd = {k: v.detach().clone() for k, v in m0.state_dict().items()}
lr = mtrain.DEFAULT_TRAIN_CFG['lr']


def draws(i):
    with torch.device('cpu'):
        g = torch.Generator().manual_seed(10_000 + i)
        b = synthetic_real_view_batch(N, seed=3000 + i, frame=(37 * i) % bench.NUM_FRAMES)
        jit, noi = torch.rand(N, generator=g), torch.randn(N * S, 3, generator=g)
    return {k: v.to(dev) for k, v in b.items()}, jit.to(dev), noi.to(dev)


# ---- (a) ours ----
model = bench.make_state().to(dev).train()
tr = dict(mtrain.DEFAULT_TRAIN_CFG)
R = Renderer(model, OccGridEstimator(torch.tensor([-1.01] * 3 + [1.01] * 3), 128).to(dev), dict(bench.CONFIG, train=tr), bench.NUM_FRAMES, uniform_samples=S)
opt = mtrain.FlatAdam(model, lr)
ours = []
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(steps):
    b, jit, noi = draws(i)
    ours.append(mtrain.train_step(R, opt, b, tr, jitter=jit, perturb_noise=noi))
ours = [float(x) for x in ours]
torch.cuda.synchronize()
t_ours = time.perf_counter() - t0

# ---- (b) reference GPU path ----
backend = rgs.load_ref_backend()
res = {'steps': steps, 'rays': N, 'samples_per_ray': S, 'ours': ours}
if backend is None:
    res['reference_gpu'] = {'unavailable': 'oracle/_ref not built'}
else:
    grid_fn = rgs.make_grid_fn(backend)

    def grid(self, which, x):
        u = (x + self.bound) / (2 * self.bound)
        return grid_fn.apply(u, self.sd[which + '.embeddings'], self.sd[which + '.offsets'], self.S, self.H, u.requires_grad, self.max_level)
    of.SceneOracle.grid = grid
    with torch.device(dev):
        params = ots.make_params({k: v.to(dev) for k, v in sd.items()})
        groups = [{'params': [v], 'lr': lr * (0.1 if n == 'pose_array.data' else 0.5 if n == 'sdf2density.beta' else 1.0)}
                  for n, v in params.items() if v.is_floating_point() and v.requires_grad]
        topt = torch.optim.Adam(groups, lr=lr, betas=(0.9, 0.99), eps=1e-15)
        ref = []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(steps):
            b, jit, noi = draws(i)
            topt.zero_grad()
            loss, _ = ots.step_loss(params, b, S, bench.MAX_LEVEL, jitter=jit, perturb_noise=noi)
            loss.backward()
            topt.step()
            ref.append(loss.detach())
        ref = [float(x) for x in ref]
        torch.cuda.synchronize()
        t_ref = time.perf_counter() - t0
    rel = [abs(a - b) / abs(b) for a, b in zip(ours, ref)]
    res.update(reference_gpu=ref, max_rel_diff=max(rel), mean_rel_diff=sum(rel) / len(rel), rel_diff_last_10=rel[-10:],
               final_loss={'ours': ours[-1], 'reference_gpu': ref[-1]}, wall_s={'ours_eager': t_ours, 'reference_gpu': t_ref},
               what='same init (bench.make_state), same batches, same injected jitter / perturbation noise; ours = train.train_step (eager launches), '
                    'reference = oracle.train_step on the GPU + unmodified reference gridencoder kernel + torch.optim.Adam')
print(json.dumps(res))

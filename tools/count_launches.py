"""Count the CUDA kernels one eager real-view training step launches (torch.profiler), grouped by name and by the
coarse phase of the step.  python tools/count_launches.py [n_rays]   (run under gpurun)"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile, record_function

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from morpheus_b200 import train as mtrain  # noqa: E402
from morpheus_b200.nerfacc_compat import OccGridEstimator  # noqa: E402
from morpheus_b200.rays import synthetic_real_view_batch  # noqa: E402
from morpheus_b200.render import Renderer  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device('cuda:0')
model = bench.make_state().to(dev).train()
tr = dict(mtrain.DEFAULT_TRAIN_CFG)
cfg = dict(bench.CONFIG, train=tr)
R = Renderer(model, OccGridEstimator(torch.tensor([-1.01] * 3 + [1.01] * 3), 128).to(dev), cfg, bench.NUM_FRAMES, uniform_samples=bench.N_SAMPLES)
opt = mtrain.FlatAdam(model, tr['lr'])
b = {k: v.to(dev) for k, v in synthetic_real_view_batch(N, seed=1).items()}
for _ in range(3):
    mtrain.train_step(R, opt, b, tr, 1)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    opt.zero_grad()
    with record_function('PH_render'):
        out = R.render_rays(b['rays_o'], b['rays_d'], b['rays_t'], b['rays_id'], bg_color=b['bg'], shading='albedo_normal', real_view=True,
                            rays_depth=b['depth'], rays_mask=b['mask'], optimize_pose=True)
    with record_function('PH_loss'):
        loss = mtrain.real_view_loss(out, b, model, tr)
    with record_function('PH_backward'):
        loss.backward()
    with record_function('PH_opt'):
        opt.all_reduce()
        opt.step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
phases = [e for e in prof.events() if e.name.startswith('PH_') and e.device_type == torch.autograd.DeviceType.CPU]
print('total CUDA kernels/memops in one step:', len(ev))
import collections
cnt = collections.Counter()
tm = collections.Counter()
for e in ev:
    cnt[e.name[:90]] += 1
    tm[e.name[:90]] += e.device_time if hasattr(e, 'device_time') else e.cuda_time
for k, v in cnt.most_common(40):
    print(f'{v:5d} {tm[k]:9.1f} us  {k}')
# launches per phase, from the CPU-side launch calls
launch = [e for e in prof.events() if e.name in ('cudaLaunchKernel', 'cudaMemsetAsync', 'cudaMemcpyAsync', 'cuLaunchKernel', 'cudaLaunchKernelExC')]
for ph in phases:
    n = sum(1 for e in launch if e.time_range.start >= ph.time_range.start and e.time_range.end <= ph.time_range.end)
    print(ph.name, n, 'launch calls')

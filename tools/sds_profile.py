"""torch.profiler breakdown of one eager strict-fp32 SDS step (own convolutions on): which kernels the 20 ms go to.
    python tools/sds_profile.py        (run under gpurun)"""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device('cuda:0')
z, emb = bench.make_guidance(dev, 'fp32')
z.graph = False
pol, az, rad = torch.tensor([10.0]), torch.tensor([30.0]), torch.tensor([0.1])
for it in range(3):
    pred = torch.rand(1, 3, 72, 72, device=dev, requires_grad=True)
    z.train_step(emb, pred, pol, az, rad, guidance_scale=5, grad_scale=0.01)[0].backward()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    pred = torch.rand(1, 3, 72, 72, device=dev, requires_grad=True)
    z.train_step(emb, pred, pol, az, rad, guidance_scale=5, grad_scale=0.01)[0].backward()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
cnt, tm = collections.Counter(), collections.Counter()
for e in ev:
    k = e.name[:100]
    cnt[k] += 1
    tm[k] += e.device_time if hasattr(e, 'device_time') else e.cuda_time
T = sum(tm.values())
print(f'{len(ev)} kernels, {T / 1e3:.2f} ms of device time')
for k, v in sorted(tm.items(), key=lambda kv: -kv[1])[:28]:
    print(f'{cnt[k]:5d} {v / 1e3:8.3f} ms {100 * v / T:5.1f}%  {k}')

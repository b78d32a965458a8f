"""Small driver for ncu captures of the fused field kernels (run under gpurun):
   ncu --set full --clock-control none --import-source on -k regex:field_fwd_tc -s 1 -c 1 -o gpurun_out/fwd_tc python tools/profile_field.py fwd
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else 'fwd'
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4096 * 128
dev = torch.device('cuda:0')
m = bench.make_state().to(dev).train()
# "trained-like": the geometric init zeroes the sdf_net columns that read the hash grid (decoders.py:36-38), which would let the
# backward skip the whole table scatter; a small perturbation makes every gradient path live, as after the first Adam step
torch.manual_seed(1)
with torch.no_grad():
    for name, prm in m.named_parameters():
        if 'embeddings' not in name:
            prm.add_(0.02 * torch.randn_like(prm))
m.invalidate()
g = torch.Generator().manual_seed(0)
x = ((torch.rand(M, 3, generator=g) * 2 - 1) * 0.6).to(dev)
t = torch.full((M, 1), 0.3, device=dev)
light = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1).to(dev)
for it in range(3):
    if mode == 'fwd':
        with torch.no_grad():
            out = m(x, t, light, ratio=1.0, shading='albedo_normal')
    elif mode == 'aux':   # the perturbed-normal query of render_rays: 6 SDF queries, topo = None
        n, raw = m.normal(x, topo=None)
        n.sum().backward()
    else:
        xg = x.clone().requires_grad_(True)
        out = m(xg, t, light, ratio=1.0, shading='albedo_normal')
        (out[0].sum() + out[2].sum() + out[3].sum()).backward()
    torch.cuda.synchronize()
print('done')

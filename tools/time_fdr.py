"""CUDA-event timing + per-phase cycle breakdown of the fused FD regulariser (mb_fd_regulariser_tc):
    MB_NVCC_EXTRA=-DMB_FDR_PHASE_TIMING=1 python -m morpheus_b200.build -f && python tools/time_fdr.py [M]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from morpheus_b200 import _lib  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 * 128
dev = torch.device('cuda:0')
m = bench.make_state().to(dev).train()
torch.manual_seed(1)
with torch.no_grad():
    for name, prm in m.named_parameters():
        if 'embeddings' not in name:
            prm.add_(0.02 * torch.randn_like(prm))
m.invalidate()
g = torch.Generator().manual_seed(0)
x = ((torch.rand(M, 3, generator=g) * 2 - 1) * 0.6).to(dev)
# ray-ordered points (128 consecutive samples along a line), like the render's packed samples
rays = M // 128
o = ((torch.rand(rays, 1, 3, generator=g) * 2 - 1) * 0.5)
d = torch.nn.functional.normalize(torch.randn(rays, 1, 3, generator=g), dim=-1)
tt = torch.linspace(-0.6, 0.6, 128).view(1, 128, 1)
xr = (o + d * tt).reshape(-1, 3).to(dev)
topo = (torch.randn(M, 2, generator=g) * 0.1).to(dev)
noise = torch.randn(M, 3, generator=g).to(dev)
_lib.PROFILE.enabled = True
for pts, tag in ((x, 'random points'), (xr[:M], 'ray-ordered points')):
    buf = (C.c_ulonglong * 16)()
    _lib.lib().mb_debug_fdr_phases(buf, 1)
    for it in range(5):
        if it == 2:
            _lib.PROFILE.reset()
        xx = pts.clone().requires_grad_(True)
        tp = topo.clone().requires_grad_(True)
        l, n, r = m.fd_regulariser(xx, tp, noise, 0.005, 1.0 / (3 * M))
        torch.cuda.synchronize()
    for k, v in _lib.PROFILE.summary().items():
        print(f'{tag}: {k}: {v["avg_ms"]:.3f} ms')
    _lib.check(_lib.lib().mb_debug_fdr_phases(buf, 1), 'debug_fdr_phases')
    names = ['tile prologue', 'row setup+bar', 'gather', 'freq enc', 'wait fwd0', 'epi A1', 'wait fwd1', 'pass1+bar', 'normals+bar', 'pass2 dZ1',
             'wait bwd1', 'epi dZ0', 'wait bwd0', 'epi dS0+bar', 'scatter+bar', 'fold']
    tot = sum(buf[:16])
    if tot:
        print(f'phases ({tag}; cycles of worker thread 0, all CTAs):')
        for nme, v in zip(names, buf[:16]):
            print(f'  {nme:16s} {100.0 * v / tot:5.1f}%')

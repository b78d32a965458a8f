"""CUDA-event timing of the fused field kernels on a fixed synthetic query set:  python tools/time_field.py [fwd|aux|bwd] [M]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from morpheus_b200 import _lib  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else 'fwd'
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4096 * 128
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
t = torch.full((M, 1), 0.3, device=dev)
light = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1).to(dev)
_lib.PROFILE.enabled = True
for it in range(6):
    if it == 2:
        _lib.PROFILE.reset()
    if mode == 'fwd':
        with torch.no_grad():
            out = m(x, t, light, ratio=1.0, shading='albedo_normal')
    elif mode == 'aux':
        n, raw = m.normal(x, topo=None)
        n.sum().backward()
    else:
        xg = x.clone().requires_grad_(True)
        out = m(xg, t, light, ratio=1.0, shading='albedo_normal')
        (out[0].sum() + out[2].sum() + out[3].sum()).backward()
    torch.cuda.synchronize()
for k, v in _lib.PROFILE.summary().items():
    print(f'{mode} {k}: {v["avg_ms"]:.3f} ms')

if mode in ('aux', 'bwd'):
    import ctypes as C
    buf = (C.c_ulonglong * 16)()
    _lib.check(_lib.lib().mb_debug_fd_phases(buf, 1), 'debug_fd_phases')
    names = ['tile prologue', 'row setup+bar', 'gather+enc', 'wait fwd0', 'epi A1', 'wait fwd1', 'epi A2->dZ1', 'wait bwd1', 'epi dZ0', 'wait bwd0',
             'epi dS0+bar', 'scatter+bar', 'fold(last)', 'flush+outputs']
    tot = sum(buf[:14]) or 1
    print('FD kernel phases (cycles of worker thread 0, all CTAs, all launches of this process):')
    for n, v in zip(names, buf[:14]):
        print(f'  {n:16s} {100.0 * v / tot:5.1f}%')

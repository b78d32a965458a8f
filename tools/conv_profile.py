"""a few launches of the own convolution kernel for ncu (VAE-sized and UNet-sized layers):
   ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 4 -c 3 -o gpurun_out/r02_conv python tools/conv_profile.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morpheus_b200 import guidance  # noqa: E402

dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
cases = [(1, 128, 128, 256, 256), (2, 320, 320, 32, 32), (2, 1280, 1280, 8, 8)]     # VAE level 0, UNet level 1, UNet level 4
with guidance._precision('fp32'), torch.no_grad():
    for rep in range(2):
        for (B, Ci, Co, H, W) in cases:
            x = torch.randn(B, Ci, H, W, generator=g).to(dev)
            sd = {'c.weight': (torch.randn(Co, Ci, 3, 3, generator=g) / (Ci * 9) ** 0.5).to(dev), 'c.bias': torch.zeros(Co, device=dev)}
            y = guidance._conv(x, sd, 'c', pre_silu=True)
            torch.cuda.synchronize()
print('done')

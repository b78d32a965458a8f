"""Golden vectors for the SDS networks from the UNMODIFIED reference classes
(ldm.modules.diffusionmodules.openaimodel.UNetModel, ldm.modules.diffusionmodules.model.Encoder) on CPU with seeded
random weights (tests/ldm_util.py).  Build container only.  Stubs needed to import ldm offline: matplotlib, omegaconf."""
import json
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))
sys.path.insert(0, '/root/reference')
for name in ('matplotlib', 'matplotlib.pyplot', 'omegaconf', 'omegaconf.listconfig'):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules['omegaconf.listconfig'].ListConfig = type('ListConfig', (list,), {})
from ldm.modules.diffusionmodules.model import Encoder  # noqa: E402
from ldm.modules.diffusionmodules.openaimodel import UNetModel  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

unet = UNetModel(image_size=32, in_channels=8, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
                 channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True, transformer_depth=1, context_dim=768,
                 use_checkpoint=False, legacy=False).eval()
enc = Encoder(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4], num_res_blocks=2,
              attn_resolutions=[], dropout=0.0).eval()
table = {'unet': {k: list(v.shape) for k, v in unet.state_dict().items()}, 'encoder': {k: list(v.shape) for k, v in enc.state_dict().items()},
         'quant_conv': {'weight': [8, 8, 1, 1], 'bias': [8]}, 'cc_projection': {'weight': [768, 772], 'bias': [768]}}
json.dump(table, open(os.path.join(OUT, 'ldm_keys.json'), 'w'))
from ldm_util import seeded_state  # noqa: E402

unet.load_state_dict(seeded_state(table['unet'], 1))
enc.load_state_dict(seeded_state(table['encoder'], 2))
qc = seeded_state(table['quant_conv'], 3)
g = torch.Generator().manual_seed(0)
x = torch.randn(2, 8, 32, 32, generator=g)
t = torch.tensor([260, 260])
ctx = torch.randn(2, 1, 768, generator=g)
ctx[0] = 0                                                     # unconditional half of the CFG batch
img = torch.rand(1, 3, 256, 256, generator=g) * 2 - 1
with torch.no_grad():
    eps = unet(x, t, ctx)
imgg = img.clone().requires_grad_(True)
moments = torch.nn.functional.conv2d(enc(imgg), qc['weight'], qc['bias'])
wv = torch.randn(moments.shape, generator=g)
(moments * wv).sum().backward()
np.savez_compressed(os.path.join(OUT, 'sds_nets.npz'), x=x.numpy(), t=t.numpy(), ctx=ctx.numpy(), eps=eps.numpy(), img=img.numpy(),
                    moments=moments.detach().numpy(), wv=wv.numpy(), g_img=imgg.grad.numpy())
print('saved; eps std', float(eps.std()), 'moments std', float(moments.std()))

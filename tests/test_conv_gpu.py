"""GPU: the tensor-core convolution of the SDS networks (csrc/conv_tc.cu through morpheus_b200.guidance._conv) against torch's float64
convolution on the same inputs -- forward (with and without the fused SiLU), the input-gradient backward (transposed pack), 3x3 and 1x1,
both N tiles (128 / 160), split-K, partial last tile, tiles that straddle two images.  Bar: 3-term fp16 split => fp32-grade accuracy
(rel-L2 <= 2e-6); cuDNN's own fp32 result is measured next to it."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


CASES = [  # B, Cin, Cout, H, W, k
    (2, 320, 320, 32, 32, 3),      # UNet level 1: N tile 160, split-K
    (1, 128, 128, 64, 64, 3),      # VAE-like: N tile 128
    (2, 1280, 1280, 8, 8, 3),      # deep UNet level: one M tile holding both images, 180 K stages split over the grid
    (2, 640, 320, 16, 16, 1),      # 1x1 skip connection
    (1, 64, 128, 20, 20, 3),       # partial last tile (400 pixels)
    (2, 256, 512, 12, 12, 3),      # tiles straddle the two images (144 pixels per image)
]


@pytest.mark.parametrize('case', CASES)
def test_conv_tc_vs_float64(case):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from morpheus_b200 import guidance
    B, Cin, Cout, H, W, k = case
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(B * 1000 + Cin + Cout + H)
    x = torch.randn(B, Cin, H, W, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    sd = {'c.weight': w, 'c.bias': b}
    pad = 1 if k == 3 else 0
    assert guidance.OWN_CONV
    with guidance._precision('fp32'):
        assert guidance._own_conv_ok(x, w, 1, pad)
        with torch.no_grad():
            y = guidance._conv(x, sd, 'c', padding=pad)
            y_silu = guidance._conv(x, sd, 'c', padding=pad, pre_silu=True)
        xg = x.clone().requires_grad_(True)
        yg = guidance._conv(xg, sd, 'c', padding=pad, pre_silu=True)
        gout = torch.randn(yg.shape, generator=g).to(dev) * 1e-6        # VAE-backward-sized gradients: exercises the dynamic power-of-two scale
        yg.backward(gout)
        y_cudnn = F.conv2d(x, w, b, padding=pad)
    xd = x.double().requires_grad_(True)
    ref = F.conv2d(xd, w.double(), b.double(), padding=pad)
    ref_silu = F.conv2d(F.silu(xd), w.double(), b.double(), padding=pad)
    ref_silu.backward(gout.double())
    e = rel(y, ref)
    assert e < 3e-6, e          # (K up to 11 520 with split-K atomics; cuDNN's fp32 result sits at 2.6e-6 on the 320-channel case)
    assert rel(y_silu, ref_silu) < 2e-6
    assert rel(yg, ref_silu) < 2e-6
    assert rel(xg.grad, xd.grad) < 5e-6, rel(xg.grad, xd.grad)      # (includes the fp32 SiLU backward of torch)
    print(f'conv {case}: own {e:.1e}, cuDNN fp32 {rel(y_cudnn, ref):.1e}')


def test_sds_chain_with_own_convs_matches_cudnn_path():
    """the whole SDS step (VAE forward + input-gradient backward, UNet) with the own convolutions against the same step on cuDNN fp32"""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from ldm_util import load_key_table, seeded_state
    from morpheus_b200 import guidance
    dev = torch.device('cuda:0')
    table = load_key_table()
    sd = {}
    sd.update(seeded_state(table['unet'], 1, 'model.diffusion_model.'))
    sd.update(seeded_state(table['encoder'], 2, 'first_stage_model.encoder.'))
    sd.update(seeded_state(table['quant_conv'], 3, 'first_stage_model.quant_conv.'))
    sd.update(seeded_state(table['cc_projection'], 4, 'cc_projection.'))
    z = guidance.Zero123(dev, state_dict=sd, t_range=[0.02, 0.5])
    g = torch.Generator().manual_seed(3)
    emb = {'c_crossattn': [torch.randn(1, 1, 768, generator=g).to(dev)], 'c_concat': [torch.randn(1, 4, 32, 32, generator=g).to(dev)],
           'ref_radii': [2.5], 'ref_polars': [90.0], 'ref_azimuths': [0.0], 'zero123_ws': [1]}
    pred = torch.rand(1, 3, 72, 72, generator=g).to(dev)
    args = dict(guidance_scale=5, grad_scale=0.01, t=torch.tensor([260], device=dev), noise=torch.randn(1, 4, 32, 32, generator=g).to(dev),
                vae_noise=torch.randn(1, 4, 32, 32, generator=g).to(dev))
    res = {}
    for own in (True, False):
        guidance.OWN_CONV = own
        try:
            pg = pred.clone().requires_grad_(True)
            loss, _, _, _ = z.train_step(emb, pg, torch.tensor([10.0]), torch.tensor([200.0]), torch.tensor([0.1]), **args)
            with guidance._precision('fp32'):
                loss.backward()
            res[own] = (float(loss), pg.grad.clone())
        finally:
            guidance.OWN_CONV = True
    assert abs(res[True][0] - res[False][0]) <= 1e-4 * abs(res[False][0])
    assert rel(res[True][1], res[False][1]) < 1e-4, rel(res[True][1], res[False][1])


@pytest.mark.parametrize('shape', [(2, 1024, 320, 320), (2, 1024, 320, 2560), (2, 256, 2560, 640), (2, 64, 1280, 1280), (1, 200, 64, 128)])
def test_linear_tc_vs_float64(shape):
    """token-major linear layers of the UNet transformer blocks through the same kernel (1x1 convolution over tokens, row-major output)"""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from morpheus_b200 import guidance
    Bt, T, K, N = shape
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(K + N + T)
    x = torch.randn(Bt, T, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    sd = {'l.weight': w, 'l.bias': b, 'n.weight': w}
    with guidance._precision('fp32'), torch.no_grad():
        y = guidance._lin(x, sd, 'l')
        y_nobias = guidance._lin(x, sd, 'n')
    ref = F.linear(x.double(), w.double(), b.double())
    assert y.shape == ref.shape
    assert rel(y, ref) < 2e-6, rel(y, ref)
    assert rel(y_nobias, F.linear(x.double(), w.double())) < 2e-6

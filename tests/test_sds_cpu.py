"""CPU: the functional latent-diffusion networks of morpheus_b200.guidance (they are plain torch, device agnostic)
against golden vectors from the UNMODIFIED reference classes UNetModel / Encoder with seeded weights
(tests/golden/make_sds_golden.py), and the schedule / pose-token / loss algebra against the oracle."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ldm_util import load_key_table, seeded_state  # noqa: E402


@pytest.fixture(scope='module')
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, 'sds_nets.npz'))


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def test_unet_forward_matches_reference_class(golden):
    from morpheus_b200.guidance import _KeyIndex, unet_forward
    table = load_key_table()
    sd = _KeyIndex(seeded_state(table['unet'], 1))
    with torch.no_grad():
        eps = unet_forward(sd, torch.from_numpy(golden['x']), torch.from_numpy(golden['t']), torch.from_numpy(golden['ctx']))
    assert eps.shape == (2, 4, 32, 32)
    assert rel(eps.numpy(), golden['eps']) < 1e-4          # north_star: SDS gradient within 1e-3


def test_vae_encoder_forward_and_input_grad_match_reference_class(golden):
    from morpheus_b200.guidance import _KeyIndex, vae_encode_moments
    table = load_key_table()
    sd = {('encoder.' + k): v for k, v in seeded_state(table['encoder'], 2).items()}
    sd.update({('quant_conv.' + k): v for k, v in seeded_state(table['quant_conv'], 3).items()})
    sd = _KeyIndex({k: v.requires_grad_(False) for k, v in sd.items()})
    img = torch.from_numpy(golden['img']).requires_grad_(True)
    m = vae_encode_moments(sd, img)
    (m * torch.from_numpy(golden['wv'])).sum().backward()
    assert rel(m.detach().numpy(), golden['moments']) < 1e-4
    assert rel(img.grad.numpy(), golden['g_img']) < 1e-3


def test_schedule_pose_token_and_loss_algebra():
    from morpheus_b200 import guidance
    from oracle import sds as osds
    ac = guidance.alphas_cumprod()
    torch.testing.assert_close(ac, osds.alphas_cumprod())
    assert abs(float(ac[0]) - (1 - 0.00085)) < 1e-6 and 0.0046 < float(ac[-1]) < 0.0048      # SD-1.x scaled-linear schedule
    # angle_between: vectorised vs the reference's double loop
    g = torch.Generator().manual_seed(0)
    s1 = torch.stack([2 + torch.rand(3, generator=g), torch.rand(3, generator=g) * 3, torch.rand(3, generator=g) * 6 - 3], -1)
    s2 = torch.stack([2 + torch.rand(4, generator=g), torch.rand(4, generator=g) * 3, torch.rand(4, generator=g) * 6 - 3], -1)
    torch.testing.assert_close(torch.rad2deg(guidance.Zero123.angle_between(s1, s2)), osds.angle_between_deg(s1, s2), rtol=1e-4, atol=1e-3)
    # d loss / d latents == grad exactly (zero123_utils.py:233-234)
    z = torch.randn(1, 4, 32, 32, generator=g, requires_grad=True)
    grad = torch.randn(1, 4, 32, 32, generator=g)
    osds.sds_loss(z, grad).backward()
    torch.testing.assert_close(z.grad, grad)

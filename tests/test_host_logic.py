"""CPU tests of the host-side training utilities that mirror the reference trainer:
learning-rate schedule (morpheus.py:471-502) and EMA (torch_ema semantics, morpheus.py:160-162,1432-1433)."""
import math

import torch

from morpheus_b200 import train as mtrain


def test_learning_factor_matches_reference_formula():
    W, E = 200, 2000                      # configs/snoopy.yaml warm_up_end / n_epochs
    assert mtrain.learning_factor(0, W, E) == 0.01
    assert mtrain.learning_factor(99, W, E) == 0.01
    assert abs(mtrain.learning_factor(100, W, E) - 0.01) < 1e-12
    assert abs(mtrain.learning_factor(150, W, E) - (0.01 + 0.5 * 0.99)) < 1e-12
    assert abs(mtrain.learning_factor(200, W, E) - 1.0) < 1e-12
    assert abs(mtrain.learning_factor(1100, W, E) - ((math.cos(math.pi * 0.5) + 1) * 0.5 * 0.95 + 0.05)) < 1e-12
    assert abs(mtrain.learning_factor(2000, W, E) - 0.05) < 1e-12
    assert abs(mtrain.learning_factor(300, W, E, scale_factor=0.5) - 0.5 * mtrain.learning_factor(300, W, E)) < 1e-12


def test_flat_ema_matches_torch_ema_recurrence():
    g = torch.Generator().manual_seed(0)
    flat = torch.randn(1000, generator=g)
    ema = mtrain.FlatEMA(flat, decay=0.95)
    shadow = flat.clone()
    for n in range(1, 40):
        flat.add_(torch.randn(1000, generator=g) * 0.1)
        ema.update()
        decay = min(0.95, (1 + n) / (10 + n))
        shadow.sub_((shadow - flat) * (1 - decay))      # torch_ema: s_param.sub_((s_param - param) * one_minus_decay)
        assert torch.allclose(ema.shadow, shadow, rtol=1e-6, atol=1e-7)
    before = flat.clone()
    ema.store()
    ema.copy_to()
    assert torch.equal(flat, ema.shadow)
    ema.restore()
    assert torch.equal(flat, before)

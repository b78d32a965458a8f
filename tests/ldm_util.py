"""Seeded random weights for the latent-diffusion networks (the Zero-1-to-3 checkpoint cannot be downloaded here):
the same procedure fills the reference classes (tests/golden/make_sds_golden.py, build container) and the product's
functional networks (GPU box), from the committed key->shape table tests/golden/ldm_keys.json."""
import json
import math
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def load_key_table():
    return json.load(open(os.path.join(HERE, 'golden', 'ldm_keys.json')))


def seeded_state(table, seed, prefix=''):
    """deterministic per-key init: matrices/convs ~ N(0, gain/fan_in), norm scales ~ 1 + 0.1 N, biases ~ 0.05 N"""
    sd = {}
    for idx, key in enumerate(sorted(table)):
        shape = table[key]
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        if len(shape) >= 2:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            w = torch.randn(shape, generator=g) * math.sqrt(1.5 / fan_in)
        elif key.endswith('weight'):
            w = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            w = 0.05 * torch.randn(shape, generator=g)
        sd[prefix + key] = w
    return sd

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


def test_per_iteration_host_choices_vs_reference_golden():
    """get_shading / get_bg_color / progressive_level / the ground-truth background blend (morpheus.py:808-813, 865-903, 929-944)
    executed from the reference source with seeded python / torch RNGs (tests/golden/make_loss_golden.py) vs the mirrors in
    morpheus_b200.train driven by the same seeds."""
    import os
    import random
    import numpy as np
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'host_choices.npz'))
    tr = dict(mtrain.DEFAULT_TRAIN_CFG)
    for k in ('albedo_iter_ratio', 'min_ambient_ratio', 'textureless_ratio'):
        assert abs(tr[k] - float(z[k])) < 1e-12
    names = {0: 'albedo', 1: 'albedo_normal', 2: 'lambertian', 3: 'textureless'}
    random.seed(5)
    for ratio, rv, amb, sid in z['shade']:
        a, s = mtrain.get_shading(float(ratio), bool(rv), tr)
        assert s == names[int(sid)] and abs(a - float(amb)) < 1e-12
    for ratio, lvl in z['levels']:
        assert abs(mtrain.progressive_level(float(ratio)) - float(lvl)) < 1e-12
    random.seed(6)
    torch.manual_seed(6)
    bg = mtrain.get_bg_color(True, 1, 9, 'cpu')
    assert torch.equal(bg, torch.from_numpy(z['bg_real']))
    for is_none, ref in zip(z['bg_virtual_none'], z['bg_virtual']):
        b = mtrain.get_bg_color(False, None, None, 'cpu')
        assert (b is None) == bool(is_none)
        if b is not None:
            assert torch.equal(b, torch.from_numpy(ref))
    img = torch.from_numpy(z['img'])[0, :, :, 0].t()          # [N,3]
    gt, m = mtrain.blend_gt_background(img, torch.from_numpy(z['mask']).reshape(-1), bg)
    assert torch.allclose(gt, torch.from_numpy(z['gt_rgb'])[0, :, :, 0].t(), rtol=0, atol=1e-7)
    assert torch.equal(m, torch.from_numpy(z['gt_mask']).reshape(-1))


def test_virtual_view_regulariser_assembly_vs_reference_source_golden():
    """train.virtual_view_loss_terms (incl. normal_smoothness * normal_reg, which the reference does NOT gate on real_view) against
    MorpheuS.get_regularization_loss executed from the reference source with the shipped weights (tests/golden/make_virtual_reg_golden.py)."""
    import ast
    import os
    import types

    import numpy as np
    import torch
    from morpheus_b200 import train as mtrain
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'virtual_view_reg_loss.npz'))
    weights = ast.literal_eval(str(z['weights'][0]))
    tr = mtrain.FULL_TRAIN_CFG
    for k in ('ori_weight', 'normal_smooth_3d', 'normal_smoothness', 'code_reg', 'beta_weight'):
        assert abs(tr[k] - weights[k]) < 1e-12, k              # the weights the step uses are the shipped yaml's
    for c in (ast.literal_eval(str(s)) for s in z['cases']):
        out = {k: torch.tensor(c[k]) for k in ('loss_orient', 'loss_normal_perturb', 'normal_reg', 'loss_code') if k in c}
        model = types.SimpleNamespace(sdf2density=types.SimpleNamespace(get_beta=lambda b=c['beta']: torch.tensor(b)))
        ours = float(mtrain.virtual_view_loss_terms(out, model, tr))
        assert abs(ours - c['total']) <= 1e-6 * abs(c['total']), (ours, c)


def test_flat_ema_invalidates_the_packed_arena_and_adam_active_groups():
    """FlatEMA.copy_to / restore rewrite the flat parameter buffer behind the parameters' version counters: the model's packed-arena cache
    must be dropped; FlatAdam.set_active mirrors torch.optim.Adam's skip of parameters without a gradient."""
    import torch
    from morpheus_b200 import train as mtrain

    class M:
        def __init__(self):
            self.invalidated = 0

        def invalidate(self):
            self.invalidated += 1
    m = M()
    flat = torch.arange(8, dtype=torch.float32)
    ema = mtrain.FlatEMA(flat, 0.95, model=m)
    flat += 1.0
    ema.update()
    ema.store()
    ema.copy_to()
    assert m.invalidated == 1 and torch.allclose(flat, ema.shadow)
    ema.restore()
    assert m.invalidated == 2 and torch.allclose(flat, torch.arange(8, dtype=torch.float32) + 1.0)
    # active-group logic (host side of mb_adam_step_groups)
    names = ['encoder_sdf', 'encoder_color', 'decoder_sdf', 'decoder_topo', 'decoder_color', 'density', 'decoder_deform', 'code_deform', 'pose', 'decoder_bg']
    fake = types_ns(group_names=names, NEVER=mtrain.FlatAdam.NEVER, _active_host=None, group_active=torch.ones(len(names), dtype=torch.uint8))
    mtrain.FlatAdam.set_active(fake, real_view=True)
    assert dict(zip(names, fake._active_host)) == {n: (0 if n == 'decoder_bg' else 1) for n in names}
    mtrain.FlatAdam.set_active(fake, real_view=False, shading='textureless')
    off = {n for n, a in zip(names, fake._active_host) if a == 0}
    assert off == {'decoder_bg', 'pose', 'encoder_color', 'decoder_color'}
    assert fake.group_active.tolist() == list(fake._active_host)


def types_ns(**kw):
    import types
    return types.SimpleNamespace(**kw)


def test_weighted_sum_matches_eager_assembly():
    import torch
    from morpheus_b200 import train as mtrain
    ts = [torch.tensor(float(i + 1), requires_grad=True) for i in range(5)]
    ws = [1.0, 10.0, 0.0, 0.1, 0.5]
    out = mtrain.weighted_sum(ts, ws)
    out.backward()
    assert abs(float(out) - sum(w * float(t) for t, w in zip(ts, ws))) < 1e-6
    for t, w in zip(ts, ws):
        assert (t.grad is None and w == 0.0) or abs(float(t.grad) - w) < 1e-7

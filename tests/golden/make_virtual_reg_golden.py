"""Golden for the regulariser assembly of a VIRTUAL-view step: `MorpheuS.get_regularization_loss` (/root/reference/morpheus.py:1090-1145)
executed from the reference's own source text (morpheus.py cannot be imported offline: nerfacc / Zero-1-to-3 stack) on a stand-in `self`
carrying the SHIPPED weights (configs/snoopy.yaml `train` section).  The outputs dict is what render_rays produces on a virtual view after
the first 10 % of training: loss_orient, loss_normal_perturb, normal_reg (NOT gated on real_view, :778-785, :1127-1128), loss_code.
    python tests/golden/make_virtual_reg_golden.py     (this container only: needs /root/reference)"""
import ast
import os
import types

import numpy as np
import torch
import yaml

SRC = '/root/reference/morpheus.py'
OUT = os.path.dirname(os.path.abspath(__file__))
tree = ast.parse(open(SRC).read())
cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'MorpheuS')
fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == 'get_regularization_loss')
env = {'torch': torch}
exec(compile(ast.Module(body=[fn], type_ignores=[]), 'morpheus.get_regularization_loss', 'exec'), env)
ytrain = yaml.safe_load(open('/root/reference/configs/snoopy.yaml'))['train']
g = torch.Generator().manual_seed(77)
cases = []
for beta_p, with_orient in ((0.1, True), (-0.03, True), (0.2, False)):
    outs = {'loss_normal_perturb': torch.rand((), generator=g), 'loss_code': torch.rand((), generator=g), 'normal_reg': torch.rand((), generator=g),
            'normal_raw': torch.randn(40, 3, generator=g), 'deform': torch.randn(40, 3, generator=g), 'weights': torch.rand(40, generator=g)}
    if with_orient:
        outs['loss_orient'] = torch.rand((), generator=g) * 30
    beta = torch.tensor(beta_p).abs() + 1e-4
    fake = types.SimpleNamespace(config={'train': ytrain, 'exp': {'end_iter': 1}}, global_step=0,
                                 model=types.SimpleNamespace(sdf2density=types.SimpleNamespace(get_beta=lambda b=beta: b)))
    with torch.no_grad():
        total = env['get_regularization_loss'](fake, outs, None, cano=False)
    cases.append({'beta': float(beta), 'total': float(total), **{k: float(v) for k, v in outs.items() if v.dim() == 0}})
np.savez_compressed(os.path.join(OUT, 'virtual_view_reg_loss.npz'), cases=np.array([str(c) for c in cases]),
                    weights=np.array([str({k: float(v) for k, v in ytrain.items() if isinstance(v, (int, float)) and not isinstance(v, bool)})]))
print(cases)

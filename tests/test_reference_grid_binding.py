"""INTEGRATION.md section 1, exercised against the reference's OWN module: /root/reference/external/encoders/gridencoder/grid.py is loaded
unchanged with its `_gridencoder` import satisfied by a recording stand-in, driven through `GridEncoder.forward` / backward on the CPU,
and every call it makes into `_backend` must bind -- positionally, argument for argument -- to `morpheus_b200.gridencoder._backend`
(the shim a maintainer swaps in).  Also pins the host-side table sizing (grid.py:125-141: offsets, embedding count, per-level scale)
of our `GridEncoder` to the reference class for the shipped constructor arguments (models/model.py:144-157).

Runs only where the reference tree exists (this container); the GPU box has no /root/reference (the kernels themselves are checked
there against oracle/_ref, the reference kernel compiled from the same sources)."""
import importlib.util
import inspect
import os
import sys
import types

import pytest
import torch

REF = '/root/reference/external/encoders/gridencoder/grid.py'
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason='reference tree not present')


class _Recorder(types.ModuleType):
    def __init__(self):
        super().__init__('_gridencoder')
        self.calls = []

    def grid_encode_forward(self, *args):
        self.calls.append(('grid_encode_forward', args))
        args[3].fill_(0.25)                  # outputs [L, B, C]
        if args[11] is not None:
            args[11].fill_(0.5)              # dy_dx

    def grid_encode_backward(self, *args):
        self.calls.append(('grid_encode_backward', args))


def load_reference_grid(recorder):
    sys.modules['_gridencoder'] = recorder
    try:
        spec = importlib.util.spec_from_file_location('ref_grid_module', REF)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.modules.pop('_gridencoder', None)
    return mod


def test_reference_call_sites_bind_to_our_backend():
    from morpheus_b200 import gridencoder as ours
    rec = _Recorder()
    ref = load_reference_grid(rec)
    enc = ref.GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=15, desired_resolution=128)
    x = (torch.rand(10, 3) * 2 - 1).requires_grad_(True)
    out = enc(x, bound=1.0, max_level=0.77)
    assert out.shape == (10, 32) and float(out[0, 0]) == 0.25
    out.sum().backward()
    names = [c[0] for c in rec.calls]
    assert names == ['grid_encode_forward', 'grid_encode_backward']
    for name, args in rec.calls:
        fn = getattr(ours._Backend, name)
        bound = inspect.signature(fn).bind(*args)           # raises TypeError on any arity mismatch
        assert list(bound.arguments) == list(inspect.signature(fn).parameters)      # every parameter filled, none left to defaults
    f = dict(zip(inspect.signature(ours._Backend.grid_encode_forward).parameters, rec.calls[0][1]))
    assert (f['B'], f['D'], f['C'], f['L']) == (10, 3, 2, 16) and f['max_level'] == 13 and f['H'] == 16
    assert f['gridtype'] == 0 and f['align_corners'] is False and f['interp'] == 0
    assert f['outputs'].shape == (16, 10, 2) and f['dy_dx'].shape == (10, 96)
    assert f['offsets'].dtype == torch.int32 and f['inputs'].dtype == torch.float32
    b = dict(zip(inspect.signature(ours._Backend.grid_encode_backward).parameters, rec.calls[1][1]))
    assert b['grad'].shape == (16, 10, 2) and b['grad'].is_contiguous() and b['grad_embeddings'].shape == enc.embeddings.shape
    assert b['grad_inputs'].shape == (10, 3) and float(b['grad_embeddings'].abs().max()) == 0.0 and float(b['grad_inputs'].abs().max()) == 0.0


@pytest.mark.parametrize('kw', [dict(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=15, desired_resolution=128),
                                dict(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048),
                                dict(input_dim=2, num_levels=8, level_dim=4, base_resolution=8, log2_hashmap_size=12, desired_resolution=256)])
def test_table_sizing_matches_reference_class(kw):
    from morpheus_b200 import gridencoder as ours
    ref = load_reference_grid(_Recorder())
    r = ref.GridEncoder(**kw)
    o = ours.GridEncoder(**kw)
    assert torch.equal(r.offsets.cpu(), o.offsets.cpu())
    assert r.embeddings.shape == o.embeddings.shape and r.output_dim == o.output_dim
    assert abs(float(r.per_level_scale) - float(o.per_level_scale)) < 1e-12
    assert set(r.state_dict().keys()) == set(o.state_dict().keys())

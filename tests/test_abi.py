"""CPU: the C-ABI library builds, loads and exports every symbol include/morpheus_b200.h declares
(no compute calls: there is no GPU here), and argument errors come back as codes + messages."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    g.build()
    from morpheus_b200 import _lib
    return _lib.lib()


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'morpheus_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(mb_[a-z0-9_]+)\s*\(', src)))


def test_header_and_library_agree(lib):
    from morpheus_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/morpheus_b200.h but not exported'
    assert sorted(_lib.SYMBOLS) == syms, 'morpheus_b200._lib.SYMBOLS out of sync with the header'


def test_version_and_error_reporting(lib):
    assert lib.mb_version() >= 100
    # null pointers are rejected before any CUDA call, with a message
    rc = lib.mb_grid_encode_forward(None, None, None, None, 8, 3, 2, 16, 16, C.c_float(0.2), 16, None, 0, 0, 0, 0, None)
    assert rc == -1
    assert b'null' in lib.mb_last_error()
    # B == 0 is a no-op (grid.py would launch a 0-block grid)
    assert lib.mb_composite_forward(None, 0, 0, None, None, None, None, None, None, None, None, None, None, None) == 0


def test_struct_layouts_match_header():
    """sizeof of the ctypes mirrors == what the C compiler lays out (checked by compiling a probe with gcc)."""
    import subprocess
    import tempfile
    from morpheus_b200 import _lib
    probe = '#include <stdio.h>\n#include "morpheus_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(mb_layer_desc), sizeof(mb_field_params), sizeof(mb_field_io), sizeof(mb_field_grads));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, 'p.c'), 'w').write(probe)
        subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), os.path.join(d, 'p.c'), '-o', os.path.join(d, 'p')], check=True)
        out = subprocess.run([os.path.join(d, 'p')], capture_output=True, text=True, check=True).stdout.split()
    sizes = [int(x) for x in out]
    assert sizes == [C.sizeof(_lib.LayerDesc), C.sizeof(_lib.FieldParams), C.sizeof(_lib.FieldIO), C.sizeof(_lib.FieldGrads)]


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under morpheus_b200/ may import or execute it"""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, 'morpheus_b200')):
        for f in fs:
            if not f.endswith(('.py', '.cu', '.cuh', '.h')):
                continue
            txt = open(os.path.join(dp, f)).read()
            if f.endswith('.py'):
                code = '\n'.join(l for l in txt.splitlines() if not l.lstrip().startswith('#'))
                code = re.sub(r'(\'\'\'|\"\"\").*?\1', '', code, flags=re.S)          # docstrings may cite the oracle
                if re.search(r'\boracle\b', code):
                    bad.append(os.path.join(dp, f))
            elif re.search(r'#\s*include[^\n]*oracle', txt):
                bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_missing_library_fails_loudly(monkeypatch):
    from morpheus_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libmorpheus_b200.so')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _lib.lib()


def test_packing_layout():
    import torch
    from morpheus_b200 import packing
    assert packing.ARENA_FLOATS == 364128
    rows = packing.LAYOUT['deform']
    assert rows[0][3:] == (87, 128, 96, 128) and rows[5][3:] == (128, 3, 128, 16)
    assert all(r[0] % 16 == 0 and r[1] % 16 == 0 and r[2] % 16 == 0 for net in packing.NET_ORDER for r in packing.LAYOUT[net])
    g = torch.Generator().manual_seed(0)
    layers = {net: [(torch.randn(N, K, generator=g, requires_grad=True), torch.randn(N, generator=g, requires_grad=True))
                    for (K, N) in packing.NET_DIMS[net]] for net in packing.NET_ORDER}
    arena = packing.pack(layers)
    assert arena.numel() == packing.ARENA_FLOATS
    wt_off, w_off, b_off, K, N, Kp, Np = packing.LAYOUT['sdf'][2]
    W, b = layers['sdf'][2]
    assert torch.equal(arena[wt_off:wt_off + Kp * Np].view(Kp, Np)[:K, :N], W.t())
    assert torch.equal(arena[w_off:w_off + Kp * Np].view(Np, Kp)[:N, :K], W)
    assert torch.equal(arena[b_off:b_off + N], b)
    assert float(arena[wt_off:wt_off + Kp * Np].view(Kp, Np)[K:].abs().sum()) == 0
    # autograd maps the flat gradient arena back onto the layer tensors (Wt slot only carries dW)
    ga = torch.zeros_like(arena)
    ga[wt_off:wt_off + Kp * Np] = 1.0
    arena.backward(ga)
    assert torch.equal(W.grad, torch.ones_like(W))

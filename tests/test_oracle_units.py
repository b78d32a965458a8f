"""CPU: unit checks of the oracle restatements against closed forms / known constants of the reference."""
import numpy as np
import torch

from oracle import grid as og
from oracle import render as orr


def test_level_resolutions_and_table_sizes():
    """float32 resolution rule gives 16..128 with 32/64/128 at levels 5/10/15; tables sized by grid.py:125-136"""
    S = float(np.log2(og.per_level_scale()))
    res = [og.level_resolution(l, S, 16) for l in range(16)]
    assert res == [16, 19, 22, 25, 28, 32, 37, 43, 49, 56, 64, 74, 85, 98, 112, 128]
    off = og.make_offsets()
    sizes = np.diff(off).tolist()
    assert sizes[:5] == [4096, 6864, 10648, 15632, 21952] and all(s == 32768 for s in sizes[5:])
    assert int(off[-1]) == 419640


def test_hash_index_wraparound():
    """gridencoder.cu:46-58: uint32 multiply wraps; checked against python big-int arithmetic"""
    coords = [np.array([5, 127], dtype=np.uint32), np.array([77, 126], dtype=np.uint32), np.array([120, 3], dtype=np.uint32)]
    idx = og._grid_index(0, 32768, 128, coords)
    for i in range(2):
        x, y, z = (int(c[i]) for c in coords)
        h = ((x * 1) & 0xFFFFFFFF) ^ ((y * 2654435761) & 0xFFFFFFFF) ^ ((z * 805459861) & 0xFFFFFFFF)
        assert idx[i] == h % 32768
    small = [np.array([5], dtype=np.uint32), np.array([17], dtype=np.uint32), np.array([30], dtype=np.uint32)]
    dense = og._grid_index(0, 32768, 32, small)      # level 5: 32^3 == table size -> dense, no hash
    assert dense[0] == 5 + 17 * 32 + 30 * 1024


def test_grid_forward_is_trilinear_and_oob_zero():
    off = og.make_offsets()
    S = float(np.log2(og.per_level_scale()))
    rng = np.random.default_rng(0)
    emb = rng.uniform(-1, 1, size=(int(off[-1]), 2)).astype(np.float32)
    # at a cell centre of level 0 (res 16) the encoding equals the table entry itself
    ijk = np.array([[3, 7, 11]])
    x = ((ijk + 0.5) / 16).astype(np.float32)
    out, dy = og.grid_encode_forward(x, emb, off, 16, S, 16, True)
    np.testing.assert_allclose(out[0, 0], emb[3 + 7 * 16 + 11 * 256], rtol=1e-6)
    oob = np.array([[1.0000001, 0.5, 0.5], [0.5, -1e-7, 0.5]], dtype=np.float32)
    out, dy = og.grid_encode_forward(oob, emb, off, 16, S, 16, True)
    assert not out.any() and not dy.any()
    # partial levels: the rest stays zero (grid.py:53)
    out, _ = og.grid_encode_forward(x, emb, off, 9, S, 16, False)
    assert out[:9].any() and not out[9:].any()


def test_grid_dydx_matches_finite_differences():
    off = og.make_offsets()
    S = float(np.log2(og.per_level_scale()))
    rng = np.random.default_rng(1)
    emb = rng.uniform(-1, 1, size=(int(off[-1]), 2)).astype(np.float32)
    x = rng.uniform(0.2, 0.8, size=(6, 3)).astype(np.float32)
    _, dy = og.grid_encode_forward(x, emb, off, 4, S, 16, True)
    dy = dy.reshape(6, 16, 3, 2)
    h = 1e-3
    for d in range(3):
        e = np.zeros(3, dtype=np.float32); e[d] = h
        op, _ = og.grid_encode_forward(x + e, emb, off, 4, S, 16, False)
        om, _ = og.grid_encode_forward(x - e, emb, off, 4, S, 16, False)
        fd = (op - om)[:4] / (2 * h)       # [L,B,C]
        # piecewise-linear: matches unless the +-h step crosses a cell boundary
        ok = np.isclose(np.transpose(dy[:, :4, d], (1, 0, 2)), fd, rtol=5e-2, atol=5e-2)
        assert ok.mean() > 0.8


def test_grid_backward_is_adjoint_of_forward():
    off = og.make_offsets()
    S = float(np.log2(og.per_level_scale()))
    rng = np.random.default_rng(2)
    emb = rng.uniform(-1, 1, size=(int(off[-1]), 2)).astype(np.float32)
    x = rng.uniform(0, 1, size=(50, 3)).astype(np.float32)
    g = rng.normal(size=(16, 50, 2)).astype(np.float32)
    out, _ = og.grid_encode_forward(x, emb, off, 16, S, 16, False)
    ge, _ = og.grid_encode_backward(g, x, emb, off, 16, S, 16)
    # <g, F(emb)> == <F^T g, emb> because F is linear in the table
    np.testing.assert_allclose((g.astype(np.float64) * out).sum(), (ge.astype(np.float64) * emb).sum(), rtol=1e-4)


def test_compositing_closed_form_and_empty_rays():
    counts = torch.tensor([0, 3, 0, 2])
    ri = torch.arange(4).repeat_interleave(counts)
    t0 = torch.tensor([0.0, 0.1, 0.2, 1.0, 1.5])
    t1 = t0 + 0.1
    sig = torch.tensor([1.0, 2.0, 3.0, 10.0, 0.0])
    w, T, a = orr.render_weight_from_density(t0, t1, sig, ri, 4)
    np.testing.assert_allclose(T.numpy(), [1.0, np.exp(-0.1), np.exp(-0.3), 1.0, np.exp(-1.0)], rtol=1e-6)
    np.testing.assert_allclose(a.numpy(), 1 - np.exp(-sig.numpy() * 0.1), rtol=1e-6)
    op = orr.accumulate_along_rays(w, None, ri, 4)
    assert op[0, 0] == 0 and op[2, 0] == 0
    np.testing.assert_allclose(op[1, 0], 1 - np.exp(-0.6), rtol=1e-6)    # sum of weights telescopes to 1 - T_end


def test_sampler_invariants():
    """sorted/packed, step lattice, interval centre inside an occupied cell of the AABB"""
    g = torch.Generator().manual_seed(0)
    c2w = orr.look_at_pose(60.0, -40.0, 2.5)
    dirs = orr.camera_dirs(360, 360, 517.0, 517.0, 180.0, 180.0).reshape(-1, 3)
    idx = torch.randint(0, dirs.shape[0], (40,), generator=g)
    o, d = orr.rays_from_pose(dirs[idx], c2w)
    aabb = torch.tensor([-1.01, -1.01, -1.01, 1.01, 1.01, 1.01])
    r = torch.arange(32)
    cx, cy, cz = torch.meshgrid(r, r, r, indexing='ij')
    centre = (torch.stack([cx, cy, cz], -1).float() + 0.5) / 32 * 2.02 - 1.01
    binaries = centre.norm(dim=-1) < 0.5
    jit = torch.rand(40, generator=g)
    ri, t0, t1 = orr.sample_occgrid(o, d, binaries, aabb, 0.01, jit)
    assert ri.numel() > 0 and bool((ri[1:] >= ri[:-1]).all())
    np.testing.assert_allclose((t1 - t0).numpy(), 0.01, rtol=1e-4)
    mid = (t0 + t1) / 2
    p = o[ri] + d[ri] * mid[:, None]
    cell = torch.floor((p + 1.01) / 2.02 * 32).long().clamp(0, 31)
    assert bool(binaries[cell[:, 0], cell[:, 1], cell[:, 2]].all())
    tmin, _ = orr.ray_aabb(o, d, aabb)
    k = (t0 - (tmin[ri] + jit[ri] * 0.01)) / 0.01
    np.testing.assert_allclose(k.numpy(), np.round(k.numpy()), atol=2e-2)    # on the per-ray lattice


def test_sdf_loss_matches_reference_quirks():
    z = torch.tensor([[0.5], [0.95], [1.05], [2.0], [0.2]])
    tgt = torch.tensor([[1.0], [1.0], [1.0], [0.0], [-1.0]])
    sdf = torch.tensor([0.4, 0.06, -0.04, 0.3, 0.1])
    fs, sl = orr.get_sdf_loss(z, tgt, sdf, 0.1, mask=torch.ones(5, 1))
    # samples 1,2 in the band (|d - z| <= .1): |0.06-0.05| + |-0.04+0.05| = 0.02, each /(1+1e-8), / 4 non-zero depths
    np.testing.assert_allclose(float(sl), 0.02 / 4, rtol=1e-5)
    assert float(fs) >= 0

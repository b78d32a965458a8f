"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Every test drives the CUDA product path
through the C ABI (morpheus_b200/libmorpheus_b200.so) and compares it with
  * the CPU oracle (oracle/) on the same seeded inputs,
  * the golden vectors produced by the unmodified reference Python (tests/golden/scene_*.npz),
  * and, for the grid encoder, the reference CUDA kernel itself (oracle/_ref, built from
    /root/reference/external/encoders/gridencoder/src by oracle/build_ref.sh) -- bit for bit.
Tolerances: integer/index outputs bit-exact; fp32 outputs within the stated rel/abs bounds
(BASELINE.json north_star: rendered RGB/depth <= 1e-4 relative L2).
"""
import glob
import importlib.util
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIG = {'model': {'bg_radius': 1.4, 'activation': 'exp'}}


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def cpu(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')


def make_model(sd, max_level, dev):
    from morpheus_b200.model import scene_representation
    m = scene_representation(CONFIG, 1.01, num_frames=200, deform_dim=16, use_app=False, use_t=False, amb_dim=2,
                             color_grid=True, use_joint=True, encode_topo=False)
    m.load_state_dict(sd, strict=True)   # reference state_dict keys load unchanged
    m.max_level = max_level
    return m.to(dev)


def load_case(golden_dir, tag):
    from oracle.fields import init_reference_like_state
    z = np.load(os.path.join(golden_dir, f'scene_{tag}.npz'))
    ml = float(z['max_level'])
    sd = init_reference_like_state(200, seed=int(z['seed']), randomize=bool(z['randomize']), emb_scale=float(z['emb_scale']))
    return z, sd, (None if ml < 0 else ml)


# ------------------------------------------------------------------------------------------------
# grid encoder
# ------------------------------------------------------------------------------------------------
def _grid_inputs(B, seed, dev):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, generator=g)
    x[0] = torch.tensor([0.0, 0.0, 0.0]); x[1] = torch.tensor([1.0, 1.0, 1.0]); x[2] = torch.tensor([0.5, 0.5, 0.5])
    x[3] = torch.tensor([1.0001, 0.2, 0.3]); x[4] = torch.tensor([-1e-6, 0.2, 0.3])      # out of bounds
    x[5] = torch.tensor([31.5 / 32, 0.5 / 32, 16.5 / 32])                                  # exactly on cell centres
    x[6] = x[7] + torch.tensor([2e-3 / 2.02, 0, 0])                                        # FD-offset pair
    return x


@pytest.mark.parametrize('max_level_frac', [None, 0.5, 0.53])
@pytest.mark.parametrize('emb_scale', [1e-4, 1.0])
def test_grid_encode_vs_oracle(dev, max_level_frac, emb_scale):
    from morpheus_b200.gridencoder import GridEncoder
    from oracle import grid as og
    B = 1000
    enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=15, desired_resolution=128).to(dev)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        enc.embeddings.copy_(((torch.rand(enc.embeddings.shape, generator=g) * 2 - 1) * emb_scale).to(dev))
    x01 = _grid_inputs(B, 5, dev)
    xw = (x01 * 2.02 - 1.01).to(dev).requires_grad_(True)
    out = enc(xw, bound=1.01, max_level=max_level_frac)
    gout = torch.randn(out.shape, generator=g).to(dev)
    out.backward(gout)
    # oracle on the exact same [0,1] inputs the kernel saw
    u = cpu((xw.detach() + 1.01) / (2 * 1.01))
    S = float(np.log2(enc.per_level_scale))
    ml = og.resolve_max_level(max_level_frac, 16)
    o_ref, dydx_ref = og.grid_encode_forward(u, cpu(enc.embeddings), cpu(enc.offsets), ml, S, 16, True)
    o_ref = np.transpose(o_ref, (1, 0, 2)).reshape(B, 32)
    np.testing.assert_allclose(cpu(out), o_ref, rtol=1e-6, atol=1e-7 * emb_scale)
    g_lbc = np.ascontiguousarray(np.transpose(cpu(gout).reshape(B, 16, 2), (1, 0, 2)))
    ge_ref, gi_ref = og.grid_encode_backward(g_lbc, u, cpu(enc.embeddings), cpu(enc.offsets), ml, S, 16, dydx_ref)
    assert rel_l2(cpu(enc.embeddings.grad), ge_ref) < 1e-5
    assert rel_l2(cpu(xw.grad), gi_ref / 2.02) < 1e-5


def _load_ref_backend():
    paths = glob.glob(os.path.join(ROOT, 'oracle', '_ref', '_gridencoder_ref*.so'))
    if not paths:
        return None
    spec = importlib.util.spec_from_file_location('_gridencoder_ref', paths[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize('calc_dydx', [False, True])
def test_grid_encode_bit_exact_vs_reference_kernel(dev, calc_dydx):
    """Our kernel against the reference CUDA kernel compiled from the reference sources (oracle/_ref)."""
    ref = _load_ref_backend()
    if ref is None:
        pytest.skip('oracle/_ref not built (needs /root/reference at build time)')
    from morpheus_b200.gridencoder import _backend as ours
    from oracle import grid as og
    B, D, C, L, H = 20000, 3, 2, 16, 16
    S = float(np.log2(og.per_level_scale()))
    offsets = torch.from_numpy(og.make_offsets()).to(dev)
    g = torch.Generator().manual_seed(11)
    emb = ((torch.rand(int(offsets[-1]), C, generator=g) * 2 - 1)).to(dev)
    x = _grid_inputs(B, 9, dev).to(dev)
    for max_level in (16, 9):
        outs, dys = [], []
        for be in (ref, ours):
            o = torch.zeros(L, B, C, device=dev)
            dy = torch.zeros(B, L * D * C, device=dev) if calc_dydx else None
            be.grid_encode_forward(x, emb, offsets, o, B, D, C, L, max_level, S, H, dy, 0, False, 0)
            outs.append(o); dys.append(dy)
        torch.cuda.synchronize()
        assert torch.equal(outs[0], outs[1]), f'forward differs: max abs {(outs[0] - outs[1]).abs().max().item()}'
        if calc_dydx:
            assert torch.equal(dys[0], dys[1]), f'dy_dx differs: max abs {(dys[0] - dys[1]).abs().max().item()}'
            grad = torch.randn(L, B, C, generator=g).to(dev)
            res = []
            for be in (ref, ours):
                ge = torch.zeros_like(emb)
                gi = torch.zeros(B, D, device=dev)
                be.grid_encode_backward(grad, x, emb, offsets, ge, B, D, C, L, max_level, S, H, dys[0], gi, 0, False, 0)
                res.append((ge, gi))
            torch.cuda.synchronize()
            assert torch.equal(res[0][1], res[1][1]), 'grad_inputs differs'
            assert rel_l2(cpu(res[1][0]), cpu(res[0][0])) < 1e-6   # atomics: order differs


def test_grid_encode_errors(dev):
    from morpheus_b200.gridencoder import _backend as ours
    x = torch.rand(8, 5, device=dev)
    emb = torch.rand(64, 2, device=dev)
    off = torch.tensor([0, 64], dtype=torch.int32, device=dev)
    out = torch.zeros(1, 8, 2, device=dev)
    with pytest.raises(RuntimeError):   # reference: std::runtime_error "D must be 2, 3, 4 or 5" -> we support {2,3}
        ours.grid_encode_forward(x, emb, off, out, 8, 5, 2, 1, 1, 0.0, 16, None, 0, False, 0)
    with pytest.raises(RuntimeError):
        ours.grid_encode_forward(x.cpu(), emb, off, out, 8, 3, 2, 1, 1, 0.0, 16, None, 0, False, 0)


# ------------------------------------------------------------------------------------------------
# scene model: forward in every shading mode, vs golden (reference Python) and oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('tag', ['init_full', 'rand_c2f', 'rand_full'])
def test_scene_forward_vs_reference_golden(dev, golden_dir, tag):
    z, sd, ml = load_case(golden_dir, tag)
    m = make_model(sd, ml, dev).eval()
    x, t, light = (torch.from_numpy(z[k]).to(dev) for k in ('x', 't', 'light'))

    def close(a, name, rtol=1e-4, atol=1e-5):
        np.testing.assert_allclose(cpu(a), z[name], rtol=rtol, atol=atol, err_msg=name)
    with torch.no_grad():
        for shading, ratio in (('albedo', 1.0), ('albedo_normal', 1.0), ('lambertian', 0.3), ('textureless', 0.55), ('normal', 1.0)):
            sdf, sigma, color, normal, deform, raw = m(x, t, light, ratio=ratio, shading=shading)
            close(sdf, f'{shading}.sdf'); close(sigma, f'{shading}.sigma', rtol=5e-4, atol=1e-4); close(deform, f'{shading}.deform')
            shaded = shading in ('lambertian', 'textureless', 'normal')
            close(color, f'{shading}.color', rtol=2e-3 if shaded else 1e-4, atol=5e-4 if shaded else 1e-5)
            if normal is not None:
                close(raw, f'{shading}.normal_raw', rtol=2e-3, atol=5e-4)   # FD of fp32 values: cancellation noise
                close(normal, f'{shading}.normal', rtol=2e-3, atol=5e-4)
        d = m.density(x, t)
        close(d['sdf'], 'density.sdf'); close(d['albedo'], 'density.albedo')
        d = m.density(x, None)
        close(d['sdf'], 'density_cano.sdf'); close(d['albedo'], 'density_cano.albedo')
        d = m.density(x, t[:3], allow_shape=True, return_color=False)
        close(d['sigma'], 'density_allow_shape.sigma', rtol=5e-4, atol=1e-4)
        n, raw = m.normal(x, t=t)
        close(raw, 'normal_warped.raw', rtol=2e-3, atol=5e-4)
        n, raw = m.normal(x, topo=None)
        close(raw, 'normal_cano.raw', rtol=2e-3, atol=5e-4)
        deform, topo, _ = m.warp(x, t)
        close(deform, 'warp.deform'); close(topo, 'warp.topo')
        close(m.get_deform_code(t), 'code', rtol=1e-5, atol=1e-6)
        close(m.background(light, t), 'background')
        o2, d2 = m.pose_optimisation(x, light, torch.from_numpy(z['pose.ids']).to(dev))
        close(o2, 'pose.o', rtol=1e-5); close(d2, 'pose.d', rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('tag', ['init_full', 'rand_c2f', 'rand_full'])
def test_scene_backward_vs_reference_golden(dev, golden_dir, tag):
    """Param / input gradients of the fused backward kernel vs autograd through the unmodified reference."""
    z, sd, ml = load_case(golden_dir, tag)
    m = make_model(sd, ml, dev).train()
    x, t, light = (torch.from_numpy(z[k]).to(dev) for k in ('x', 't', 'light'))
    M = x.shape[0]
    xg = x.clone().requires_grad_(True)
    sdf, sigma, color, normal, deform, raw = m(xg, t, light, ratio=1.0, shading='albedo_normal')
    g = torch.Generator().manual_seed(77)
    r = [torch.randn(M, generator=g), torch.randn(M, generator=g), torch.randn(M, 3, generator=g), torch.randn(M, 3, generator=g),
         torch.randn(M, 3, generator=g)]
    r = [v.to(dev) for v in r]
    loss = (sdf * r[0]).sum() + (sigma * r[1]).sum() * 1e-2 + (color * r[2]).sum() + (normal * r[3]).sum() + (deform * r[4]).sum()
    loss.backward()
    bad = []
    errs = {'x': rel_l2(cpu(xg.grad), z['grad.x'])}
    for name, p in m.named_parameters():
        key = 'grad.' + name
        if key in z.files and p.grad is not None and np.abs(z[key]).max() > 0:
            errs[name] = rel_l2(cpu(p.grad), z[key])
    for k, e in errs.items():
        # beta: with |sdf| >> beta the reference's own fp32 `0.5 + 0.5*expm1(-|s|/beta)` cancels catastrophically (its
        # autograd value is ~1% off the fp64 value, see DESIGN.md "numerics"); the well-conditioned beta-gradient check
        # is test_scene_forward_backward_vs_oracle_large.
        if e > (5e-2 if k == 'sdf2density.beta' else 2e-3):
            bad.append((k, e))
    assert not bad, f'gradient mismatch: {bad}\nall: {errs}'


def test_scene_forward_backward_vs_oracle_large(dev):
    """Bigger ragged batch (partial last tile, two frames) against the CPU oracle incl. autograd."""
    from oracle.fields import SceneOracle, init_reference_like_state
    sd = init_reference_like_state(200, seed=5, randomize=True, emb_scale=0.3)
    sd['sdf2density.beta'] = torch.tensor(0.6)     # |sdf| ~ beta: well-conditioned sigma / d sigma / d beta
    M = 1000 + 37
    g = torch.Generator().manual_seed(123)
    x = (torch.rand(M, 3, generator=g) * 2 - 1) * 0.95
    t = torch.full((M, 1), 63.0 / 200)
    light = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1)
    w = [torch.randn(M, generator=g), torch.randn(M, 3, generator=g), torch.randn(M, 3, generator=g)]
    # oracle
    sdo = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    sc = SceneOracle(sdo, 1.01, 200, 0.77)
    xo = x.clone().requires_grad_(True)
    so, sgo, co, no, do, ro = sc.forward(xo, t, light, ratio=0.4, shading='lambertian')
    lo = (so * w[0]).sum() + (co * w[1]).sum() + (no * w[2]).sum() + (sgo * w[0]).sum()
    lo.backward()
    # ours
    m = make_model(sd, 0.77, dev).train()
    xd = x.to(dev).requires_grad_(True)
    s, sg, c, n, d, r = m(xd, t.to(dev), light.to(dev), ratio=0.4, shading='lambertian')
    l = (s * w[0].to(dev)).sum() + (c * w[1].to(dev)).sum() + (n * w[2].to(dev)).sum() + (sg * w[0].to(dev)).sum()
    l.backward()
    assert rel_l2(cpu(s), cpu(so)) < 1e-5
    assert rel_l2(cpu(sg), cpu(sgo)) < 1e-4
    assert rel_l2(cpu(c), cpu(co)) < 1e-3
    assert rel_l2(cpu(d), cpu(do)) < 1e-5
    errs = {'x': rel_l2(cpu(xd.grad), cpu(xo.grad))}
    for name, p in m.named_parameters():
        if name in sdo and sdo[name].grad is not None and p.grad is not None and float(sdo[name].grad.abs().max()) > 0:
            errs[name] = rel_l2(cpu(p.grad), cpu(sdo[name].grad))
    bad = {k: e for k, e in errs.items() if e > 5e-3}
    assert not bad, f'{bad}\nall: {errs}'


# ------------------------------------------------------------------------------------------------
# sampling + compositing
# ------------------------------------------------------------------------------------------------
def _rays(N, seed):
    from oracle.render import camera_dirs, look_at_pose, rays_from_pose
    g = torch.Generator().manual_seed(seed)
    c2w = look_at_pose(70.0, 35.0, 2.5)
    dirs = camera_dirs(360, 360, 517.0, 517.0, 180.0, 180.0).reshape(-1, 3)
    idx = torch.randint(0, dirs.shape[0], (N,), generator=g)
    o, d = rays_from_pose(dirs[idx], c2w)
    return o.contiguous(), d.contiguous(), g


def test_sampler_vs_oracle(dev):
    from morpheus_b200.nerfacc_compat import OccGridEstimator
    from oracle.render import sample_occgrid
    N = 64
    o, d, g = _rays(N, 1)
    aabb = torch.tensor([-1.01, -1.01, -1.01, 1.01, 1.01, 1.01])
    est = OccGridEstimator(aabb, resolution=128).to(dev)
    r = torch.arange(128)
    cx, cy, cz = torch.meshgrid(r, r, r, indexing='ij')
    centre = (torch.stack([cx, cy, cz], -1).float() + 0.5) / 128 * 2.02 - 1.01
    binaries = (centre.norm(dim=-1) < 0.6) & (centre.norm(dim=-1) > 0.3)     # spherical shell: ragged, some empty rays
    est.binaries = binaries[None].to(dev)
    jitter = torch.rand(N, generator=g)
    ri, t0, t1 = est.sampling(o.to(dev), d.to(dev), render_step_size=0.01, stratified=True, jitter=jitter.to(dev))
    ri_o, t0_o, t1_o = sample_occgrid(o, d, binaries, aabb, 0.01, jitter)
    assert ri.dtype == torch.int64
    assert torch.equal(ri.cpu(), ri_o)                       # index work: bit-exact
    np.testing.assert_array_equal(cpu(t0), t0_o.numpy())     # same float32 lattice
    np.testing.assert_array_equal(cpu(t1), t1_o.numpy())
    assert bool((ri[1:] >= ri[:-1]).all())


def test_composite_vs_oracle(dev):
    from morpheus_b200 import nerfacc_compat as nf
    from oracle import render as orr
    g = torch.Generator().manual_seed(2)
    counts = torch.tensor([0, 5, 1, 0, 77, 33, 0, 130, 2, 0])
    N = counts.numel()
    ri = torch.arange(N).repeat_interleave(counts)
    M = ri.numel()
    t0 = torch.rand(M, generator=g) * 3
    t1 = t0 + 0.01 + torch.rand(M, generator=g) * 0.02
    sig = (torch.rand(M, generator=g) * 30).requires_grad_(True)
    rgb = torch.rand(M, 3, generator=g).requires_grad_(True)
    gw, go, gd, gc = torch.randn(M, generator=g), torch.randn(N, 1, generator=g), torch.randn(N, 1, generator=g), torch.randn(N, 3, generator=g)
    w, _, _ = orr.render_weight_from_density(t0, t1, sig, ri, N)
    op = orr.accumulate_along_rays(w, None, ri, N)
    dp = orr.accumulate_along_rays(w, ((t0 + t1) / 2)[:, None], ri, N)
    cl = orr.accumulate_along_rays(w, rgb, ri, N)
    ((w * gw).sum() + (op * go).sum() + (dp * gd).sum() + (cl * gc).sum()).backward()
    sig_d = sig.detach().to(dev).requires_grad_(True)
    rgb_d = rgb.detach().to(dev).requires_grad_(True)
    w2, o2, d2, c2 = nf.composite(sig_d, rgb_d, t0.to(dev), t1.to(dev), ri.to(dev), N)
    ((w2 * gw.to(dev)).sum() + (o2 * go[:, 0].to(dev)).sum() + (d2 * gd[:, 0].to(dev)).sum() + (c2 * gc.to(dev)).sum()).backward()
    np.testing.assert_allclose(cpu(w2), cpu(w), rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(cpu(o2), cpu(op)[:, 0], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(cpu(d2), cpu(dp)[:, 0], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(cpu(c2), cpu(cl), rtol=2e-5, atol=1e-6)
    assert rel_l2(cpu(sig_d.grad), cpu(sig.grad)) < 1e-4
    assert rel_l2(cpu(rgb_d.grad), cpu(rgb.grad)) < 1e-5
    # API-compatible split calls
    w3, tr3, al3 = nf.render_weight_from_density(t0.to(dev), t1.to(dev), sig_d.detach(), ray_indices=ri.to(dev), n_rays=N)
    np.testing.assert_allclose(cpu(w3), cpu(w), rtol=2e-5, atol=1e-7)
    acc = nf.accumulate_along_rays(w3, values=rgb_d.detach(), ray_indices=ri.to(dev), n_rays=N)
    np.testing.assert_allclose(cpu(acc), cpu(cl), rtol=2e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------------
# end to end: render_rays (BASELINE cfg-1: 256 rays x 64 samples, SDS off) vs the oracle
# ------------------------------------------------------------------------------------------------
def test_render_rays_cfg1_vs_oracle(dev):
    from morpheus_b200.nerfacc_compat import OccGridEstimator
    from morpheus_b200.render import Renderer
    from oracle import render as orr
    from oracle.fields import SceneOracle, init_reference_like_state
    N, S = 256, 64
    sd = init_reference_like_state(200, seed=7, randomize=True, emb_scale=0.02, sphere=True)
    o, d, g = _rays(N, 3)
    aabb = torch.tensor([-1.01, -1.01, -1.01, 1.01, 1.01, 1.01])
    jitter = torch.rand(N, generator=g)
    samples = orr.sample_uniform(o, d, aabb, S, jitter)
    t = torch.full((1, N, 1), 41.0 / 200)
    ids = torch.full((1, N, 1), 41, dtype=torch.long)
    bg = torch.rand(N, 3, generator=g)
    sc = SceneOracle(sd, 1.01, 200, 0.8)
    with torch.no_grad():
        ref = orr.render_rays(sc, o[None], d[None], t, ids, samples, bg_color=bg, shading='albedo')
    m = make_model(sd, 0.8, dev).eval()
    cfg = {'render': {'step_size': 0.01}, 'model': CONFIG['model'], 'train': {}}
    R = Renderer(m, OccGridEstimator(aabb, 128).to(dev), cfg, 200)
    with torch.no_grad():
        out = R.render_rays(o[None].to(dev), d[None].to(dev), t.to(dev), ids.to(dev), bg_color=bg.to(dev), shading='albedo',
                            samples=tuple(s.to(dev) for s in samples))
    e_img = rel_l2(cpu(out['image']).reshape(-1, 3), cpu(ref['image']))
    e_dep = rel_l2(cpu(out['depth']).reshape(-1), cpu(ref['depth']))
    assert e_img < 1e-4 and e_dep < 1e-4, (e_img, e_dep)


def test_uniform_sampler_matches_oracle(dev):
    import ctypes as C
    from morpheus_b200 import _lib
    from oracle import render as orr
    N, S = 100, 16
    o, d, g = _rays(N, 4)
    aabb = torch.tensor([-1.01, -1.01, -1.01, 1.01, 1.01, 1.01])
    jitter = torch.rand(N, generator=g)
    ri_o, t0_o, t1_o = orr.sample_uniform(o, d, aabb, S, jitter)
    ri = torch.empty(N * S, dtype=torch.int64, device=dev)
    t0 = torch.empty(N * S, device=dev)
    t1 = torch.empty(N * S, device=dev)
    od, dd, jd = o.to(dev), d.to(dev), jitter.to(dev)   # keep the device tensors alive across the raw-pointer call
    _lib.check(_lib.lib().mb_sample_rays_uniform(_lib.ptr(od), _lib.ptr(dd), N, S, (C.c_float * 6)(*aabb.tolist()),
                                                 _lib.ptr(jd), _lib.ptr(ri), _lib.ptr(t0), _lib.ptr(t1), _lib.stream()))
    torch.cuda.synchronize()
    assert torch.equal(ri.cpu(), ri_o)
    np.testing.assert_allclose(cpu(t0), t0_o.numpy(), rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(cpu(t1), t1_o.numpy(), rtol=2e-6, atol=1e-6)


def test_adam_vs_torch(dev):
    import ctypes as C
    from morpheus_b200 import _lib
    n = 10000
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(n, generator=g)
    grads = [torch.randn(n, generator=g) * (10 ** float(torch.randn(1, generator=g))) for _ in range(5)]
    pt = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pt], lr=5e-4, betas=(0.9, 0.99), eps=1e-15)
    p = p0.clone().to(dev)
    m = torch.zeros(n, device=dev)
    v = torch.zeros(n, device=dev)
    gid = torch.zeros(n, dtype=torch.uint8, device=dev)
    lr = torch.tensor([5e-4], device=dev)
    step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    for step, gr in enumerate(grads, 1):
        pt.grad = gr.clone()
        opt.step()
        grd = gr.to(dev)
        if step % 2:
            _lib.check(_lib.lib().mb_adam_step(_lib.ptr(p), _lib.ptr(grd), _lib.ptr(m), _lib.ptr(v), _lib.ptr(gid), _lib.ptr(lr),
                                               C.c_uint64(n), C.c_float(0.9), C.c_float(0.99), C.c_float(1e-15), step, _lib.stream()))
        else:   # device-resident step count variant (CUDA-graph replayable)
            step_dev.fill_(step)
            _lib.check(_lib.lib().mb_adam_step_dev(_lib.ptr(p), _lib.ptr(grd), _lib.ptr(m), _lib.ptr(v), _lib.ptr(gid), _lib.ptr(lr),
                                                   C.c_uint64(n), C.c_float(0.9), C.c_float(0.99), C.c_float(1e-15), _lib.ptr(step_dev), _lib.stream()))
    np.testing.assert_allclose(cpu(p), pt.detach().numpy(), rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------------------------------------
# tensor-core (tcgen05) forward engine vs the fp32 SIMT engine, same inputs, both on the GPU
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('M', [128, 1000 + 37])
def test_tc_forward_matches_simt(dev, M):
    from morpheus_b200 import _lib
    from oracle.fields import init_reference_like_state
    sd = init_reference_like_state(200, seed=9, randomize=True, emb_scale=0.3)
    m = make_model(sd, 0.9, dev).eval()
    g = torch.Generator().manual_seed(5)
    x = ((torch.rand(M, 3, generator=g) * 2 - 1) * 0.95).to(dev)
    t = torch.full((M, 1), 77.0 / 200, device=dev)
    light = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1).to(dev)
    outs = {}
    old = _lib.USE_TC
    try:
        for tc in (False, True):
            _lib.USE_TC = tc
            with torch.no_grad():
                a = m(x, t, light, ratio=0.4, shading='lambertian')
                b = m.warp(x, t)
                c = m.density(x, None)
            torch.cuda.synchronize()
            outs[tc] = [a[0], a[1], a[2], a[3], a[4], a[5], b[0], b[1], c['sdf'], c['albedo']]
    finally:
        _lib.USE_TC = old
    names = ['sdf', 'sigma', 'color', 'normal', 'deform', 'normal_raw', 'warp.deform', 'warp.topo', 'cano.sdf', 'cano.albedo']
    errs = {n: rel_l2(cpu(b), cpu(a)) for n, a, b in zip(names, outs[False], outs[True])}
    tol = {'normal': 2e-3, 'normal_raw': 2e-3, 'color': 2e-3, 'sigma': 1e-4}
    bad = {n: e for n, e in errs.items() if e > tol.get(n, 2e-5)}
    assert not bad, f'{bad}\nall: {errs}'


def test_tc_backward_matches_simt(dev):
    """tensor-core backward of the deform/topology nets (stash + wgrad/dgrad on tcgen05) vs the fp32 SIMT backward"""
    from morpheus_b200 import _lib
    from oracle.fields import init_reference_like_state
    sd = init_reference_like_state(200, seed=13, randomize=True, emb_scale=0.3)
    M = 700
    g = torch.Generator().manual_seed(8)
    x = ((torch.rand(M, 3, generator=g) * 2 - 1) * 0.95).to(dev)
    t = torch.full((M, 1), 113.0 / 200, device=dev)
    light = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1).to(dev)
    w = [torch.randn(M, generator=g).to(dev), torch.randn(M, 3, generator=g).to(dev), torch.randn(M, 3, generator=g).to(dev),
         torch.randn(M, 3, generator=g).to(dev)]
    grads = {}
    old = (_lib.USE_TC, _lib.USE_TC_BWD)
    try:
        for tcb in (False, True):
            _lib.USE_TC, _lib.USE_TC_BWD = True, tcb
            _lib.PROFILE.reset()
            _lib.PROFILE.enabled = True
            m = make_model(sd, 0.9, dev).train()
            xg = x.clone().requires_grad_(True)
            sdf, sigma, color, normal, deform, raw = m(xg, t, light, ratio=1.0, shading='albedo_normal')
            loss = (sdf * w[0]).sum() + (color * w[1]).sum() + (normal * w[2]).sum() + (deform * w[3]).sum() * 1e-3
            loss.backward()
            torch.cuda.synchronize()
            launched = set(_lib.PROFILE.summary())
            assert ('field_bwd_warp_tc' in launched) == tcb, launched      # the tensor-core kernel really ran (or not)
            grads[tcb] = {'x': xg.grad.clone(), **{n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}}
    finally:
        _lib.USE_TC, _lib.USE_TC_BWD = old
        _lib.PROFILE.enabled = False
    errs = {n: rel_l2(cpu(grads[True][n]), cpu(grads[False][n])) for n in grads[False] if float(grads[False][n].abs().max()) > 0}
    bad = {n: e for n, e in errs.items() if e > 2e-4}
    assert not bad, f'{bad}\nall: {errs}'


@pytest.mark.parametrize('case', ['albedo_normal', 'lambertian', 'albedo', 'normal_aux', 'normal_warped', 'density'])
def test_tc_backward_sdf_matches_simt(dev, case):
    """tensor-core backward of the SDF / colour nets, shading and FD queries (field_bwd_sdf_tc) vs the fp32 SIMT backward,
    for every query type the training step issues"""
    from morpheus_b200 import _lib
    from oracle.fields import init_reference_like_state
    sd = init_reference_like_state(200, seed=21, randomize=True, emb_scale=0.3)
    sd['sdf2density.beta'] = torch.tensor(0.5)
    M = 500
    g = torch.Generator().manual_seed(4)
    x = ((torch.rand(M, 3, generator=g) * 2 - 1) * 0.95).to(dev)
    t = torch.full((M, 1), 57.0 / 200, device=dev)
    light = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1).to(dev)
    w = [torch.randn(M, generator=g).to(dev), torch.randn(M, 3, generator=g).to(dev), torch.randn(M, 3, generator=g).to(dev),
         torch.randn(M, generator=g).to(dev)]

    def run(m, xg):
        if case in ('albedo_normal', 'lambertian'):
            sdf, sigma, color, normal, deform, raw = m(xg, t, light, ratio=1.0 if case == 'albedo_normal' else 0.3,
                                                       shading='albedo_normal' if case == 'albedo_normal' else 'lambertian')
            return (sdf * w[0]).sum() + (color * w[1]).sum() + (normal * w[2]).sum() + (sigma * w[3]).sum() * 0.05
        if case == 'albedo':
            sdf, sigma, color, _, deform, _ = m(xg, t, None, shading='albedo')
            return (sdf * w[0]).sum() + (color * w[1]).sum() + (sigma * w[3]).sum() * 0.05
        if case == 'normal_aux':
            n, raw = m.normal(xg, topo=None)
            return (n * w[2]).sum()
        if case == 'normal_warped':
            n, raw = m.normal(xg, t=t)
            return (n * w[2]).sum()
        d = m.density(xg, t)
        return (d['sdf'] * w[0]).sum() + (d['albedo'] * w[1]).sum()

    grads = {}
    old = (_lib.USE_TC, _lib.USE_TC_BWD, _lib.USE_TC_BWD_SDF)
    try:
        for tc in (False, True):
            _lib.USE_TC, _lib.USE_TC_BWD, _lib.USE_TC_BWD_SDF = True, tc, tc
            _lib.PROFILE.reset()
            _lib.PROFILE.enabled = True
            m = make_model(sd, 0.9, dev).train()
            xg = x.clone().requires_grad_(True)
            run(m, xg).backward()
            torch.cuda.synchronize()
            launched = set(_lib.PROFILE.summary())
            assert any(k.startswith(('field_bwd_sdf_tc', 'field_bwd_fd_tc')) for k in launched) == tc, launched
            grads[tc] = {'x': xg.grad.clone(), **{n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}}
    finally:
        _lib.USE_TC, _lib.USE_TC_BWD, _lib.USE_TC_BWD_SDF = old
        _lib.PROFILE.enabled = False
    errs = {n: rel_l2(cpu(grads[True][n]), cpu(grads[False][n])) for n in grads[False] if float(grads[False][n].abs().max()) > 0}
    if case in ('normal_aux', 'normal_warped'):
        errs.pop('sdf_net.net.2.bias', None)   # the last-layer bias cancels exactly in the +-eps differences: its gradient is pure round-off
    bad = {n: e for n, e in errs.items() if e > 3e-4}
    assert not bad, f'{bad}\nall: {errs}'


# ------------------------------------------------------------------------------------------------
# host-glue kernels (csrc/glue.cu)
# ------------------------------------------------------------------------------------------------
def test_pack_arena_matches_torch_weight_norm(dev):
    """mb_pack_arena_* == the differentiable torch construction (packing.pack over MLP.effective(): v * g / ||v||,
    models/decoders.py:51-52), forward bit-close and parameter gradients to 1e-6."""
    from morpheus_b200 import packing
    from oracle.fields import init_reference_like_state
    sd = init_reference_like_state(200, seed=5, randomize=True, emb_scale=0.05)
    m = make_model(sd, 1.0, dev).train()
    ref = packing.pack({'deform': m.deform_net.effective(), 'topo': m.topo_net.effective(), 'sdf': m.sdf_net.effective(),
                        'color': m.color_net.effective()})
    arena, tcw = m.packed_arena()
    assert arena.shape == ref.shape
    assert rel_l2(cpu(arena), cpu(ref)) < 1e-7
    assert float((arena - ref).abs().max()) < 1e-6
    g = torch.Generator().manual_seed(1)
    w = torch.randn(ref.shape, generator=g).to(dev)
    # only the Wt and b slots of the gradient arena are populated by the kernels: mask the W (n-major) slots
    mask = torch.zeros_like(w)
    for net in packing.NET_ORDER:
        for (wt_off, w_off, b_off, K, N, Kp, Np) in packing.LAYOUT[net]:
            mask[wt_off:wt_off + Kp * Np] = 1
            mask[b_off:b_off + Np] = 1
    w = w * mask
    params = [p for n_, p in m.named_parameters() if any(k in n_ for k in ('deform_net', 'topo_net', 'sdf_net', 'color_net'))]
    g_ref = torch.autograd.grad((ref * w).sum(), params)
    g_new = torch.autograd.grad((arena * w).sum(), params)
    for a, b, (n_, _) in zip(g_new, g_ref, [(n_, p) for n_, p in m.named_parameters() if any(k in n_ for k in ('deform_net', 'topo_net', 'sdf_net', 'color_net'))]):
        assert rel_l2(cpu(a), cpu(b)) < 2e-6, n_
    # the cache is dropped once the arena's backward has run, and after invalidate()
    arena2, _ = m.packed_arena()
    assert arena2 is not arena


def test_ray_points_and_sdf_loss_match_torch(dev):
    from morpheus_b200 import render as mr
    from morpheus_b200.nerfacc_compat import ray_segments
    g = torch.Generator().manual_seed(3)
    N = 257
    counts = torch.randint(0, 40, (N,), generator=g)
    counts[5] = 0
    ri = torch.repeat_interleave(torch.arange(N), counts).to(dev)
    M = ri.shape[0]
    o = torch.randn(N, 3, generator=g).to(dev).requires_grad_(True)
    d = torch.randn(N, 3, generator=g).to(dev).requires_grad_(True)
    t0 = (torch.rand(M, generator=g) * 3).to(dev)
    t1 = t0 + 0.01
    seg = ray_segments(ri, N)
    xyz = mr._RayPoints.apply(o, d, ri, t0, t1, seg)
    ref = o[ri] + d[ri] * ((t0 + t1) / 2.0)[:, None]
    assert torch.equal(xyz, ref)          # bit-exact: same two roundings (mul, add)
    w = torch.randn(M, 3, generator=g).to(dev)
    ga = torch.autograd.grad((xyz * w).sum(), [o, d])
    gb = torch.autograd.grad((ref * w).sum(), [o, d])
    for a, b in zip(ga, gb):
        assert float((a - b).abs().max()) < 1e-4 * float(b.abs().max())
    # get_sdf_loss (utils.py:91-113): depth with holes (0), invalid (-1), masks, samples in front / inside / behind the band
    depth = (torch.rand(N, 1, generator=g) * 3 + 0.2)
    depth[::7] = 0.0
    depth[3::11] = -1.0
    mask = (torch.rand(N, 1, generator=g) > 0.3).float()
    depth, mask = depth.to(dev), mask.to(dev)
    sdf = (torch.randn(M, generator=g) * 0.2).to(dev).requires_grad_(True)
    z = ((t0 + t1) / 2.0)[:, None]
    fs_ref, sl_ref = mr.get_sdf_loss(z, depth[ri], sdf, 0.1, mask=mask[ri])
    fs, sl = mr.packed_sdf_loss(sdf, t0, t1, ri, depth, mask, 0.1)
    assert abs(float(fs) - float(fs_ref)) < 1e-5 * max(1.0, abs(float(fs_ref)))
    assert abs(float(sl) - float(sl_ref)) < 1e-5 * max(1.0, abs(float(sl_ref)))
    ga = torch.autograd.grad(2.0 * fs + 3.0 * sl, sdf)[0]
    gb = torch.autograd.grad(2.0 * fs_ref + 3.0 * sl_ref, sdf)[0]
    assert rel_l2(cpu(ga), cpu(gb)) < 1e-5


def test_pose_rays_and_ray_loss_match_torch(dev):
    """fused pose correction (models/model.py:335-346) and loss heads (morpheus.py:946-983) vs their eager torch forms"""
    from morpheus_b200 import train as mtrain
    from oracle.fields import init_reference_like_state
    sd = init_reference_like_state(200, seed=2, randomize=True, emb_scale=0.02)
    m = make_model(sd, 1.0, dev).train()
    g = torch.Generator().manual_seed(9)
    with torch.no_grad():
        m.pose_array.data.copy_((torch.randn(200, 6, generator=g) * 0.05).to(dev))
    N = 777
    o = torch.randn(N, 3, generator=g).to(dev).requires_grad_(True)
    d = torch.randn(N, 3, generator=g).to(dev).requires_grad_(True)
    ids = torch.randint(0, 200, (N, 1), generator=g).to(dev)
    ids[:400] = 17                                  # warp-uniform and mixed warps
    o2, d2 = m.pose_optimisation(o, d, ids)
    R = m.pose_array.get_rotation_matrices(ids.squeeze())
    o2r, d2r = o + m.pose_array.get_translations(ids.squeeze()), torch.sum(d[..., None, :] * R, -1)
    assert torch.equal(o2, o2r)
    assert float((d2 - d2r).abs().max()) < 1e-6
    w1, w2 = torch.randn(N, 3, generator=g).to(dev), torch.randn(N, 3, generator=g).to(dev)
    ga = torch.autograd.grad((o2 * w1).sum() + (d2 * w2).sum(), [m.pose_array.data, o, d])
    gb = torch.autograd.grad((o2r * w1).sum() + (d2r * w2).sum(), [m.pose_array.data, o, d])
    for a, b in zip(ga, gb):
        assert rel_l2(cpu(a), cpu(b)) < 1e-5
    # loss heads
    tr = dict(mtrain.DEFAULT_TRAIN_CFG)
    img = torch.rand(N, 3, generator=g).to(dev).requires_grad_(True)
    dep = (torch.rand(N, generator=g) * 3).to(dev).requires_grad_(True)
    opa = torch.rand(N, generator=g)
    opa[:5] = 0.0
    opa[5:9] = 1.0
    opa = opa.to(dev).requires_grad_(True)
    gt_depth = torch.rand(N, generator=g) * 2
    gt_depth[::5] = 0.0
    batch = {'rgb': torch.rand(N, 3, generator=g).to(dev), 'depth': gt_depth.to(dev), 'mask': (torch.rand(N, generator=g) > 0.4).float().to(dev),
             'rays_o': (torch.randn(N, 3, generator=g) * 0.3).to(dev), 'rays_d': (torch.randn(N, 3, generator=g) * 0.3).to(dev)}
    out = {'image': img, 'depth': dep, 'weights_sum': opa}
    la = mtrain.real_view_loss(out, batch, m, tr)
    lb = mtrain.real_view_loss_torch(out, batch, m, tr)
    assert abs(float(la) - float(lb)) < 1e-5 * max(1.0, abs(float(lb)))
    ga = torch.autograd.grad(la, [img, dep, opa])
    gb = torch.autograd.grad(lb, [img, dep, opa])
    for a, b in zip(ga, gb):
        assert rel_l2(cpu(a), cpu(b)) < 1e-5


def test_normal_smoothness_and_surface_point_losses_vs_oracle(dev):
    """get_normal_smoothness_loss (morpheus.py:530-556) and the surface-point terms of
    get_real_view_point_loss (:1005-1029) on the fused normal(x, t) / density(x, t) queries vs the CPU oracle scene."""
    import math
    from morpheus_b200 import train as mtrain
    from morpheus_b200.nerfacc_compat import OccGridEstimator
    from morpheus_b200.render import Renderer
    from oracle.fields import SceneOracle, init_reference_like_state
    sd = init_reference_like_state(200, seed=21, randomize=True, emb_scale=0.3, sphere=True)
    m = make_model(sd, 1.0, dev).train()
    tr = dict(mtrain.FULL_TRAIN_CFG)
    cfg = {'model': {'bg_radius': 1.4, 'activation': 'exp'}, 'render': {'step_size': 0.01}, 'train': tr}
    R = Renderer(m, OccGridEstimator(torch.tensor([-1.01] * 3 + [1.01] * 3), 128).to(dev), cfg, 200)
    g = torch.Generator().manual_seed(4)
    N = 40
    o = (torch.randn(N, 3, generator=g) * 0.2)
    d = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=-1) * 0.5
    depth = torch.rand(N, generator=g) * 1.5 + 0.2
    depth[:4] = 5.0                                        # points outside the 1.1 sphere are dropped
    t = torch.full((N, 1), 31.0 / 200)
    noise = torch.rand(11, generator=g)
    phi = torch.rand(11 * N, 1, generator=g) * 2 * math.pi
    # oracle (CPU, eager torch, boolean indexing as the reference)
    sc = SceneOracle({k: v.clone() for k, v in sd.items()}, 1.01, 200, 1.0)
    oc, dc, depc = o.clone().requires_grad_(True), d.clone(), depth.clone().requires_grad_(True)
    tn = torch.linspace(-0.05, 0.05, 11) + 0.01 * noise
    pts = ((depc[None, :] + tn[:, None])[..., None] * dc[None] + oc[None]).view(-1, 3)
    ts = t[None].repeat(11, 1, 1).view(-1, 1)
    keep = torch.linalg.norm(pts, dim=-1) < 1.1
    n1, _ = sc.normal(pts[keep], t=ts[keep])
    w = Renderer.get_ortho_normal_dir(n1, phi[keep])
    n2, _ = sc.normal(pts[keep] + w * tr['smoothness_std'], t=ts[keep])
    ref = torch.mean(torch.square(n1 - n2))
    g_ref = torch.autograd.grad(ref, [oc, depc])
    og, dg = o.to(dev).requires_grad_(True), depth.to(dev).requires_grad_(True)
    out = R.get_normal_smoothness_loss(og, d.to(dev), t.to(dev), dg, trunc_noise=noise.to(dev), phi=phi.to(dev))
    g_out = torch.autograd.grad(out, [og, dg])
    assert abs(float(out) - float(ref)) < 2e-3 * abs(float(ref)) + 1e-7, (float(out), float(ref))
    for a, b in zip(g_out, g_ref):
        assert rel_l2(cpu(a), b.numpy()) < 2e-2            # second differences of FD normals (eps 2e-3): ill-conditioned in fp32
    # surface-point terms
    batch = {'rays_o': o.to(dev), 'rays_d': d.to(dev), 'rays_t': t.to(dev), 'depth': depth.to(dev), 'mask': (torch.rand(N, generator=g) > 0.3).float().to(dev),
             'rgb': torch.rand(N, 3, generator=g).to(dev)}
    loss = mtrain.surface_point_loss(m, batch, tr)
    xyz = o + depth[:, None] * d
    dm = ((depth > 0) & (xyz.norm(dim=-1) <= 1.1) & (batch['mask'].cpu() > 0.5))
    res = sc.density(xyz, t=t)
    ref = 10.0 * torch.mean(torch.square(res['sdf'][dm])) + 5.0 * torch.nn.functional.mse_loss(res['albedo'] * dm[:, None].float(), batch['rgb'].cpu() * dm[:, None].float())
    assert abs(float(loss) - float(ref)) < 1e-4 * abs(float(ref)) + 1e-7, (float(loss), float(ref))


def test_forward_only_consumers(dev):
    """dense SDF volume for mesh export (morpheus.py:384-396) and chunked eval render vs the CPU oracle."""
    from morpheus_b200.nerfacc_compat import OccGridEstimator
    from morpheus_b200.render import Renderer
    from oracle import render as orr
    from oracle.fields import SceneOracle, init_reference_like_state
    sd = init_reference_like_state(200, seed=3, randomize=True, emb_scale=0.05, sphere=True)
    m = make_model(sd, 1.0, dev).eval()
    cfg = {'model': {'bg_radius': 1.4, 'activation': 'exp'}, 'render': {'step_size': 0.01}, 'train': {}}
    R = Renderer(m, OccGridEstimator(torch.tensor([-1.01] * 3 + [1.01] * 3), 128).to(dev), cfg, 200, uniform_samples=32)
    vol = R.sdf_volume(resolution=12, S=5, t=0.25)
    sc = SceneOracle(sd, 1.01, 200, 1.0)
    ax = torch.linspace(-1, 1, 12)
    xx, yy, zz = torch.meshgrid(ax, ax, ax, indexing='ij')
    pts = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], -1)
    with torch.no_grad():
        ref = sc.density(pts, t=torch.full((pts.shape[0], 1), 0.25), return_color=False)['sdf'].reshape(12, 12, 12)
    assert rel_l2(cpu(vol), ref.numpy()) < 1e-4
    # chunked eval render == one-shot render
    g = torch.Generator().manual_seed(1)
    c2w = orr.look_at_pose(70.0, 40.0, 2.5)
    dirs = orr.camera_dirs(360, 360, 517.0, 517.0, 180.0, 180.0).reshape(-1, 3)
    idx = torch.randint(0, dirs.shape[0], (300,), generator=g)
    o, d = orr.rays_from_pose(dirs[idx], c2w)
    t = torch.full((300, 1), 0.25)
    ids = torch.full((300, 1), 50, dtype=torch.long)
    jit = torch.rand(300, generator=g).to(dev)
    a = R.render_image(o.to(dev), d.to(dev), t.to(dev), ids.to(dev), chunk=128, jitter=None, shading='albedo')
    assert a['image'].shape == (300, 3) and a['depth'].shape == (300,) and torch.isfinite(a['image']).all()
    assert m.training is False


def test_gradient_sinks_and_code_regulariser(dev):
    """train.FlatAdam switches the model's gradient sink on: the field kernels, mb_pack_arena_backward (direct mode) and mb_code_reg
    accumulate straight into the flat gradient buffer.  Same step with the sink off (plain autograd accumulation) -> same gradients;
    fused code regulariser == three MultiCode.sample calls (morpheus.py:766-771)."""
    from morpheus_b200 import train as mtrain
    from morpheus_b200.nerfacc_compat import OccGridEstimator
    from morpheus_b200.rays import synthetic_real_view_batch
    from morpheus_b200.render import Renderer
    from oracle.fields import init_reference_like_state
    sd = init_reference_like_state(200, seed=17, randomize=True, emb_scale=0.05, sphere=True)
    tr = dict(mtrain.DEFAULT_TRAIN_CFG)
    cfg = {'model': {'bg_radius': 1.4, 'activation': 'exp'}, 'render': {'step_size': 0.01}, 'train': tr}
    batch = {k: v.to(dev) for k, v in synthetic_real_view_batch(96, seed=5, frame=40).items()}
    grads, losses = [], []
    for sink in (False, True):
        m = make_model(sd, 1.0, dev).train()
        R = Renderer(m, OccGridEstimator(torch.tensor([-1.01] * 3 + [1.01] * 3), 128).to(dev), cfg, 200, uniform_samples=24)
        opt = mtrain.FlatAdam(m, tr['lr'])
        m.grad_sink = sink
        torch.manual_seed(3)
        losses.append(float(mtrain.train_step_compute(R, opt, batch, tr)))
        torch.cuda.synchronize()
        grads.append(opt.grad.clone())
        # fused code regulariser vs the eager formulation
        t = torch.full((1, 1), 40 / 200, device=dev)
        codes = m.get_deform_code(torch.cat([t, t - 1 / 200, t + 1 / 200], dim=0))
        ref = torch.square(2 * codes[0:1] - codes[1:2] - codes[2:3]).mean()
        assert abs(float(m.code_regulariser(t, 200)) - float(ref)) < 1e-6 * max(1.0, abs(float(ref)))
    assert abs(losses[0] - losses[1]) < 1e-6 * max(1.0, abs(losses[0]))
    assert float(grads[0].abs().max()) > 0
    assert rel_l2(cpu(grads[1]), cpu(grads[0])) < 1e-5


def test_edge_cases_empty_tiny_and_errors(dev):
    """Edge cases the reference handles (or trips over): empty sample set -> white image / zero depth (morpheus.py:663-670);
    M = 0 launches are no-ops; null pointers surface as RuntimeError (the reference raises from TORCH_CHECK); tiny and
    non-multiple-of-16 query counts through the FD backward kernel (partial sub-tiles) match the fp32 SIMT engine."""
    import ctypes as C
    from morpheus_b200 import _lib
    from morpheus_b200.nerfacc_compat import OccGridEstimator
    from morpheus_b200.render import Renderer
    from oracle.fields import init_reference_like_state
    sd = init_reference_like_state(200, seed=8, randomize=True, emb_scale=0.2)
    m = make_model(sd, 1.0, dev).train()
    cfg = {'model': {'bg_radius': 1.4, 'activation': 'exp'}, 'render': {'step_size': 0.01},
           'train': {'ori_weight': 0.0, 'normal_smooth_3d': 0.0, 'code_reg': 0.0, 'trunc': 0.1}}
    est = OccGridEstimator(torch.tensor([-1.01] * 3 + [1.01] * 3), 128).to(dev)      # binaries all False: every ray is empty
    R = Renderer(m, est, cfg, 200)
    N = 7
    o = torch.tensor([[0.0, 0.0, 2.5]]).repeat(N, 1).to(dev)
    d = torch.tensor([[0.0, 0.0, -1.0]]).repeat(N, 1).to(dev)
    out = R.render_rays(o[None], d[None], torch.full((1, N, 1), 0.1, device=dev), torch.zeros(1, N, 1, dtype=torch.long, device=dev))
    assert out['image'].shape == (1, N, 3) and float((out['image'] - 1).abs().max()) == 0.0
    assert out['depth'].shape == (1, N) and float(out['depth'].abs().max()) == 0.0 and out['sdf'] is None
    # M = 0 / N = 0 launches
    L = _lib.lib()
    assert L.mb_ray_points_forward(None, None, None, None, None, 0, None, _lib.stream()) == 0
    assert L.mb_sdf_loss_forward(_lib.ptr(o), _lib.ptr(o), _lib.ptr(o), _lib.ptr(o), None, _lib.ptr(o), 0, C.c_float(0.1), _lib.ptr(o), _lib.stream()) == 0
    assert L.mb_pose_rays_forward(None, None, None, None, 0, None, None, _lib.stream()) == 0
    # null pointers -> error code + message -> RuntimeError in the Python shim
    rc = L.mb_ray_points_forward(None, None, None, None, None, 5, None, _lib.stream())
    assert rc != 0 and b'null' in L.mb_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(L.mb_pack_arena_forward(None, 18, None, _lib.stream()), 'pack_arena_forward')
    with pytest.raises(RuntimeError):
        _lib.check(L.mb_field_backward_fd_tc(None, None, None, None, None, None, None, 0, _lib.stream()), 'field_backward_fd_tc')
    with pytest.raises(RuntimeError):
        m.normal(torch.zeros(0, 3, device=dev), topo=None)          # empty query: explicit error instead of the reference's NameError
    # tiny / ragged FD queries: tensor-core FD kernel vs SIMT engine
    g = torch.Generator().manual_seed(2)
    old = (_lib.USE_TC, _lib.USE_TC_BWD, _lib.USE_TC_BWD_SDF)
    try:
        for M in (1, 17, 130):
            x = ((torch.rand(M, 3, generator=g) * 2 - 1) * 0.9).to(dev)
            w = torch.randn(M, 3, generator=g).to(dev)
            res = {}
            for tc in (False, True):
                _lib.USE_TC, _lib.USE_TC_BWD, _lib.USE_TC_BWD_SDF = True, tc, tc
                mm = make_model(sd, 1.0, dev).train()
                xg = x.clone().requires_grad_(True)
                n, raw = mm.normal(xg, topo=None)
                (n * w).sum().backward()
                res[tc] = (xg.grad.clone(), mm.encoder.embeddings.grad.clone(), mm.sdf_net.net[0].weight.grad.clone())
            for a, b in zip(res[True], res[False]):
                assert rel_l2(cpu(a), cpu(b)) < 3e-4, M
    finally:
        _lib.USE_TC, _lib.USE_TC_BWD, _lib.USE_TC_BWD_SDF = old


# ------------------------------------------------------------------------------------------------
# fused FD-normal regulariser (csrc/field_fd_reg_tc.cu) -- morpheus.py:714-741 on a real view
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('M,max_level', [(1, 1.0), (17, 0.77), (130, 1.0), (1000 + 37, 0.77), (4096 + 5, 1.0)])
def test_fd_regulariser_vs_fp64_oracle_and_general_path(dev, M, max_level):
    """loss_normal_perturb = mean |normal(x, topo) - normal(x + noise * std, topo=None)|: the one-launch forward+backward kernel against
    (a) the CPU oracle in float64 with autograd and (b) the library's general path (two `normal` queries + torch L1).  Points include the
    AABB faces (clamped +-eps rows), exact cell centres / faces of the finest level and out-of-range samples; ragged M exercises partial
    tiles, partial sub-tiles and (M = 4101) several tiles per CTA."""
    from oracle.fields import SceneOracle, init_reference_like_state
    sd = init_reference_like_state(200, seed=5, randomize=True, emb_scale=0.3)
    g = torch.Generator().manual_seed(100 + M)
    x = (torch.rand(M, 3, generator=g) * 2 - 1) * 1.0
    if M >= 17:
        x[0] = torch.tensor([1.01, -1.01, 0.3]); x[1] = torch.tensor([1.0095, 0.0, -1.0095])      # on / next to the AABB faces
        x[2] = (torch.tensor([64.0, 17.0, 100.0]) / 128) * 2.02 - 1.01                           # exactly on cell faces of the finest level
        x[3] = (torch.tensor([64.5, 17.5, 100.5]) / 128) * 2.02 - 1.01                           # exactly on cell centres
        x[4] = torch.tensor([1.2, 0.1, 0.1])                                                      # outside the AABB (clamped rows)
    topo = torch.randn(M, 2, generator=g) * 0.3
    noise = torch.randn(M, 3, generator=g)
    std = 0.005
    # (a) fp64 oracle
    sdo = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        sc = SceneOracle(sdo, 1.01, 200, max_level)
        xo, to = x.double().requires_grad_(True), topo.double().requires_grad_(True)
        n1, _ = sc.normal(xo, topo=to)
        n2, _ = sc.normal(xo + noise.double() * std, topo=None)
        lo = (n1 - n2).abs().mean()
        lo.backward()
    finally:
        torch.set_default_dtype(prev)
    # (b) general path
    m = make_model(sd, max_level, dev).train()
    xg, tg = x.to(dev).requires_grad_(True), topo.to(dev).requires_grad_(True)
    a1, _ = m.normal(xg, topo=tg)
    a2, _ = m.normal(xg + noise.to(dev) * std, topo=None)
    lg = (a1 - a2).abs().mean()
    lg.backward()
    gen = {'x': xg.grad.clone(), 'topo': tg.grad.clone(), 'emb': m.encoder.embeddings.grad.clone(),
           **{f'sdf{l}.{k}': getattr(m.sdf_net.net[l], k).grad.clone() for l in range(3) for k in ('weight', 'bias') if getattr(m.sdf_net.net[l], k).grad is not None}}
    m.zero_grad()
    # fused
    xf, tf = x.to(dev).requires_grad_(True), topo.to(dev).requires_grad_(True)
    lf, nf, rf = m.fd_regulariser(xf, tf, noise.to(dev), std, 1.0 / (3 * M))
    (lf * 1.7).backward()                  # a non-unit upstream gradient: backward() only scales the stored gradients
    fus = {'x': xf.grad / 1.7, 'topo': tf.grad / 1.7, 'emb': m.encoder.embeddings.grad / 1.7,
           **{f'sdf{l}.{k}': getattr(m.sdf_net.net[l], k).grad / 1.7 for l in range(3) for k in ('weight', 'bias') if getattr(m.sdf_net.net[l], k).grad is not None}}
    # (every FD normal differences fp32 SDF values 4e-3 apart: golden tolerance 2e-3 per normal)
    assert abs(float(lf) - float(lo)) <= 2e-4 * abs(float(lo)) + 1e-7, (float(lf), float(lo))
    assert abs(float(lf) - float(lg)) <= 2e-5 * abs(float(lg)) + 1e-7, (float(lf), float(lg))
    assert rel_l2(cpu(nf), cpu(a1)) < 1e-5 and rel_l2(cpu(nf), cpu(n1)) < 2e-3      # FD normals difference fp32 SDF values 4e-3 apart
    ora = {'x': xo.grad, 'topo': to.grad, 'emb': sdo['encoder.embeddings'].grad,
           **{f'sdf{l}.{k}': sdo[f'sdf_net.net.{l}.{k}'].grad for l in range(3) for k in ('weight', 'bias') if sdo[f'sdf_net.net.{l}.{k}'].grad is not None}}
    errs_o = {k: rel_l2(cpu(fus[k]), cpu(ora[k])) for k in fus if k in ora and float(ora[k].abs().max()) > 0}
    errs_g = {k: rel_l2(cpu(fus[k]), cpu(gen[k])) for k in fus if k in gen and float(gen[k].abs().max()) > 0}
    floor = {k: rel_l2(cpu(gen[k]), cpu(ora[k])) for k in gen if k in ora and float(ora[k].abs().max()) > 0}
    # bar: within 1e-3 of the exact (fp64) gradient, or at least as close to it as the general fp32 path
    bad = {k: e for k, e in errs_o.items() if e > max(1e-3, 1.5 * floor.get(k, 0.0))}
    assert len(errs_o) >= 7 and not bad, f'vs fp64 oracle: {bad}\nfused-vs-oracle {errs_o}\ngeneral-vs-oracle {floor}\nfused-vs-general {errs_g}'


# ------------------------------------------------------------------------------------------------
# occupancy refresh (morpheus.py:905-913 -> nerfacc update_every_n_steps)
# ------------------------------------------------------------------------------------------------
def test_occupancy_refresh_vs_oracle(dev):
    """`Renderer.update_occ_grid` (the fused no-grad density query over the probed cells + mb_occ_update / mb_occ_binarize_dev) against the
    oracle restatement with the cell jitter and the post-warm-up cell subset injected: the warm-up refresh (every cell) and a later
    refresh (uniform + occupied subset, EMA decay of the previous values) on a 32^3 grid, then the shipped 128^3 grid against itself
    (second refresh with the same jitter must reproduce occs / binaries exactly: max(0.95 occs, occ) with occ unchanged)."""
    from morpheus_b200.nerfacc_compat import OccGridEstimator
    from morpheus_b200.render import Renderer
    from oracle import render as orr
    from oracle.fields import SceneOracle, init_reference_like_state
    sd = init_reference_like_state(200, seed=13, randomize=True, emb_scale=0.02, sphere=True)
    sd['sdf2density.beta'] = torch.tensor(0.05)
    aabb = torch.tensor([-1.01, -1.01, -1.01, 1.01, 1.01, 1.01])
    res = 32
    m = make_model(sd, 0.8, dev).train()
    cfg = {'render': {'step_size': 0.01}, 'model': CONFIG['model'], 'train': {}}
    est = OccGridEstimator(aabb, res).to(dev).train()
    R = Renderer(m, est, cfg, 200)
    rays_t = torch.full((7, 1), 41.0 / 200)
    sc = SceneOracle(sd, 1.01, 200, 0.8)

    def occ_fn_oracle(x):
        with torch.no_grad():
            return sc.density(x, rays_t, allow_shape=True, return_color=False)['sigma'] * 0.01

    def occ_fn_ours(x):
        return m.density(x, rays_t.to(dev), allow_shape=True, return_color=False)['sigma'] * 0.01
    g = torch.Generator().manual_seed(5)
    n = res ** 3
    # ---- warm-up refresh: every cell ----
    jit = torch.rand(n, 3, generator=g)
    est._update(0, occ_fn_ours, 1e-2, 0.95, 256, cell_jitter=jit)
    occs_o, bin_o = orr.occ_grid_update(torch.zeros(n), torch.arange(n), jit, aabb, res, occ_fn_oracle)
    assert rel_l2(cpu(est.occs), cpu(occs_o)) < 1e-4
    mism = (est.binaries.flatten().cpu() != bin_o)
    near = (occs_o - torch.clamp(occs_o[occs_o >= 0].mean(), max=1e-2)).abs() < 1e-4 * occs_o.abs().clamp(min=1e-6)      # cells sitting ON the threshold
    assert int((mism & ~near).sum()) == 0 and 0 < int(bin_o.sum()) < n
    # ---- later refresh: uniform + occupied subset, EMA over the previous state ----
    uni = torch.randint(n, (n // 4,), generator=g)
    occ_idx = torch.nonzero(bin_o)[:, 0]
    if occ_idx.shape[0] > n // 4:
        occ_idx = occ_idx[torch.randint(occ_idx.shape[0], (n // 4,), generator=g)]
    idx = torch.cat([uni, occ_idx])
    jit2 = torch.rand(idx.shape[0], 3, generator=g)
    # (duplicate cells in idx race in both implementations: keep the first occurrence only)
    first = torch.zeros(n, dtype=torch.bool)
    keep = []
    for k, c in enumerate(idx.tolist()):
        if not first[c]:
            first[c] = True
            keep.append(k)
    idx, jit2 = idx[keep], jit2[keep]
    est._update(400, occ_fn_ours, 1e-2, 0.95, 256, cell_idx=idx, cell_jitter=jit2)
    occs_o2, bin_o2 = orr.occ_grid_update(occs_o, idx, jit2, aabb, res, occ_fn_oracle)
    assert rel_l2(cpu(est.occs), cpu(occs_o2)) < 1e-4
    mism = (est.binaries.flatten().cpu() != bin_o2)
    near = (occs_o2 - torch.clamp(occs_o2[occs_o2 >= 0].mean(), max=1e-2)).abs() < 1e-4 * occs_o2.abs().clamp(min=1e-6)
    assert int((mism & ~near).sum()) == 0
    # ---- the shipped 128^3 grid through the public entry point (2 097 152 points, step 0), idempotence of a repeated refresh ----
    est128 = OccGridEstimator(aabb, 128).to(dev).train()
    R128 = Renderer(m, est128, cfg, 200)
    torch.manual_seed(3)
    R128.update_occ_grid(rays_t.to(dev), step=0)
    occs1, bin1 = est128.occs.clone(), est128.binaries.clone()
    assert 0 < int(bin1.sum()) < 128 ** 3
    torch.manual_seed(3)
    R128.update_occ_grid(rays_t.to(dev), step=16)        # same jitter: occ unchanged -> max(0.95 occs, occ) == occs
    assert torch.equal(est128.occs, occs1) and torch.equal(est128.binaries, bin1)
    R128.update_occ_grid(rays_t.to(dev), step=17)        # not a multiple of 16: no refresh
    assert torch.equal(est128.occs, occs1)


# ------------------------------------------------------------------------------------------------
# ray generation on the device, eval render, virtual-view render with the orientation loss
# ------------------------------------------------------------------------------------------------
def test_ray_generation_on_device_vs_reference_goldens(dev, golden_dir):
    """SURVEY 8a-1 / 8f-3 on the GPU: the device-resident dataset (rays of the selected pixels computed on the fly, no host gather, no
    H2D copy) and the novel-view ray generator against the goldens produced by the unmodified reference dataset methods
    (datasets/dataset.py:336-433, :435-578; tests/golden/make_loss_golden.py)."""
    from morpheus_b200 import rays
    z = np.load(os.path.join(golden_dir, 'real_view_rays.npz'))
    data = rays.RealViewData(z['images'], z['depths'], z['masks'], z['poses'], z['K'], device=dev)
    s = data.sample_real_view_rays(idx=torch.from_numpy(z['idx']), ray_num=13, index=torch.from_numpy(z['index']))
    assert all(s[k].is_cuda for k in ('rays_o', 'rays_d', 'rays_t', 'rays_id', 'image', 'depth', 'mask'))
    for k in ('rays_o', 'rays_d', 'rays_t', 'image', 'depth'):
        assert s[k].shape == z['s_' + k].shape and np.allclose(cpu(s[k]), z['s_' + k], rtol=0, atol=1e-6), k
    for k in ('rays_id', 'mask'):
        assert np.array_equal(cpu(s[k]), z['s_' + k]), k
    f = data.sample_real_view_rays(idx=2)
    for k in ('rays_o', 'rays_d', 'rays_t', 'image', 'depth'):
        assert np.allclose(cpu(f[k]), z['f_' + k], rtol=0, atol=1e-6), k
    v = np.load(os.path.join(golden_dir, 'virtual_views.npz'))
    for i in range(int(v['n'])):
        g = lambda k: v[f'v{i}_{k}']      # noqa: E731
        w = rays.virtual_view_rays(frame=12, num_frames=200, H=360, W=360, focal=517.0, scale=0.2, theta_deg=float(g('polar')[0]) + 90.0,
                                   phi_deg=float(g('azimuth')[0]), device=dev)
        assert w['rays_o'].is_cuda and np.allclose(cpu(w['rays_o']), g('rays_o'), rtol=0, atol=2e-5)
        assert np.allclose(cpu(w['rays_d']), g('rays_d'), rtol=0, atol=2e-5)
        assert np.allclose(cpu(w['rays_t']), g('rays_t')) and np.array_equal(cpu(w['rays_id']), g('rays_id'))


def test_eval_render_image_vs_oracle(dev):
    """`Renderer.render_image` (eval_step, morpheus.py:1238-1269: model.eval(), chunked rays, shading 'albedo') against the
    oracle render on the same samples -- image and depth within the north star's 1e-4."""
    from morpheus_b200.nerfacc_compat import OccGridEstimator
    from morpheus_b200.render import Renderer
    from oracle import render as orr
    from oracle.fields import SceneOracle, init_reference_like_state
    sd = init_reference_like_state(200, seed=3, randomize=True, emb_scale=0.05, sphere=True)
    sd['sdf2density.beta'] = torch.tensor(0.3)
    N, S = 300, 32
    o, d, g = _rays(N, 11)
    aabb = torch.tensor([-1.01, -1.01, -1.01, 1.01, 1.01, 1.01])
    jitter = torch.rand(N, generator=g)
    t = torch.full((N, 1), 0.25)
    ids = torch.full((N, 1), 50, dtype=torch.long)
    sc = SceneOracle(sd, 1.01, 200, 1.0)
    with torch.no_grad():
        ref = orr.render_rays(sc, o, d, t, ids, orr.sample_uniform(o, d, aabb, S, jitter), bg_color=None, shading='albedo')
    m = make_model(sd, 1.0, dev).train()
    cfg = {'model': {'bg_radius': 1.4, 'activation': 'exp'}, 'render': {'step_size': 0.01}, 'train': {}}
    R = Renderer(m, OccGridEstimator(aabb, 128).to(dev), cfg, 200, uniform_samples=S)
    out = R.render_image(o.to(dev), d.to(dev), t.to(dev), ids.to(dev), chunk=512, jitter=jitter.to(dev), shading='albedo')
    assert m.training is True                      # the previous mode is restored
    assert rel_l2(cpu(out['image']), cpu(ref['image'])) < 1e-4 and rel_l2(cpu(out['depth']), cpu(ref['depth'])) < 1e-4
    assert rel_l2(cpu(out['weights_sum']).reshape(-1), cpu(ref['weights_sum']).reshape(-1)) < 1e-4


def test_virtual_view_render_with_orientation_loss_vs_oracle(dev):
    """MorpheuS.render_rays on a VIRTUAL view (morpheus.py:558-794 with real_view=False): lambertian shading (the colour depends on the FD
    normal and the injected light direction), white background, loss_orient = sum w.detach() max(n . d, 0)^2 (:709-712), loss_normal_perturb --
    values and the gradients of image.sum() + loss_orient + loss_normal_perturb against the oracle's autograd.  (General path: forward FD
    chains + field_bwd_fd_tc; the fused FD regulariser only serves real views.)"""
    from morpheus_b200.nerfacc_compat import OccGridEstimator
    from morpheus_b200.render import Renderer
    from oracle import render as orr
    from oracle.fields import SceneOracle, init_reference_like_state, safe_normalize
    sd = init_reference_like_state(200, seed=9, randomize=True, emb_scale=0.05, sphere=True)
    sd['sdf2density.beta'] = torch.tensor(0.3)
    N, S = 96, 32
    o, d, g = _rays(N, 5)
    aabb = torch.tensor([-1.01, -1.01, -1.01, 1.01, 1.01, 1.01])
    jitter = torch.rand(N, generator=g)
    samples = orr.sample_uniform(o, d, aabb, S, jitter)
    t = torch.full((N, 1), 0.4)
    ids = torch.full((N, 1), 80, dtype=torch.long)
    light = safe_normalize(o + torch.randn(3, generator=g))
    noise = torch.randn(N * S, 3, generator=g)
    sdo = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    sc = SceneOracle(sdo, 1.01, 200, 0.9)
    ref = orr.render_rays(sc, o, d, t, ids, samples, bg_color=None, ambient_ratio=0.4, light_d=light, shading='lambertian', perturb_noise=noise,
                          training=True, real_view=False)
    lo = ref['image'].sum() + 0.01 * ref['loss_orient'] + ref['loss_normal_perturb']
    lo.backward()
    m = make_model(sd, 0.9, dev).train()
    tr = {'ori_weight': 0.01, 'normal_smooth_3d': 0.1, 'smoothness_std': 0.005, 'topo_none': True, 'normal_dir': False, 'code_reg': 0.0, 'trunc': 0.1}
    cfg = {'model': {'bg_radius': 1.4, 'activation': 'exp'}, 'render': {'step_size': 0.01}, 'train': tr}
    R = Renderer(m, OccGridEstimator(aabb, 128).to(dev), cfg, 200)
    out = R.render_rays(o.to(dev), d.to(dev), t.to(dev), ids.to(dev), bg_color=None, ambient_ratio=0.4, light_d=light.to(dev), shading='lambertian',
                        real_view=False, samples=tuple(s.to(dev) for s in samples), perturb_noise=noise.to(dev))
    l = out['image'].sum() + 0.01 * out['loss_orient'] + out['loss_normal_perturb']
    l.backward()
    assert rel_l2(cpu(out['image']).reshape(-1, 3), cpu(ref['image'])) < 1e-4
    assert abs(float(out['loss_orient']) - float(ref['loss_orient'])) <= 2e-3 * abs(float(ref['loss_orient'])) + 1e-7
    assert abs(float(out['loss_normal_perturb']) - float(ref['loss_normal_perturb'])) <= 2e-4 * abs(float(ref['loss_normal_perturb']))
    errs = {}
    for name, p in m.named_parameters():
        if name in sdo and sdo[name].grad is not None and p.grad is not None and float(sdo[name].grad.abs().max()) > 0:
            errs[name] = rel_l2(cpu(p.grad), cpu(sdo[name].grad))
    bad = {k: e for k, e in errs.items() if e > 1e-2}
    assert len(errs) >= 20 and not bad, f'{bad}\nall: {errs}'

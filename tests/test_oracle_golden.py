"""CPU: the oracle restatement (oracle/fields.py) against golden vectors produced by the UNMODIFIED
reference Python (tests/golden/make_scene_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle.fields import SceneOracle, init_reference_like_state

CASES = ['init_full', 'rand_c2f', 'rand_full']


def load(golden_dir, tag):
    z = np.load(os.path.join(golden_dir, f'scene_{tag}.npz'))
    ml = float(z['max_level'])
    sd = init_reference_like_state(200, seed=int(z['seed']), randomize=bool(z['randomize']), emb_scale=float(z['emb_scale']))
    return z, sd, (None if ml < 0 else ml)


def close(a, ref, rtol=2e-5, atol=2e-6):
    a = a.detach().numpy() if torch.is_tensor(a) else a
    np.testing.assert_allclose(a, ref, rtol=rtol, atol=atol)


@pytest.mark.parametrize('tag', CASES)
def test_forward_modes(golden_dir, tag):
    z, sd, ml = load(golden_dir, tag)
    sc = SceneOracle(sd, 1.01, 200, ml)
    x, t, light = (torch.from_numpy(z[k]) for k in ('x', 't', 'light'))
    with torch.no_grad():
        for shading, ratio in (('albedo', 1.0), ('albedo_normal', 1.0), ('lambertian', 0.3), ('textureless', 0.55), ('normal', 1.0)):
            sdf, sigma, color, normal, deform, raw = sc.forward(x, t, light, ratio=ratio, shading=shading)
            close(sdf, z[f'{shading}.sdf']); close(sigma, z[f'{shading}.sigma'], rtol=1e-4)
            shaded = shading in ('lambertian', 'textureless', 'normal')   # colour then depends on FD normals
            close(color, z[f'{shading}.color'], rtol=1e-3 if shaded else 2e-5, atol=2e-4 if shaded else 2e-6)
            close(deform, z[f'{shading}.deform'])
            if normal is not None:
                close(raw, z[f'{shading}.normal_raw'], rtol=1e-3, atol=2e-4)   # FD of fp32 sdf: cancellation
                close(normal, z[f'{shading}.normal'], rtol=1e-3, atol=2e-4)
        d = sc.density(x, t)
        close(d['sdf'], z['density.sdf']); close(d['albedo'], z['density.albedo'])
        d = sc.density(x, None)
        close(d['sdf'], z['density_cano.sdf']); close(d['albedo'], z['density_cano.albedo'])
        d = sc.density(x, t[:3], allow_shape=True, return_color=False)
        close(d['sigma'], z['density_allow_shape.sigma'], rtol=1e-4)
        n, raw = sc.normal(x, t=t)
        close(raw, z['normal_warped.raw'], rtol=1e-3, atol=2e-4)
        n, raw = sc.normal(x, topo=None)
        close(raw, z['normal_cano.raw'], rtol=1e-3, atol=2e-4)
        deform, topo = sc.warp(x, t)
        close(deform, z['warp.deform']); close(topo, z['warp.topo'])
        close(sc.code(t), z['code'], rtol=1e-6, atol=1e-6)
        close(sc.background(light, t), z['background'])
        o2, d2 = sc.pose_optimisation(x, light, torch.from_numpy(z['pose.ids']))
        close(o2, z['pose.o'], rtol=1e-6); close(d2, z['pose.d'], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('tag', CASES)
def test_gradients(golden_dir, tag):
    z, sd, ml = load(golden_dir, tag)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    sc = SceneOracle(sd, 1.01, 200, ml)
    x, t, light = (torch.from_numpy(z[k]) for k in ('x', 't', 'light'))
    M = x.shape[0]
    xg = x.clone().requires_grad_(True)
    sdf, sigma, color, normal, deform, raw = sc.forward(xg, t, light, ratio=1.0, shading='albedo_normal')
    g = torch.Generator().manual_seed(77)
    loss = (sdf * torch.randn(M, generator=g)).sum() + (sigma * torch.randn(M, generator=g)).sum() * 1e-2 \
        + (color * torch.randn(M, 3, generator=g)).sum() + (normal * torch.randn(M, 3, generator=g)).sum() \
        + (deform * torch.randn(M, 3, generator=g)).sum()
    loss.backward()

    def relerr(a, b):
        return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))
    assert relerr(xg.grad.numpy(), z['grad.x']) < 2e-3
    for name in ('encoder.embeddings', 'encoder_c.embeddings', 'sdf_net.net.0.weight', 'sdf_net.net.2.bias',
                 'color_net.net.1.weight_v', 'color_net.net.1.weight_g', 'deform_net.net.0.weight_v', 'deform_net.net.5.weight_g',
                 'topo_net.net.3.weight_v', 'deform_code.volumes.0', 'deform_code.volumes.2', 'sdf2density.beta'):
        assert relerr(sd[name].grad.numpy(), z['grad.' + name]) < 2e-3, name


@pytest.mark.parametrize('case', ['masked', 'nomask'])
def test_sdf_loss_restatements_vs_reference_golden(case):
    """utils.get_sdf_loss (utils.py:91-113): the oracle restatement (oracle/render.py) and the product's torch form
    (morpheus_b200.render.get_sdf_loss, which the GPU test pins the mb_sdf_loss_* kernels to) against values and gradients
    produced by the UNMODIFIED reference function (tests/golden/make_loss_golden.py)."""
    import numpy as np
    import torch
    from morpheus_b200 import render as mr
    from oracle import render as orr
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'loss_sdf.npz'))
    get = lambda k: torch.from_numpy(z[f'{case}_{k}'])      # noqa: E731
    mask = get('mask') if f'{case}_mask' in z.files else None
    for fn in (orr.get_sdf_loss, mr.get_sdf_loss):
        sdf = get('sdf').clone().requires_grad_(True)
        fs, sl = fn(get('z'), get('d'), sdf, 0.1, mask=mask)
        assert abs(float(fs) - float(get('fs'))) <= 1e-6 * max(1.0, abs(float(get('fs'))))
        assert abs(float(sl) - float(get('sl'))) <= 1e-6 * max(1.0, abs(float(get('sl'))))
        gfs, = torch.autograd.grad(fs, sdf, retain_graph=True)
        gsl, = torch.autograd.grad(sl, sdf)
        assert torch.allclose(gfs, get('gfs'), rtol=1e-5, atol=1e-9)
        assert torch.allclose(gsl, get('gsl'), rtol=1e-5, atol=1e-9)


def test_camera_rays_vs_reference_golden():
    """datasets/utils.py:28-65 get_camera_rays (OpenGL, un-normalised): product (morpheus_b200.render.get_camera_rays) and oracle
    (oracle.render.camera_dirs) against the unmodified reference function's output."""
    import numpy as np
    import torch
    from morpheus_b200 import render as mr
    from oracle import render as orr
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'camera_rays.npz'))
    H, W = int(z['H']), int(z['W'])
    ours = mr.get_camera_rays(H, W, float(z['fx']), float(z['fy']), float(z['cx']), float(z['cy']), device='cpu')
    assert torch.equal(ours, torch.from_numpy(z['dirs']))
    assert torch.equal(orr.camera_dirs(H, W, float(z['fx']), float(z['fy']), float(z['cx']), float(z['cy'])), torch.from_numpy(z['dirs']))
    assert torch.equal(mr.get_camera_rays(H, W, 20.0, device='cpu'), torch.from_numpy(z['dirs_default']))


def test_lookat_poses_vs_reference_golden():
    """datasets/dataset.py:225-266 get_c2w_from_cam_center (OpenGL, keep_chirality): morpheus_b200.rays.c2w_from_cam_center against
    the unmodified reference method's output."""
    import numpy as np
    import torch
    from morpheus_b200 import rays
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'lookat_poses.npz'))
    ours = rays.c2w_from_cam_center(torch.from_numpy(z['centers']), 0.0)
    assert torch.allclose(ours, torch.from_numpy(z['poses']), rtol=0, atol=1e-7)


def test_sds_view_angles_vs_reference_golden():
    """zero123_utils.py:102-120 angle_between (drives the SDS grad_scale, :123-134): product (guidance.Zero123.angle_between, radians)
    and oracle (oracle.sds.angle_between_deg) against the reference function's output."""
    import numpy as np
    import torch
    from morpheus_b200 import guidance
    from oracle import sds as osds
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'sds_angles.npz'))
    v1, v2, ref = torch.from_numpy(z['v1']), torch.from_numpy(z['v2']), torch.from_numpy(z['angles'])
    assert torch.allclose(guidance.Zero123.angle_between(v1, v2), ref, rtol=0, atol=2e-6)
    assert torch.allclose(osds.angle_between_deg(v1, v2), torch.rad2deg(ref), rtol=0, atol=2e-4)


def test_trainer_pieces_vs_reference_golden():
    """Pieces of the reference trainer executed from its own source text (tests/golden/make_loss_golden.py):
    get_real_view_render_loss (morpheus.py:946-983) vs train.real_view_loss_torch (the eager form the fused mb_ray_loss kernel is
    pinned to on the GPU), get_ortho_normal_dir (:518-528) vs Renderer.get_ortho_normal_dir, update_learning_rate (:471-502) vs
    train.learning_factor."""
    import types
    import numpy as np
    import torch
    from morpheus_b200 import train as mtrain
    from morpheus_b200.render import Renderer
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'trainer_pieces.npz'))
    t = lambda k: torch.from_numpy(z[k])      # noqa: E731
    # ---- render loss heads ----
    img, dep, opa = t('pred_rgb').clone().requires_grad_(True), t('pred_depth').clone().requires_grad_(True), t('pred_mask').clone().requires_grad_(True)
    out = {'image': img, 'depth': dep.reshape(-1), 'weights_sum': opa}
    batch = {'rgb': t('gt_rgb'), 'depth': t('gt_depth'), 'mask': t('gt_mask'), 'rays_o': t('rays_o').reshape(-1, 3), 'rays_d': t('rays_d').reshape(-1, 3)}
    tr = dict(mtrain.DEFAULT_TRAIN_CFG, beta_weight=0.0)
    model = types.SimpleNamespace(sdf2density=types.SimpleNamespace(get_beta=lambda: torch.zeros(())))
    loss = mtrain.real_view_loss_torch(out, batch, model, tr)
    assert abs(float(loss) - float(z['loss'])) < 1e-6 * max(1.0, abs(float(z['loss'])))
    g = torch.autograd.grad(loss, [img, dep, opa])
    assert torch.allclose(g[0], t('g_rgb'), rtol=1e-5, atol=1e-9)
    assert torch.allclose(g[1].reshape(-1), t('g_depth').reshape(-1), rtol=1e-5, atol=1e-9)
    assert torch.allclose(g[2].reshape(-1), t('g_mask').reshape(-1), rtol=1e-5, atol=1e-7)
    # ---- random tangent direction: same draw (torch.manual_seed(7); rand(50, 1) * 2 pi) ----
    torch.manual_seed(7)
    phi = torch.rand(50, 1) * 2.0 * np.pi
    w = Renderer.get_ortho_normal_dir(t('normals'), phi)
    assert torch.allclose(w, t('wdir'), rtol=0, atol=1e-6)
    # ---- learning-rate schedule: groups (encoder_sdf, encoder_color, decoder_sdf, density, code_deform, pose) ----
    for epoch, lrs in zip(z['lr_epochs'], z['lr_values']):
        lr = 5e-4 * mtrain.learning_factor(int(epoch), 200, 2000)
        assert np.allclose(lrs[:5], lr, rtol=1e-12, atol=0)
        assert np.isclose(lrs[5], lr * 0.1, rtol=1e-12, atol=0)


def test_normal_smoothness_loss_vs_reference_golden():
    """get_normal_smoothness_loss (morpheus.py:530-556) executed from the reference source (scene field = oracle) vs the oracle
    restatement oracle.render.normal_smoothness_loss with the reference's two RNG draws replayed from the same seed; the product's
    Renderer.get_normal_smoothness_loss is compared with the same formula on the GPU."""
    import numpy as np
    import torch
    from oracle import render as orr
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'normal_smoothness.npz'))
    sd = init_reference_like_state(200, seed=21, randomize=True, emb_scale=0.3, sphere=True)
    scene = SceneOracle({k: v.clone() for k, v in sd.items()}, 1.01, 200, 1.0)
    o = torch.from_numpy(z['rays_o']).clone().requires_grad_(True)
    dep = torch.from_numpy(z['depth']).clone().requires_grad_(True)
    N = o.shape[0]
    torch.manual_seed(int(z['seed']))
    trunc_noise = torch.rand(11)
    phi = torch.rand(11 * N, 1) * 2.0 * np.pi
    loss = orr.normal_smoothness_loss(scene, o, torch.from_numpy(z['rays_d']), torch.from_numpy(z['t']), dep, trunc_noise, phi)
    assert abs(float(loss) - float(z['loss'])) < 1e-5 * abs(float(z['loss']))
    g_o, g_dep = torch.autograd.grad(loss, [o, dep])
    assert np.linalg.norm(g_o.numpy() - z['g_o']) < 1e-3 * np.linalg.norm(z['g_o'])
    assert np.linalg.norm(g_dep.numpy() - z['g_depth']) < 1e-3 * np.linalg.norm(z['g_depth'])


def test_render_rays_glue_vs_reference_source_golden():
    """MorpheuS.render_rays (morpheus.py:558-794) executed from the reference source (stand-ins: oracle scene field, oracle
    compositing, injected samples; tests/golden/make_loss_golden.py) vs oracle.render.render_rays -- the checker every GPU render
    test and smoke() compare the product with -- with the reference's RNG draws (light offset, perturbation noise) replayed."""
    import numpy as np
    import torch
    from oracle import render as orr
    from oracle.fields import safe_normalize
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'render_rays_ref.npz'))
    sd = init_reference_like_state(200, seed=21, randomize=True, emb_scale=0.3, sphere=True)
    scene = SceneOracle({k: v.clone() for k, v in sd.items()}, 1.01, 200, 1.0)
    t = lambda k: torch.from_numpy(z[k])      # noqa: E731
    o, d = t('rays_o'), t('rays_d')
    N = o.shape[0]
    samples = (t('ray_indices'), t('t_starts'), t('t_ends'))
    M = samples[0].shape[0]
    rays_t = torch.full((1, N, 1), 31.0 / 200)
    ids = torch.full((1, N, 1), 31, dtype=torch.long)
    torch.manual_seed(int(z['seed']))
    loff = torch.randn(3)
    noise = torch.randn(M, 3)
    with torch.no_grad():
        o2, _ = scene.pose_optimisation(o, d, ids.view(-1, 1))
        out = orr.render_rays(scene, o[None], d[None], rays_t, ids, samples, bg_color=t('bg'), ambient_ratio=1.0, light_d=safe_normalize(o2 + loff),
                              shading='albedo_normal', optimize_pose=True, rays_depth=t('depth_gt')[None], rays_mask=t('mask_gt')[None],
                              perturb_noise=noise, trunc=0.1, smoothness_std=0.005, training=True, real_view=True)
        ts = rays_t.view(-1, 1)[:1]
        loss_code = torch.square(2 * scene.code(ts) - scene.code(ts - 1 / 200) - scene.code(ts + 1 / 200)).mean()

    def close(a, b, tol):
        a, b = np.asarray(a, dtype=np.float64).reshape(-1), np.asarray(b, dtype=np.float64).reshape(-1)
        return np.linalg.norm(a - b) <= tol * (np.linalg.norm(b) + 1e-12)
    assert close(out['image'].numpy(), z['image'], 1e-6)
    assert close(out['depth'].numpy(), z['depth'], 1e-6)
    assert close(out['weights'].numpy(), z['weights'], 1e-6)
    assert close(out['weights_sum'].numpy(), z['weights_sum'], 1e-6)
    assert close(out['sdf'].numpy(), z['sdf'], 1e-6)
    assert close(out['normal'].numpy(), z['normal'], 1e-5)
    assert close(out['normal_raw'].numpy(), z['normal_raw'], 1e-5)
    assert close(out['deform'].numpy(), z['deform'], 1e-6)
    assert close(out['loss_normal_perturb'].numpy(), z['loss_normal_perturb'], 1e-5)
    assert close(out['sdf_loss'].numpy(), z['sdf_loss'], 1e-6)
    assert close(out['fs_loss'].numpy(), z['fs_loss'], 1e-6)
    assert close(loss_code.numpy(), z['loss_code'], 1e-6)


def test_sds_scalar_chain_vs_reference_source_golden():
    """Zero123.train_step (zero123_utils.py:138-236) executed from the reference source with small stand-in networks
    (tests/golden/make_loss_golden.py) vs the oracle chain of oracle/sds.py assembled exactly as tests/test_sds_gpu.py assembles it
    to check morpheus_b200.guidance.Zero123.train_step on the GPU: loss, grad_scale and d(loss)/d(pred_rgb)."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    from oracle import sds as osds
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'sds_chain.npz'))
    t_ = lambda k: torch.from_numpy(z[k])      # noqa: E731

    def standin_encode(img256):
        p = F.avg_pool2d(img256 * 2 - 1, 8)
        return 0.18215 * torch.cat([p, p.mean(1, keepdim=True)], 1)

    def standin_apply(x_in, t_in, cc, ca):
        return torch.tanh(0.5 * x_in + 0.1 * cc + ca.mean(dim=(1, 2))[:, None, None, None] + 1e-3 * t_in[:, None, None, None].float())

    ac = osds.alphas_cumprod()
    pred = t_('pred').clone().requires_grad_(True)
    polar, azimuth, radius, t, noise = t_('polar'), t_('azimuth'), t_('radius'), t_('t'), t_('noise')
    lat = standin_encode(F.interpolate(pred, (256, 256), mode='bilinear', align_corners=False))
    ang = osds.angle_between_deg(torch.stack([radius + 2.5, torch.deg2rad(polar + 90.0), torch.deg2rad(azimuth + 0.0)], -1),
                                 torch.tensor([[2.5, np.deg2rad(90.0), 0.0]]))
    grad_scale = (torch.exp(ang.min(dim=1)[0] / 180.0) - 1) * 0.01
    with torch.no_grad():
        T = osds.pose_token(polar, azimuth, radius)
        clip = F.linear(torch.cat([t_('c_crossattn'), T], -1), t_('ccw'), t_('ccb'))
        x_in = torch.cat([osds.add_noise(lat.detach(), noise, t, ac)] * 2)
        eps = standin_apply(x_in, torch.cat([t, t]), torch.cat([torch.zeros(1, 4, 32, 32), t_('c_concat')]), torch.cat([torch.zeros_like(clip), clip]))
        grad = osds.sds_grad(eps[0:1], eps[1:2], noise, t, ac, 5.0, grad_scale)
    loss = osds.sds_loss(lat, grad)
    assert abs(float(grad_scale) - float(z['grad_scale'])) < 1e-6 * abs(float(z['grad_scale']))
    assert abs(float(loss) - float(z['loss'])) < 1e-5 * abs(float(z['loss']))
    g, = torch.autograd.grad(loss, pred)
    assert np.linalg.norm(g.numpy() - z['g_pred']) < 1e-5 * np.linalg.norm(z['g_pred'])


def test_real_view_total_loss_and_shipped_weights_vs_reference_source_golden():
    """The complete real-view loss of one iteration -- get_real_view_render_loss + get_real_view_point_loss + get_regularization_loss
    (morpheus.py:946-1029, 1090-1145) executed from the reference source with the weights of the shipped configs/snoopy.yaml --
    vs the product's loss assembly (train.real_view_loss_torch with FULL_TRAIN_CFG; the scene field is the oracle on both sides).
    Also pins every weight of train.FULL_TRAIN_CFG to the yaml."""
    import types
    import numpy as np
    import torch
    from morpheus_b200 import train as mtrain
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'real_view_total_loss.npz'))
    t_ = lambda k: torch.from_numpy(z[k])      # noqa: E731
    shipped = dict(zip([str(n) for n in z['weight_names']], [float(v) for v in z['weight_values']]))
    tr = dict(mtrain.FULL_TRAIN_CFG)
    for k in ('rgb_weight', 'mask_weight', 'depth_weight', 'sdf_weight', 'fs_weight', 'surf_sdf_weight', 'surf_color_weight', 'normal_smoothness',
              'normal_smooth_3d', 'smoothness_std', 'code_reg', 'beta_weight', 'ori_weight', 'trunc', 'lr'):
        assert abs(float(tr[k]) - shipped[k]) < 1e-12, k
    for k in ('normal_smooth_3d_t', 'normal_smooth_2d', 'eik_weight', 'sdf_reg', 'entropy_weight', 'deform_weight', 'deform_smooth', 'deform_smooth_t', 'topo_smooth_t'):
        assert shipped[k] == 0.0, f'{k} is active in the shipped config but not implemented'
    sd = init_reference_like_state(200, seed=21, randomize=True, emb_scale=0.3, sphere=True)
    scene = SceneOracle({k: v.clone() for k, v in sd.items()}, 1.01, 200, 1.0)
    model = types.SimpleNamespace(density=lambda x, t=None: scene.density(x, t=t), sdf2density=types.SimpleNamespace(get_beta=lambda: t_('beta')))
    out = {'image': t_('pred_rgb'), 'depth': t_('pred_depth'), 'weights_sum': t_('pred_mask'), 'sdf_loss': t_('o_sdf_loss'), 'fs_loss': t_('o_fs_loss'),
           'loss_normal_perturb': t_('o_loss_normal_perturb'), 'loss_code': t_('o_loss_code'), 'normal_reg': t_('o_normal_reg')}
    batch = {'rgb': t_('gt_rgb'), 'depth': t_('gt_depth'), 'mask': t_('gt_mask'), 'rays_o': t_('rays_o').reshape(-1, 3), 'rays_d': t_('rays_d').reshape(-1, 3),
             'rays_t': t_('rays_t').reshape(-1, 1)}
    with torch.no_grad():
        total = mtrain.real_view_loss_torch(out, batch, model, tr)
    assert abs(float(total) - float(z['total'])) < 2e-6 * abs(float(z['total'])), (float(total), float(z['total']))


def test_virtual_view_rays_vs_reference_golden():
    """DeformDataset.get_virtual_view_rays / get_virtual_view_data (datasets/dataset.py:435-578, shipped data config) vs
    morpheus_b200.rays.virtual_view_rays with the two angle draws injected."""
    import numpy as np
    import torch
    from morpheus_b200 import rays
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'virtual_views.npz'))
    for i in range(int(z['n'])):
        g = lambda k: z[f'v{i}_{k}']      # noqa: E731
        v = rays.virtual_view_rays(frame=12, num_frames=200, H=360, W=360, focal=517.0, scale=0.2, theta_deg=float(g('polar')[0]) + 90.0,
                                   phi_deg=float(g('azimuth')[0]))
        assert (v['H'], v['W']) == (int(g('H')), int(g('W')))
        assert np.allclose(v['rays_o'].numpy(), g('rays_o'), rtol=0, atol=2e-5)
        assert np.allclose(v['rays_d'].numpy(), g('rays_d'), rtol=0, atol=2e-5)
        assert np.allclose(v['rays_t'].numpy(), g('rays_t')) and np.array_equal(v['rays_id'].numpy(), g('rays_id'))
        assert np.allclose(v['polar'].numpy(), g('polar'), atol=1e-4) and np.allclose(v['azimuth'].numpy(), g('azimuth'), atol=1e-4)
        assert np.allclose(v['radius'].numpy(), g('radius'), atol=1e-6)


def test_real_view_sampling_vs_reference_golden():
    """DeformDataset.get_real_view_rays + sample_real_view_rays (datasets/dataset.py:336-433) vs the device-resident
    morpheus_b200.rays.RealViewData.sample_real_view_rays with the reference's two randint draws injected, and a full-frame fetch."""
    import numpy as np
    import torch
    from morpheus_b200 import rays
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'real_view_rays.npz'))
    data = rays.RealViewData(z['images'], z['depths'], z['masks'], z['poses'], z['K'])
    s = data.sample_real_view_rays(idx=torch.from_numpy(z['idx']), ray_num=13, index=torch.from_numpy(z['index']))
    assert (s['H'], s['W']) == (int(z['s_H']), int(z['s_W']))
    for k in ('rays_o', 'rays_d', 'rays_t', 'image', 'depth'):
        assert s[k].shape == z['s_' + k].shape and np.allclose(s[k].numpy(), z['s_' + k], rtol=0, atol=1e-6), k
    for k in ('rays_id', 'mask'):
        assert np.array_equal(s[k].numpy(), z['s_' + k]), k
    f = data.sample_real_view_rays(idx=2)
    for k in ('rays_o', 'rays_d', 'rays_t', 'image', 'depth'):
        assert f[k].shape == z['f_' + k].shape and np.allclose(f[k].numpy(), z['f_' + k], rtol=0, atol=1e-6), k
    for k in ('rays_id', 'mask'):
        assert np.array_equal(f[k].numpy(), z['f_' + k]), k

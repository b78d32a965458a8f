"""Generate tests/golden/loss_sdf.npz by running the UNMODIFIED reference `utils.get_sdf_loss`
(/root/reference/utils.py:91-113) on CPU with seeded inputs (values + autograd gradient w.r.t. the predicted SDF).

Run in the build container only (needs /root/reference):
    python tests/golden/make_loss_golden.py
One shim: `cv2` (imported at the top of the reference's utils.py, unused by get_sdf_loss) -> empty module.
"""
import os
import sys
import types

import numpy as np
import torch

sys.modules.setdefault('cv2', types.ModuleType('cv2'))
sys.path.insert(0, '/root/reference')
import utils as ref_utils  # noqa: E402  (reference, unmodified)

OUT = os.path.dirname(os.path.abspath(__file__))
g = torch.Generator().manual_seed(1234)
cases = {}
for name, M, with_mask in (('masked', 4000, True), ('nomask', 1500, False)):
    z = torch.rand(M, 1, generator=g) * 3.0
    d = torch.rand(M, 1, generator=g) * 2.5 + 0.2
    d[::7] = 0.0                      # no depth observation
    d[3::11] = -1.0                   # invalid depth marker (front-mask branch z < 3.5)
    near = torch.arange(0, M, 5)
    z[near] = d[near] + (torch.rand(near.shape[0], 1, generator=g) - 0.5) * 0.15      # samples inside / around the truncation band
    mask = (torch.rand(M, 1, generator=g) > 0.3).float() if with_mask else None
    sdf = (torch.randn(M, generator=g) * 0.2).requires_grad_(True)
    fs, sl = ref_utils.get_sdf_loss(z, d, sdf, 0.1, mask=mask)
    gfs, = torch.autograd.grad(fs, sdf, retain_graph=True)
    gsl, = torch.autograd.grad(sl, sdf)
    cases[name] = dict(z=z, d=d, sdf=sdf.detach(), fs=fs.detach(), sl=sl.detach(), gfs=gfs, gsl=gsl)
    if mask is not None:
        cases[name]['mask'] = mask
np.savez_compressed(os.path.join(OUT, 'loss_sdf.npz'), **{f'{c}_{k}': v.numpy() for c, vs in cases.items() for k, v in vs.items()})
print({c: (float(v['fs']), float(v['sl'])) for c, v in cases.items()})

# ---- pinhole ray directions: datasets/utils.py:28-65 (get_camera_rays, OpenGL convention, un-normalised) ----
import importlib.util  # noqa: E402
spec = importlib.util.spec_from_file_location('ref_dataset_utils', '/root/reference/datasets/utils.py')
ref_du = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_du)
H, W = 9, 12
dirs = ref_du.get_camera_rays(H, W, torch.tensor(517.0 / 30), torch.tensor(500.0 / 30), 6.3, 4.1)
dirs_default = ref_du.get_camera_rays(H, W, torch.tensor(20.0))
np.savez_compressed(os.path.join(OUT, 'camera_rays.npz'), dirs=dirs.numpy(), dirs_default=dirs_default.numpy(),
                    H=H, W=W, fx=517.0 / 30, fy=500.0 / 30, cx=6.3, cy=4.1)
print('camera rays', tuple(dirs.shape))

# ---- look-at poses of the virtual views: datasets/dataset.py:225-266 (get_c2w_from_cam_center, OpenGL, keep_chirality) ----
pkg = types.ModuleType('ref_datasets')      # the reference package is called `datasets` (clashes with the HuggingFace package): load it under an alias
pkg.__path__ = ['/root/reference/datasets']
sys.modules['ref_datasets'] = pkg
import importlib  # noqa: E402
ref_ds = importlib.import_module('ref_datasets.dataset')  # (reference, unmodified; cv2 stubbed above)
g2 = torch.Generator().manual_seed(77)
centers = torch.randn(16, 3, generator=g2) * 2.0
centers[0] = torch.tensor([0.0, 0.0, 2.5])
centers[1] = torch.tensor([2.5, 0.3, 0.0])
poses = ref_ds.DeformDataset.get_c2w_from_cam_center(None, centers, targets=0, camera_convention='OpenGL')
np.savez_compressed(os.path.join(OUT, 'lookat_poses.npz'), centers=centers.numpy(), poses=poses.numpy())
print('poses', tuple(poses.shape))

# ---- SDS view-angle weighting: models/guidance/zero123_utils.py:102-120 (angle_between), executed from the reference source text
#      (the module itself imports diffusers / omegaconf / ldm, absent offline) ----
import ast  # noqa: E402
src = open('/root/reference/models/guidance/zero123_utils.py').read()
fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == 'angle_between')
ns = {'torch': torch, 'np': np}
exec(compile(ast.Module(body=[fn], type_ignores=[]), 'zero123_utils.angle_between', 'exec'), ns)
g3 = torch.Generator().manual_seed(5)
v1 = torch.stack([2.5 + torch.rand(6, generator=g3) * 0.4, torch.deg2rad(90 + (torch.rand(6, generator=g3) - 0.5) * 90),
                  torch.deg2rad((torch.rand(6, generator=g3) - 0.5) * 360)], -1)
v2 = torch.tensor([[2.5, np.deg2rad(90.0), 0.0], [2.5, np.deg2rad(80.0), np.deg2rad(120.0)]], dtype=torch.float32)
angles = ns['angle_between'](None, v1, v2)
np.savez_compressed(os.path.join(OUT, 'sds_angles.npz'), v1=v1.numpy(), v2=v2.numpy(), angles=angles.numpy())
print('angles', tuple(angles.shape))

# ---- trainer-side pieces of morpheus.py, executed from the reference source text (the module imports nerfacc / the Zero-1-to-3
#      stack, absent offline): get_real_view_render_loss (:946-983), get_ortho_normal_dir (:518-528), update_learning_rate (:471-502) ----
import torch.nn.functional as F  # noqa: E402
msrc = open('/root/reference/morpheus.py').read()
mtree = ast.parse(msrc)


def ref_method(name):
    f = next(n for n in ast.walk(mtree) if isinstance(n, ast.FunctionDef) and n.name == name)
    env = {'torch': torch, 'np': np, 'F': F, 'print': lambda *a, **k: None}
    exec(compile(ast.Module(body=[f], type_ignores=[]), f'morpheus.{name}', 'exec'), env)
    return env[name]


cfg_train = {'rgb_weight': 5.0, 'mask_weight': 0.5, 'depth_weight': 0.1, 'warm_up_end': 200, 'n_epochs': 2000, 'lr': 5e-4}
fake = types.SimpleNamespace(config={'train': cfg_train})
g4 = torch.Generator().manual_seed(99)
N = 301
pred_rgb = torch.rand(N, 3, generator=g4).requires_grad_(True)
pred_depth = (torch.rand(N, 1, generator=g4) * 3).requires_grad_(True)
pred_mask = torch.rand(N, 1, generator=g4)
pred_mask[:4] = 0.0
pred_mask[4:8] = 1.0
pred_mask.requires_grad_(True)
gt_rgb = torch.rand(N, 3, generator=g4)
gt_depth = torch.rand(N, generator=g4) * 2
gt_depth[::5] = 0.0
gt_mask = (torch.rand(N, generator=g4) > 0.4).float()
rays_o = torch.randn(1, N, 3, generator=g4) * 0.3
rays_d = torch.randn(1, N, 3, generator=g4) * 0.3
# shapes as the trainer passes them for a real view (B = 1, H = N rays, W = 1; morpheus.py:915-944)
loss = ref_method('get_real_view_render_loss')(fake, pred_rgb.t().reshape(1, 3, N, 1), pred_depth.reshape(1, 1, N, 1), pred_mask.reshape(1, 1, N, 1),
                                               gt_rgb.t().reshape(1, 3, N, 1), gt_depth.reshape(1, N, 1), gt_mask.reshape(1, N, 1), rays_o, rays_d)
grads = torch.autograd.grad(loss, [pred_rgb, pred_depth, pred_mask])
torch.manual_seed(7)
normals = torch.randn(50, 3, generator=g4)
normals[0] = torch.tensor([0.0, 0.0, 1.0])
wdir = ref_method('get_ortho_normal_dir')(fake, normals)
lrs = {}
for epoch in (0, 99, 100, 150, 199, 200, 1100, 2000):
    groups = [{'name': n, 'lr': -1.0} for n in ('encoder_sdf', 'encoder_color', 'decoder_sdf', 'density', 'code_deform', 'pose')]
    fk = types.SimpleNamespace(config={'train': cfg_train}, epoch=epoch, optimizer=types.SimpleNamespace(param_groups=groups))
    ref_method('update_learning_rate')(fk)
    lrs[epoch] = [g['lr'] for g in groups]
np.savez_compressed(os.path.join(OUT, 'trainer_pieces.npz'), pred_rgb=pred_rgb.detach().numpy(), pred_depth=pred_depth.detach().numpy(),
                    pred_mask=pred_mask.detach().numpy(), gt_rgb=gt_rgb.numpy(), gt_depth=gt_depth.numpy(), gt_mask=gt_mask.numpy(),
                    rays_o=rays_o.numpy(), rays_d=rays_d.numpy(), loss=loss.detach().numpy(), g_rgb=grads[0].numpy(), g_depth=grads[1].numpy(),
                    g_mask=grads[2].numpy(), normals=normals.numpy(), wdir=wdir.numpy(),
                    lr_epochs=np.array(sorted(lrs)), lr_values=np.array([lrs[e] for e in sorted(lrs)]))
print('trainer pieces: loss', float(loss))

# ---- get_normal_smoothness_loss (morpheus.py:530-556) executed from the reference source, with the scene field provided by the
#      oracle (oracle.fields.SceneOracle.normal; the field itself is pinned by tests/golden/scene_*.npz).  Inputs keep every band
#      point inside the 1.1 sphere, so the reference's boolean indexing keeps all 11 N points and its two RNG draws
#      (rand_like(trunc_normal), then rand([11 N, 1]) in get_ortho_normal_dir) can be replayed from the same seed. ----
sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
from oracle.fields import SceneOracle, init_reference_like_state  # noqa: E402
sd = init_reference_like_state(200, seed=21, randomize=True, emb_scale=0.3, sphere=True)
scene = SceneOracle({k: v.clone() for k, v in sd.items()}, 1.01, 200, 1.0)
g5 = torch.Generator().manual_seed(31)
Nn = 10
o5 = (torch.randn(Nn, 3, generator=g5) * 0.1).requires_grad_(True)
d5 = torch.nn.functional.normalize(torch.randn(Nn, 3, generator=g5), dim=-1) * 0.4
dep5 = (torch.rand(1, Nn, generator=g5) * 1.2 + 0.2).requires_grad_(True)
t5 = torch.full((Nn, 1), 31.0 / 200)
fake5 = types.SimpleNamespace(config={'train': {'trunc': 0.1, 'smoothness_std': 0.005}}, model=types.SimpleNamespace(normal=lambda x, t=None: scene.normal(x, t=t)))
fake5.get_ortho_normal_dir = lambda normals: ref_method('get_ortho_normal_dir')(fake5, normals)
torch.manual_seed(123)
reg = ref_method('get_normal_smoothness_loss')(fake5, o5, d5, t5, dep5)
g_o5, g_dep5 = torch.autograd.grad(reg, [o5, dep5])
np.savez_compressed(os.path.join(OUT, 'normal_smoothness.npz'), rays_o=o5.detach().numpy(), rays_d=d5.numpy(), depth=dep5.detach().numpy(), t=t5.numpy(),
                    loss=reg.detach().numpy(), g_o=g_o5.numpy(), g_depth=g_dep5.numpy(), seed=123)
print('normal smoothness', float(reg))

# ---- MorpheuS.render_rays (morpheus.py:558-794) executed from the reference source on the CPU.  Stand-ins: the scene field is the
#      oracle (pinned by scene_*.npz), nerfacc's two compositing calls are the oracle restatements (nerfacc is not installable
#      offline: that boundary stays unpinned), the sampler returns an injected fixed-S lattice; get_sdf_loss / safe_normalize are the
#      reference's own.  This pins the GLUE of render_rays (ray points, light, blend, perturbed-normal loss, code regulariser, SDF
#      loss wiring) that oracle.render.render_rays restates and every GPU render test checks the product against. ----
from oracle import render as orr  # noqa: E402
fake_nerfacc = types.SimpleNamespace(
    render_weight_from_density=lambda t0, t1, sig, ray_indices=None, n_rays=None: orr.render_weight_from_density(t0, t1, sig, ray_indices, n_rays),
    accumulate_along_rays=lambda w, values=None, ray_indices=None, n_rays=None: orr.accumulate_along_rays(w, values, ray_indices, n_rays))


class _Model:
    training = True

    def __call__(self, x, t, light, ratio=1, shading='albedo', cano=False):
        return scene.forward(x, t, light, ratio=ratio, shading=shading, cano=cano)

    def normal(self, x, t=None, cano=False, topo=None):
        return scene.normal(x, t=t, cano=cano, topo=topo)

    def get_deform_code(self, t):
        return scene.code(t)

    def pose_optimisation(self, o, d, ids):
        return scene.pose_optimisation(o, d, ids)


fr = next(n for n in ast.walk(mtree) if isinstance(n, ast.FunctionDef) and n.name == 'render_rays')
env = {'torch': torch, 'np': np, 'nerfacc': fake_nerfacc, 'safe_normalize': ref_du.safe_normalize, 'get_sdf_loss': ref_utils.get_sdf_loss}
exec(compile(ast.Module(body=[fr], type_ignores=[]), 'morpheus.render_rays', 'exec'), env)
g6 = torch.Generator().manual_seed(8)
Nr, S = 24, 16
c2w = orr.look_at_pose(70.0, 30.0, 2.5)
dirs = orr.camera_dirs(360, 360, 517.0, 517.0, 180.0, 180.0).reshape(-1, 3)
idx = torch.randint(0, dirs.shape[0], (Nr,), generator=g6)
o6, d6 = orr.rays_from_pose(dirs[idx], c2w)
aabb = torch.tensor([-1.01, -1.01, -1.01, 1.01, 1.01, 1.01])
samples = orr.sample_uniform(o6, d6, aabb, S, torch.rand(Nr, generator=g6))
t6 = torch.full((1, Nr, 1), 31.0 / 200)
id6 = torch.full((1, Nr, 1), 31, dtype=torch.long)
bg6 = torch.rand(Nr, 3, generator=g6)
depth6 = torch.rand(Nr, generator=g6) * 1.5 + 1.5
depth6[::4] = 0.0
mask6 = (torch.rand(Nr, generator=g6) > 0.3).float()
train_cfg = {'ori_weight': 0.01, 'normal_smooth_3d': 0.1, 'normal_dir': False, 'smoothness_std': 0.005, 'topo_none': True, 'normal_smooth_3d_t': 0.0,
             'deform_smooth': 0.0, 'deform_smooth_t': 0.0, 'topo_smooth_t': 0.0, 'code_reg': 0.5, 'normal_smooth_2d': 0.0, 'normal_smoothness': 0.0, 'trunc': 0.1}
fake6 = types.SimpleNamespace(config={'model': {'bg_radius': 1.4}, 'render': {'step_size': 0.01}, 'train': train_cfg}, model=_Model(),
                              occupancy_grid=types.SimpleNamespace(sampling=lambda *a, **k: samples), dataset=types.SimpleNamespace(num_frames=200))
torch.manual_seed(2024)
res = env['render_rays'](fake6, o6[None], d6[None], t6, id6, Nr, 1, bg_color=bg6, ambient_ratio=1.0, shading='albedo_normal', real_view=True,
                         rays_depth=depth6[None], rays_mask=mask6[None], optimize_pose=True)
np.savez_compressed(os.path.join(OUT, 'render_rays_ref.npz'), rays_o=o6.numpy(), rays_d=d6.numpy(), ray_indices=samples[0].numpy(), t_starts=samples[1].numpy(),
                    t_ends=samples[2].numpy(), bg=bg6.numpy(), depth_gt=depth6.numpy(), mask_gt=mask6.numpy(), seed=2024,
                    **{k: v.detach().numpy() for k, v in res.items() if torch.is_tensor(v)})
print('render_rays keys', sorted(k for k, v in res.items() if torch.is_tensor(v)))

# ---- Zero123.train_step (models/guidance/zero123_utils.py:138-236) executed from the reference source with small deterministic
#      stand-ins for the three networks (encode_imgs, cc_projection, apply_model: the real ones are pinned separately by
#      sds_nets.npz against the reference UNetModel / Encoder classes) and the closed-form DDIM add_noise (diffusers is absent).
#      Pins the SCALAR CHAIN the oracle (oracle/sds.py) restates: angle-based grad scale, pose token, CFG combination, per-view
#      weights, w(t), nan_to_num, loss. ----
from pathlib import Path  # noqa: E402
ftrain = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == 'train_step')
env2 = {'torch': torch, 'np': np, 'F': F, 'Path': Path, 'save_image': None}
exec(compile(ast.Module(body=[ftrain], type_ignores=[]), 'zero123_utils.train_step', 'exec'), env2)


def standin_encode(img256):                  # [1,3,256,256] in [0,1] -> [1,4,32,32]
    p = F.avg_pool2d(img256 * 2 - 1, 8)
    return 0.18215 * torch.cat([p, p.mean(1, keepdim=True)], 1)


def standin_apply(x_in, t_in, cond):         # 'hybrid' conditioning: x_in [2,4,32,32], c_concat [2,4,32,32], c_crossattn [2,1,768]
    cc, ca = cond['c_concat'][0], cond['c_crossattn'][0]
    return torch.tanh(0.5 * x_in + 0.1 * cc + ca.mean(dim=(1, 2))[:, None, None, None] + 1e-3 * t_in[:, None, None, None].float())


betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
ac = torch.cumprod(1.0 - betas, 0)
g7 = torch.Generator().manual_seed(11)
ccw, ccb = torch.randn(768, 772, generator=g7) * 0.02, torch.randn(768, generator=g7) * 0.01
fake7 = types.SimpleNamespace(
    device='cpu', alphas=ac, min_step=20, max_step=500,
    angle_between=lambda a, b: ns['angle_between'](None, a, b),
    encode_imgs=standin_encode,
    scheduler=types.SimpleNamespace(add_noise=lambda z, e, t: (ac[t] ** 0.5).view(-1, 1, 1, 1) * z + ((1 - ac[t]) ** 0.5).view(-1, 1, 1, 1) * e),
    model=types.SimpleNamespace(cc_projection=lambda x: F.linear(x, ccw, ccb), apply_model=standin_apply))
emb = {'c_crossattn': [torch.randn(1, 1, 768, generator=g7)], 'c_concat': [torch.randn(1, 4, 32, 32, generator=g7)],
       'ref_radii': [2.5], 'ref_polars': [90.0], 'ref_azimuths': [0.0], 'zero123_ws': [1]}
pred = torch.rand(1, 3, 72, 72, generator=g7).requires_grad_(True)
polar, azimuth, radius = torch.tensor([10.0]), torch.tensor([200.0]), torch.tensor([0.1])
t7 = torch.tensor([260])
noise7 = torch.randn(1, 4, 32, 32, generator=g7)
loss7, t_out, gs7, _ = env2['train_step'](fake7, emb, pred, polar, azimuth.clone(), radius, guidance_scale=5, grad_scale=0.01, t=t7, noise=noise7)
gpred, = torch.autograd.grad(loss7, pred)
np.savez_compressed(os.path.join(OUT, 'sds_chain.npz'), ccw=ccw.numpy(), ccb=ccb.numpy(), c_crossattn=emb['c_crossattn'][0].numpy(), c_concat=emb['c_concat'][0].numpy(),
                    pred=pred.detach().numpy(), polar=polar.numpy(), azimuth=azimuth.numpy(), radius=radius.numpy(), t=t7.numpy(), noise=noise7.numpy(),
                    loss=loss7.detach().numpy(), grad_scale=gs7.numpy(), g_pred=gpred.numpy())
print('sds chain: loss', float(loss7), 'grad_scale', float(gs7))

# ---- the complete real-view loss of one iteration (morpheus.py:1202-1236): get_real_view_render_loss + get_real_view_point_loss
#      (:985-1029) + get_regularization_loss (:1090-1145), executed from the reference source with the weights of the SHIPPED
#      configs/snoopy.yaml and the oracle scene as `model.density`. ----
import yaml  # noqa: E402
ytrain = yaml.safe_load(open('/root/reference/configs/snoopy.yaml'))['train']
g8 = torch.Generator().manual_seed(321)
N8 = 37
sc8 = scene
o8 = torch.randn(1, N8, 3, generator=g8) * 0.15
d8 = torch.nn.functional.normalize(torch.randn(1, N8, 3, generator=g8), dim=-1) * 0.5
t8 = torch.full((1, N8, 1), 77.0 / 200)
gd8 = torch.rand(N8, generator=g8) * 1.2 + 0.1
gd8[::6] = 0.0
gm8 = (torch.rand(N8, generator=g8) > 0.35).float()
grgb8 = torch.rand(N8, 3, generator=g8)
prgb8, pdep8, pmask8 = torch.rand(N8, 3, generator=g8), torch.rand(N8, generator=g8) * 2, torch.rand(N8, generator=g8)
outs8 = {'sdf_loss': torch.rand((), generator=g8), 'fs_loss': torch.rand((), generator=g8), 'loss_normal_perturb': torch.rand((), generator=g8),
         'loss_code': torch.rand((), generator=g8), 'normal_reg': torch.rand((), generator=g8), 'normal_raw': torch.randn(N8 * 4, 3, generator=g8),
         'deform': torch.randn(N8 * 4, 3, generator=g8), 'weights': torch.rand(N8 * 4, generator=g8)}
beta8 = torch.tensor(0.1).abs() + 1e-4
fake8 = types.SimpleNamespace(config={'train': ytrain, 'exp': {'end_iter': 1}}, global_step=0,
                              model=types.SimpleNamespace(density=lambda x, t=None: sc8.density(x, t=t), sdf2density=types.SimpleNamespace(get_beta=lambda: beta8)))
with torch.no_grad():
    l_render = ref_method('get_real_view_render_loss')(fake8, prgb8.t().reshape(1, 3, N8, 1), pdep8.reshape(1, 1, N8, 1), pmask8.reshape(1, 1, N8, 1),
                                                       grgb8.t().reshape(1, 3, N8, 1), gd8.reshape(1, N8, 1), gm8.reshape(1, N8, 1), o8, d8)
    l_point = ref_method('get_real_view_point_loss')(fake8, grgb8.t().reshape(1, 3, N8, 1), gd8.reshape(1, N8, 1), gm8.reshape(1, N8, 1), o8, d8, t8, outs8)
    l_reg = ref_method('get_regularization_loss')(fake8, outs8, None, cano=False)
total8 = l_render + l_point + l_reg
np.savez_compressed(os.path.join(OUT, 'real_view_total_loss.npz'), rays_o=o8.numpy(), rays_d=d8.numpy(), rays_t=t8.numpy(), gt_depth=gd8.numpy(), gt_mask=gm8.numpy(),
                    gt_rgb=grgb8.numpy(), pred_rgb=prgb8.numpy(), pred_depth=pdep8.numpy(), pred_mask=pmask8.numpy(), beta=beta8.numpy(),
                    l_render=l_render.numpy(), l_point=l_point.numpy(), l_reg=l_reg.numpy(), total=total8.numpy(),
                    weight_names=np.array(sorted(k for k, v in ytrain.items() if isinstance(v, (int, float)) and not isinstance(v, bool))),
                    weight_values=np.array([float(ytrain[k]) for k in sorted(k for k, v in ytrain.items() if isinstance(v, (int, float)) and not isinstance(v, bool))]),
                    **{'o_' + k: v.numpy() for k, v in outs8.items()})
print('real-view total loss', float(total8), float(l_render), float(l_point), float(l_reg))

# ---- novel-view ray generation: DeformDataset.get_virtual_view_data / get_virtual_view_rays (datasets/dataset.py:435-578) called
#      unbound on a stand-in dataset object carrying the shipped `data` config and the synthetic snoopy-shaped camera ----
ycfg = yaml.safe_load(open('/root/reference/configs/snoopy.yaml'))
DD = ref_ds.DeformDataset
Fv = 200
K = torch.eye(4)
K[0, 0] = K[1, 1] = 517.0
K[0, 2] = K[1, 2] = 180.0
fake9 = types.SimpleNamespace(cfg=ycfg, num_frames=Fv, H=360, W=360, intrinsics=K, is_train=True, radius=torch.full((Fv,), 2.5),
                              real_view_data={'theta': torch.full((Fv,), 90.0), 'phi': torch.zeros(Fv), 'radius': torch.full((Fv,), 2.5)})
for name in ('get_radius', 'scale_intrinsics', 'get_c2w_from_cam_center', 'get_virtual_view_data'):
    setattr(fake9, name, types.MethodType(getattr(DD, name), fake9))
views = []
for seed in (1, 2, 3):
    torch.manual_seed(seed)
    data = DD.get_virtual_view_rays(fake9, bs=1, t=12)
    views.append(data)
np.savez_compressed(os.path.join(OUT, 'virtual_views.npz'), n=len(views),
                    **{f'v{i}_{k}': (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for i, dv in enumerate(views) for k, v in dv.items() if k != 'dir'})
print('virtual views', [(float(v['polar']), float(v['azimuth'])) for v in views])

# ---- real-view ray tables + sampling: DeformDataset.get_real_view_rays / sample_real_view_rays (datasets/dataset.py:336-433) on a
#      small synthetic RGB-D sequence (known_view_scale 1: cv2.resize to the same size is the identity, provided by the cv2 stub) ----
sys.modules['cv2'].resize = lambda img, size, interpolation=None: img
sys.modules['cv2'].INTER_LINEAR, sys.modules['cv2'].INTER_NEAREST = 1, 0
g10 = torch.Generator().manual_seed(404)
Fr, Hr, Wr = 5, 6, 8
imgs = torch.rand(Fr, Hr, Wr, 3, generator=g10).numpy()
deps = (torch.rand(Fr, Hr, Wr, generator=g10) * 3).numpy()
msks = (torch.rand(Fr, Hr, Wr, generator=g10) > 0.5).numpy().astype(np.int64)
cent = torch.randn(Fr, 3, generator=g10) * 2.0
poses_r = DD.get_c2w_from_cam_center(None, cent, targets=0, camera_convention='OpenGL')
Kr = torch.eye(4)
Kr[0, 0], Kr[1, 1], Kr[0, 2], Kr[1, 2] = 11.5, 10.5, 4.2, 2.9
fake10 = types.SimpleNamespace(cfg=ycfg, num_frames=Fr, H=Hr, W=Wr, intrinsics=Kr, poses=poses_r, images=imgs, depths=deps, masks=msks,
                               theta=torch.rand(Fr, generator=g10), phi=torch.rand(Fr, generator=g10), radius=torch.rand(Fr, generator=g10))
fake10.scale_intrinsics = types.MethodType(DD.scale_intrinsics, fake10)
fake10.real_view_data = DD.get_real_view_rays(fake10)
torch.manual_seed(9)
smp = DD.sample_real_view_rays(fake10, ray_num=13)
torch.manual_seed(9)
idx_draw = torch.randint(0, Fr, (1,))
index_draw = torch.randint(0, Hr * Wr, (13,))
full = DD.sample_real_view_rays(fake10, idx=2)
np.savez_compressed(os.path.join(OUT, 'real_view_rays.npz'), images=imgs, depths=deps, masks=msks, poses=poses_r.numpy(), K=Kr.numpy(),
                    idx=idx_draw.numpy(), index=index_draw.numpy(),
                    **{'s_' + k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in smp.items()},
                    **{'f_' + k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in full.items() if k in ('rays_o', 'rays_d', 'rays_t', 'rays_id', 'image', 'depth', 'mask')})
print('real-view rays', {k: tuple(v.shape) for k, v in smp.items() if torch.is_tensor(v)})

# ---- per-iteration host choices: get_shading (morpheus.py:865-888), get_bg_color (:890-903), progressive_level (:808-813),
#      get_gt_from_data blend (:929-944), executed from the reference source with the shipped weights; python / torch RNG seeded ----
import random as pyrandom  # noqa: E402
shade = []
fake11 = types.SimpleNamespace(config={'train': ytrain, 'model': {'bg_radius': 1.4}}, device='cpu', model=types.SimpleNamespace(max_level=None))
for m_name in ('get_shading', 'get_bg_color', 'progressive_level', 'get_gt_from_data'):
    env_m = {'torch': torch, 'np': np, 'random': pyrandom}
    f_m = next(n for n in ast.walk(mtree) if isinstance(n, ast.FunctionDef) and n.name == m_name)
    exec(compile(ast.Module(body=[f_m], type_ignores=[]), f'morpheus.{m_name}', 'exec'), env_m)
    setattr(fake11, m_name, types.MethodType(env_m[m_name], fake11))
pyrandom.seed(5)
for ratio in (0.0, 0.05, 0.1, 0.3, 0.6, 0.99):
    for rv in (True, False):
        a, sname = fake11.get_shading(ratio, rv)
        shade.append((ratio, float(rv), a, {'albedo': 0, 'albedo_normal': 1, 'lambertian': 2, 'textureless': 3}[sname]))
levels = []
for ratio in (0.0, 0.25, 0.5, 1.0, 1.5):
    fake11.progressive_level(ratio)
    levels.append((ratio, fake11.model.max_level))
pyrandom.seed(6)
torch.manual_seed(6)
bg_real = fake11.get_bg_color(True, B=1, N=9)
bg_virtual = [fake11.get_bg_color(False) for _ in range(6)]
g12 = torch.Generator().manual_seed(12)
img12 = torch.rand(1, 3, 9, 1, generator=g12)
dep12 = torch.rand(1, 9, 1, generator=g12)
msk12 = torch.rand(1, 9, 1, generator=g12)
gt_rgb12, gt_dep12, gt_msk12 = fake11.get_gt_from_data({'image': img12.clone(), 'depth': dep12.clone(), 'mask': msk12.clone()}, bg_real, 1, 9, 1)
np.savez_compressed(os.path.join(OUT, 'host_choices.npz'), shade=np.array(shade), levels=np.array(levels), bg_real=bg_real.numpy(),
                    bg_virtual_none=np.array([b is None for b in bg_virtual]), bg_virtual=np.stack([b.numpy() if b is not None else np.zeros(3, np.float32) for b in bg_virtual]),
                    img=img12.numpy(), mask=msk12.numpy(), gt_rgb=gt_rgb12.numpy(), gt_mask=gt_msk12.numpy(),
                    albedo_iter_ratio=ytrain['albedo_iter_ratio'], min_ambient_ratio=ytrain['min_ambient_ratio'], textureless_ratio=ytrain['textureless_ratio'])
print('host choices', len(shade), levels[-1])

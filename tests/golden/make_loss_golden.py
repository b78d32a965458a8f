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

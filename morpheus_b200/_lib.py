"""ctypes binding of the C-ABI library (include/morpheus_b200.h).

The product path has NO fallback: if libmorpheus_b200.so is missing or a call fails, a
RuntimeError is raised (the reference raises RuntimeError from TORCH_CHECK / std::runtime_error,
gridencoder.cu:15-18,392).  PyTorch is used only for device memory and streams: tensors cross
the boundary as raw device pointers (`tensor.data_ptr()`), never as torch types.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmorpheus_b200.so')

_lib = None
# tensor-core (tcgen05) forward engine; MORPHEUS_B200_TC=0 selects the fp32 SIMT engine (kept as the parity reference)
USE_TC = os.environ.get('MORPHEUS_B200_TC', '1') != '0'
# tensor-core backward of the deform/topology nets (needs the forward activation stash, i.e. USE_TC)
USE_TC_BWD = os.environ.get('MORPHEUS_B200_TC_BWD', '1') != '0'
# tensor-core backward of the SDF / colour nets + FD queries (recompute on tcgen05, per-tile TMEM weight-gradient accumulators)
USE_TC_BWD_SDF = os.environ.get('MORPHEUS_B200_TC_BWD_SDF', '1') != '0'
# FD-normal queries of the backward in the specialised two-CTAs-per-SM kernel (csrc/field_bwd_fd_tc.cu)
USE_TC_BWD_FD = os.environ.get('MORPHEUS_B200_TC_BWD_FD', '1') != '0'
# real-view FD-normal regulariser (both FD sets of a sample, forward + backward) in ONE launch (csrc/field_fd_reg_tc.cu)
USE_FD_REG = os.environ.get('MORPHEUS_B200_FD_REG', '1') != '0'


class LayerDesc(C.Structure):
    _fields_ = [('wt_off', C.c_uint32), ('w_off', C.c_uint32), ('b_off', C.c_uint32),
                ('K', C.c_uint32), ('N', C.c_uint32), ('K_pad', C.c_uint32), ('N_pad', C.c_uint32)]


class FieldParams(C.Structure):
    _fields_ = [('arena', C.c_void_p),
                ('deform', LayerDesc * 6), ('topo', LayerDesc * 6), ('sdf', LayerDesc * 3), ('color', LayerDesc * 3),
                ('emb_sdf', C.c_void_p), ('emb_col', C.c_void_p), ('offsets', C.c_void_p),
                ('code', C.c_void_p * 3), ('code_len', C.c_uint32 * 3),
                ('beta', C.c_void_p), ('bound', C.c_float), ('two_bound', C.c_float),
                ('S', C.c_float), ('H', C.c_uint32), ('n_levels', C.c_uint32), ('n_freq', C.c_uint32)]


class FieldIO(C.Structure):
    _fields_ = [('M', C.c_uint32), ('flags', C.c_uint32), ('shading', C.c_int), ('ratio', C.c_float),
                ('x', C.c_void_p), ('t', C.c_void_p), ('light', C.c_void_p), ('topo_in', C.c_void_p),
                ('sdf', C.c_void_p), ('sigma', C.c_void_p), ('color', C.c_void_p), ('normal', C.c_void_p),
                ('normal_raw', C.c_void_p), ('deform', C.c_void_p), ('topo', C.c_void_p)]


class FieldGrads(C.Structure):
    _fields_ = [('g_sdf', C.c_void_p), ('g_sigma', C.c_void_p), ('g_color', C.c_void_p), ('g_normal', C.c_void_p), ('g_normal_raw', C.c_void_p),
                ('g_deform', C.c_void_p), ('g_topo', C.c_void_p),
                ('deform', C.c_void_p), ('topo', C.c_void_p), ('normal_raw', C.c_void_p),
                ('g_arena', C.c_void_p), ('g_emb_sdf', C.c_void_p), ('g_emb_col', C.c_void_p), ('g_code', C.c_void_p * 3),
                ('g_beta', C.c_void_p), ('g_x', C.c_void_p), ('g_topo_in', C.c_void_p), ('g_def_out', C.c_void_p), ('g_topo_out', C.c_void_p),
                ('g_fd', C.c_void_p)]


F_WARP, F_MAIN, F_COLOR, F_FD, F_FD_WARPED, F_TOPO_IN, F_SKIP_WARP_BWD, F_FD_DELEGATE = 1, 2, 4, 8, 16, 32, 64, 128
SHADE = {'albedo': 0, 'lambertian': 1, 'albedo_normal': 1, 'textureless': 2, 'normal': 3}

# every symbol include/morpheus_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = ['mb_version', 'mb_last_error', 'mb_sm_count', 'mb_grid_encode_forward', 'mb_grid_encode_backward',
           'mb_sample_rays_count', 'mb_sample_rays_write', 'mb_sample_rays_uniform', 'mb_composite_forward',
           'mb_composite_backward', 'mb_field_forward', 'mb_field_backward', 'mb_occ_update', 'mb_occ_binarize', 'mb_occ_binarize_dev',
           'mb_adam_step', 'mb_sds_grad', 'mb_add_noise', 'mb_sds_grad_dev', 'mb_add_noise_dev', 'mb_conv_pack_weights', 'mb_nchw_split', 'mb_conv_tc', 'mb_rows_split', 'mb_pack_tc', 'mb_field_forward_tc', 'mb_adam_step_dev', 'mb_adam_step_groups', 'mb_field_backward_warp_tc', 'mb_field_backward_sdf_tc',
           'mb_ray_points_forward', 'mb_ray_points_backward', 'mb_pack_arena_forward', 'mb_pack_arena_backward', 'mb_sdf_loss_forward',
           'mb_sdf_loss_backward', 'mb_pose_rays_forward', 'mb_pose_rays_backward', 'mb_ray_loss', 'mb_field_backward_fd_tc', 'mb_fd_regulariser_tc', 'mb_debug_fd_phases', 'mb_debug_fdr_phases', 'mb_code_reg']


def lib():
    """Load the shared library (once).  Fails loudly: there is no CPU / eager fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} is missing: run `python -m morpheus_b200.build` (or __graft_entry__.build()). '
                               'morpheus_b200 has no CPU fallback.')
        L = C.CDLL(LIB_PATH)
        L.mb_last_error.restype = C.c_char_p
        for s in SYMBOLS:
            if s != 'mb_last_error':
                getattr(L, s).restype = C.c_int
        _lib = L
    return _lib


def check(rc, what=''):
    if rc != 0:
        msg = lib().mb_last_error().decode('utf-8', 'replace')
        raise RuntimeError(f'morpheus_b200 {what} failed (code {rc}): {msg}')


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  The tensor must be CUDA + contiguous, like the
    reference's CHECK_CUDA / CHECK_CONTIGUOUS (gridencoder.cu:468-478)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('morpheus_b200: tensor must be a CUDA tensor')
    if not t.is_contiguous():
        raise RuntimeError('morpheus_b200: tensor must be contiguous')
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Profile:
    """Optional CUDA-event timing of our own kernel launches (bench.py): events are recorded on the stream the
    kernel is launched on (torch's current stream), durations are read after a synchronize."""

    def __init__(self):
        self.enabled = False
        self.events = []

    def reset(self):
        self.events = []

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, a, b in self.events:
            d = out.setdefault(name, {'n': 0, 'total_ms': 0.0})
            d['n'] += 1
            d['total_ms'] += a.elapsed_time(b)
        for d in out.values():
            d['avg_ms'] = d['total_ms'] / d['n']
        return out


PROFILE = _Profile()


class timed:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if PROFILE.enabled:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if PROFILE.enabled:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            PROFILE.events.append((self.name, self.a, b))
        return False

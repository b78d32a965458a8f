"""Drop-in for /root/reference/external/encoders/gridencoder/grid.py.

`_backend.grid_encode_forward/backward` keep the reference extension's exact positional
signature (src/bindings.cpp:6-7, gridencoder.h:12-13) so the reference `grid.py` can import this
module's `_backend` unchanged; `GridEncoder` keeps the reference constructor,
parameter/buffer names (`embeddings`, `offsets`) and forward(inputs, bound, max_level) contract
(grid.py:103-169).  Compute is the sm_100a kernel in csrc/grid_encode.cu via the C ABI.
"""
import math

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib
from ._lib import check, ptr, stream

_gridtype_to_id = {'hash': 0, 'tiled': 1}
_interp_to_id = {'linear': 0, 'smoothstep': 1}


def _dtype_id(t):
    if t.dtype != torch.float32:
        raise RuntimeError('morpheus_b200 grid encoder: only float32 embeddings are supported '
                           '(MorpheuS runs with fp16: False, configs/snoopy.yaml:28)')
    return 0


class _Backend:
    """Same call surface as the reference `_gridencoder` extension module."""

    @staticmethod
    def grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, max_level, S, H, dy_dx, gridtype,
                            align_corners, interp):
        if inputs.dtype != torch.float32:
            raise RuntimeError('inputs must be a float32 tensor')
        if offsets.dtype != torch.int32:
            raise RuntimeError('offsets must be an int tensor')
        check(_lib.lib().mb_grid_encode_forward(
            ptr(inputs), ptr(embeddings), ptr(offsets), ptr(outputs), int(B), int(D), int(C), int(L), int(max_level),
            _lib.C.c_float(float(S)), int(H), ptr(dy_dx), int(gridtype), int(bool(align_corners)), int(interp),
            _dtype_id(embeddings), stream()), 'grid_encode_forward')

    @staticmethod
    def grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, max_level, S, H, dy_dx,
                             grad_inputs, gridtype, align_corners, interp):
        if offsets.dtype != torch.int32:
            raise RuntimeError('offsets must be an int tensor')
        check(_lib.lib().mb_grid_encode_backward(
            ptr(grad), ptr(inputs), ptr(embeddings), ptr(offsets), ptr(grad_embeddings), int(B), int(D), int(C), int(L),
            int(max_level), _lib.C.c_float(float(S)), int(H), ptr(dy_dx), ptr(grad_inputs), int(gridtype),
            int(bool(align_corners)), int(interp), _dtype_id(grad), stream()), 'grid_encode_backward')

    @staticmethod
    def grad_total_variation(*a, **k):
        raise NotImplementedError('grad_total_variation has no call site in MorpheuS (SURVEY.md 2b); not built')

    @staticmethod
    def grad_weight_decay(*a, **k):
        raise NotImplementedError('grad_weight_decay has no call site in MorpheuS (SURVEY.md 2b); not built')


_backend = _Backend()


class _grid_encode(Function):
    """grid.py:25-96: allocation, [L,B,C] layout, max_level handling and zero fills stay host-side."""

    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
                align_corners=False, interpolation=0, max_level=None):
        inputs = inputs.contiguous()
        B, D = inputs.shape
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        S = np.log2(per_level_scale)
        H = base_resolution
        max_level = L if max_level is None else max(min(int(math.ceil(max_level * L)), L), 1)
        outputs = torch.empty(L, B, C, device=inputs.device, dtype=embeddings.dtype)
        if max_level < L:
            outputs.zero_()
        dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=embeddings.dtype) if calc_grad_inputs else None
        _backend.grid_encode_forward(inputs, embeddings.contiguous(), offsets, outputs, B, D, C, L, max_level, S, H, dy_dx,
                                     gridtype, align_corners, interpolation)
        outputs = outputs.permute(1, 0, 2).reshape(B, L * C)
        ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
        ctx.dims = [B, D, C, L, S, H, gridtype, interpolation, max_level]
        ctx.align_corners = align_corners
        return outputs

    @staticmethod
    def backward(ctx, grad):
        inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, gridtype, interpolation, max_level = ctx.dims
        grad = grad.view(B, L, C).permute(1, 0, 2).contiguous()
        grad_embeddings = torch.zeros_like(embeddings)
        grad_inputs = torch.zeros_like(inputs, dtype=embeddings.dtype) if dy_dx is not None else None
        _backend.grid_encode_backward(grad, inputs, embeddings.contiguous(), offsets, grad_embeddings, B, D, C, L, max_level, S, H,
                                      dy_dx, grad_inputs, gridtype, ctx.align_corners, interpolation)
        if grad_inputs is not None:
            grad_inputs = grad_inputs.to(inputs.dtype)
        return grad_inputs, grad_embeddings, None, None, None, None, None, None, None, None


grid_encode = _grid_encode.apply


class GridEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16, log2_hashmap_size=19,
                 desired_resolution=None, gridtype='hash', align_corners=False, interpolation='linear'):
        super().__init__()
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = _gridtype_to_id[gridtype]
        self.interpolation = interpolation
        self.interp_id = _interp_to_id[interpolation]
        self.align_corners = align_corners
        # table sizing rule of grid.py:125-136 (float64 ceil, round up to a multiple of 8)
        offsets, offset = [], 0
        self.max_params = 2 ** log2_hashmap_size
        for i in range(num_levels):
            resolution = int(np.ceil(base_resolution * per_level_scale ** i))
            params_in_level = min(self.max_params, resolution ** input_dim)
            params_in_level = int(np.ceil(params_in_level / 8) * 8)
            offsets.append(offset)
            offset += params_in_level
        offsets.append(offset)
        self.register_buffer('offsets', torch.from_numpy(np.array(offsets, dtype=np.int32)))
        self.n_params = offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(offset, level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        self.embeddings.data.uniform_(-1e-4, 1e-4)   # grid.py:145-147

    def __repr__(self):
        return (f'GridEncoder(morpheus_b200): input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} '
                f'resolution={self.base_resolution} per_level_scale={self.per_level_scale:.4f} params={tuple(self.embeddings.shape)} '
                f'gridtype={self.gridtype} align_corners={self.align_corners} interpolation={self.interpolation}')

    def forward(self, inputs, bound=1, max_level=None):
        inputs = (inputs + bound) / (2 * bound)
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        outputs = grid_encode(inputs, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                              inputs.requires_grad, self.gridtype_id, self.align_corners, self.interp_id, max_level)
        return outputs.view(prefix_shape + [self.output_dim])

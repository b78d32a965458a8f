"""Oracle: multi-resolution hash/tiled grid encoder (TEST INFRASTRUCTURE, CPU only).

Restates, in numpy float32/uint32 arithmetic, the reference CUDA kernels
  external/encoders/gridencoder/src/gridencoder.cu
    :46-58   fast_hash            (uint32 wrap-around multiply, xor)
    :61-79   get_grid_index       (dense while stride <= hashmap_size, else hash; % size)
    :83-249  kernel_grid          (forward + dy_dx)
    :253-349 kernel_grid_backward (scatter of w*grad)
    :353-378 kernel_input_backward
and the Python glue external/encoders/gridencoder/grid.py:25-96 (layout [L,B,C], max_level,
zero fill) and :103-169 (offset table sizing, [-bound,bound] -> [0,1] mapping).

Floating-point contraction: nvcc fuses ``a*b+c`` into one FMA; numpy does not.  ``_fma`` emulates
the fused op through float64 (exact 48-bit product, one extra rounding that differs from a true
FMA only in vanishingly rare double-rounding ties), so oracle and kernel normally agree bit for
bit; tests still allow a few ulp.
"""
import math

import numpy as np

PRIMES = (np.uint32(1), np.uint32(2654435761), np.uint32(805459861))  # gridencoder.cu:49
F32 = np.float32


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F32)


def per_level_scale(base_resolution=16, desired_resolution=128, num_levels=16):
    """grid.py:108-109"""
    return float(np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1)))


def make_offsets(input_dim=3, num_levels=16, base_resolution=16, scale=None, log2_hashmap_size=15):
    """grid.py:125-136 -- note: uses float64 np.ceil (33/65/129 at levels 5/10/15), which only
    sizes the tables; the kernel's own float32 resolution rule is `level_resolution`."""
    if scale is None:
        scale = per_level_scale(base_resolution, 128, num_levels)
    max_params = 2 ** log2_hashmap_size
    offsets, offset = [], 0
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * scale ** i))
        params_in_level = min(max_params, resolution ** input_dim)
        params_in_level = int(np.ceil(params_in_level / 8) * 8)
        offsets.append(offset)
        offset += params_in_level
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32)


def level_resolution(level, S, H):
    """gridencoder.cu:133  (uint32_t)ceil(exp2f(level * S) * H), all in float32."""
    v = np.exp2(F32(level) * F32(S), dtype=F32) * F32(H)
    return int(np.ceil(F32(v)))


def resolve_max_level(max_level, L):
    """grid.py:42"""
    return L if max_level is None else max(min(int(math.ceil(max_level * L)), L), 1)


def _grid_index(gridtype, hashmap_size, res, coords):
    """gridencoder.cu:61-79.  coords: list of D uint32 arrays.  Returns entry index (not *C)."""
    D = len(coords)
    stride = 1
    index = np.zeros_like(coords[0], dtype=np.uint32)
    d = 0
    while d < D and stride <= hashmap_size:
        index = (index + coords[d] * np.uint32(stride & 0xFFFFFFFF)).astype(np.uint32)
        stride *= res
        d += 1
    if gridtype == 0 and stride > hashmap_size:
        index = np.zeros_like(index)
        for i in range(D):
            index ^= (coords[i] * PRIMES[i]).astype(np.uint32)
    return (index % np.uint32(hashmap_size)).astype(np.int64)


def _locate(x, res, align_corners):
    """gridencoder.cu:140-151 -> (pos fractional [B,D] f32, pos_grid [B,D] u32)."""
    r = F32(res)
    if align_corners:
        pos = (x * F32(res - 1)).astype(F32)
        pg = np.minimum(np.floor(pos).astype(np.uint32), np.uint32(res - 2))
    else:
        pos = _fma(x, np.full_like(x, r), np.full_like(x, F32(-0.5)))
        pos = np.minimum(np.maximum(pos, F32(0.0)), F32(res - 1)).astype(F32)
        pg = np.floor(pos).astype(np.uint32)
    pos = (pos - pg.astype(F32)).astype(F32)
    return pos, pg


def grid_encode_forward(inputs, embeddings, offsets, max_level, S, H, calc_dydx=False,
                        gridtype=0, align_corners=False, interp=0):
    """kernel_grid, gridencoder.cu:83-249.  inputs [B,D] f32 in [0,1]; embeddings [sO,C] f32.
    Returns outputs [L,B,C] (levels >= max_level are zero, grid.py:53) and dy_dx [B,L*D*C] or None."""
    inputs = np.ascontiguousarray(inputs, dtype=F32)
    emb = np.ascontiguousarray(embeddings, dtype=F32)
    B, D = inputs.shape
    C = emb.shape[1]
    L = len(offsets) - 1
    assert interp == 0, "oracle restates interp=linear only (the only mode MorpheuS requests, models/model.py:149)"
    out = np.zeros((L, B, C), dtype=F32)
    dydx = np.zeros((B, L, D, C), dtype=F32) if calc_dydx else None
    oob = np.any((inputs < 0) | (inputs > 1), axis=1)  # :106-112
    ok = ~oob
    x = inputs[ok]
    for l in range(max_level):
        hs = int(offsets[l + 1] - offsets[l])
        res = level_resolution(l, S, H)
        table = emb[offsets[l]:offsets[l + 1]]
        pos, pg = _locate(x, res, align_corners)
        one_m = (F32(1.0) - pos).astype(F32)
        acc = np.zeros((x.shape[0], C), dtype=F32)
        for idx in range(1 << D):  # :171-195
            w = np.ones(x.shape[0], dtype=F32)
            coords = []
            for d in range(D):
                if (idx >> d) & 1 == 0:
                    w = (w * one_m[:, d]).astype(F32)
                    coords.append(pg[:, d])
                else:
                    w = (w * pos[:, d]).astype(F32)
                    coords.append(np.minimum(pg[:, d] + np.uint32(1), np.uint32(res - 1)))
            gi = _grid_index(gridtype, hs, res, coords)
            acc = _fma(w[:, None], table[gi], acc)
        out[l, ok] = acc
        if calc_dydx:  # :205-248
            scale = F32(res - 1) if align_corners else F32(res)
            for gd in range(D):
                g = np.zeros((x.shape[0], C), dtype=F32)
                for idx in range(1 << (D - 1)):
                    w = np.full(x.shape[0], scale, dtype=F32)
                    coords = [None] * D
                    for nd in range(D - 1):
                        d = nd + 1 if nd >= gd else nd
                        if (idx >> nd) & 1 == 0:
                            w = (w * one_m[:, d]).astype(F32)
                            coords[d] = pg[:, d]
                        else:
                            w = (w * pos[:, d]).astype(F32)
                            coords[d] = np.minimum(pg[:, d] + np.uint32(1), np.uint32(res - 1))
                    coords[gd] = pg[:, gd]
                    il = _grid_index(gridtype, hs, res, coords)
                    coords[gd] = np.minimum(pg[:, gd] + np.uint32(1), np.uint32(res - 1))
                    ir = _grid_index(gridtype, hs, res, coords)
                    diff = (table[ir] - table[il]).astype(F32)
                    # acc += (w*diff)*pos_deriv with pos_deriv == 1.0f: one rounding for the
                    # product, one for the add (the fma with 1.0 is a plain add)
                    g = (g + (w[:, None] * diff).astype(F32)).astype(F32)
                dydx[ok, l, gd] = g
    if calc_dydx:
        dydx = dydx.reshape(B, L * D * C)
    return out, dydx


def grid_encode_backward(grad, inputs, embeddings, offsets, max_level, S, H, dy_dx=None,
                         gridtype=0, align_corners=False, interp=0):
    """kernel_grid_backward + kernel_input_backward, gridencoder.cu:253-378.
    grad [L,B,C].  Returns grad_embeddings [sO,C] (float64-accumulated then cast: the reference's
    atomics have no defined order) and grad_inputs [B,D] or None."""
    inputs = np.ascontiguousarray(inputs, dtype=F32)
    B, D = inputs.shape
    C = embeddings.shape[1]
    L = len(offsets) - 1
    ge = np.zeros(embeddings.shape, dtype=np.float64)
    ok = ~np.any((inputs < 0) | (inputs > 1), axis=1)
    x = inputs[ok]
    for l in range(max_level):
        hs = int(offsets[l + 1] - offsets[l])
        res = level_resolution(l, S, H)
        pos, pg = _locate(x, res, align_corners)
        one_m = (F32(1.0) - pos).astype(F32)
        g = grad[l][ok].astype(F32)
        for idx in range(1 << D):
            w = np.ones(x.shape[0], dtype=F32)
            coords = []
            for d in range(D):
                if (idx >> d) & 1 == 0:
                    w = (w * one_m[:, d]).astype(F32)
                    coords.append(pg[:, d])
                else:
                    w = (w * pos[:, d]).astype(F32)
                    coords.append(np.minimum(pg[:, d] + np.uint32(1), np.uint32(res - 1)))
            gi = _grid_index(gridtype, hs, res, coords) + int(offsets[l])
            contrib = (w[:, None] * g).astype(F32).astype(np.float64)
            for ch in range(C):
                np.add.at(ge[:, ch], gi, contrib[:, ch])
    grad_inputs = None
    if dy_dx is not None:
        dd = dy_dx.reshape(B, L, D, C).astype(np.float64)
        gg = np.transpose(grad, (1, 0, 2)).astype(np.float64)  # [B,L,C]
        grad_inputs = np.einsum('blc,bldc->bd', gg, dd).astype(F32)
    return ge.astype(F32), grad_inputs


# ---------------------------------------------------------------------------------------------
# torch wrappers (autograd) mirroring grid.py so the unchanged reference model can sit on top.
# ---------------------------------------------------------------------------------------------
import torch  # noqa: E402


class _GridEncodeOracle(torch.autograd.Function):
    """grid.py:25-96 (`_grid_encode`) on top of the numpy restatement."""

    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, S, H, calc_grad_inputs, gridtype, align_corners, interp, max_level):
        L = offsets.shape[0] - 1
        ml = resolve_max_level(max_level, L)
        B = inputs.shape[0]
        C = embeddings.shape[1]
        out, dydx = grid_encode_forward(inputs.detach().cpu().numpy(), embeddings.detach().cpu().numpy(),
                                        offsets.cpu().numpy(), ml, S, H, calc_grad_inputs, gridtype, align_corners, interp)
        ctx.save_for_backward(inputs, embeddings, offsets)
        ctx.dydx = dydx
        ctx.cfg = (ml, S, H, gridtype, align_corners, interp)
        return torch.from_numpy(out).permute(1, 0, 2).reshape(B, L * C)

    @staticmethod
    def backward(ctx, grad):
        inputs, embeddings, offsets = ctx.saved_tensors
        ml, S, H, gridtype, align_corners, interp = ctx.cfg
        B = inputs.shape[0]
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        g = grad.detach().reshape(B, L, C).permute(1, 0, 2).contiguous().cpu().numpy()
        ge, gi = grid_encode_backward(g, inputs.detach().cpu().numpy(), embeddings.detach().cpu().numpy(),
                                      offsets.cpu().numpy(), ml, S, H, ctx.dydx, gridtype, align_corners, interp)
        gi_t = torch.from_numpy(gi) if gi is not None else None
        return gi_t, torch.from_numpy(ge), None, None, None, None, None, None, None, None


class GridEncoderOracle(torch.nn.Module):
    """grid.py:103-169 (`GridEncoder`) -- same constructor arguments, parameter name
    (`embeddings`), buffer name (`offsets`) and forward(inputs, bound, max_level) contract."""

    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype='hash', align_corners=False,
                 interpolation='linear'):
        super().__init__()
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim, self.num_levels, self.level_dim = input_dim, num_levels, level_dim
        self.per_level_scale, self.base_resolution = per_level_scale, base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype_id = {'hash': 0, 'tiled': 1}[gridtype]
        self.interp_id = {'linear': 0, 'smoothstep': 1}[interpolation]
        self.align_corners = align_corners
        offsets = make_offsets(input_dim, num_levels, base_resolution, per_level_scale, log2_hashmap_size)
        self.register_buffer('offsets', torch.from_numpy(offsets))
        self.embeddings = torch.nn.Parameter(torch.empty(int(offsets[-1]), level_dim).uniform_(-1e-4, 1e-4))

    def forward(self, inputs, bound=1, max_level=None):
        inputs = (inputs + bound) / (2 * bound)  # grid.py:157
        prefix = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        S = float(np.log2(self.per_level_scale))
        out = _GridEncodeOracle.apply(inputs, self.embeddings, self.offsets, S, self.base_resolution,
                                      inputs.requires_grad, self.gridtype_id, self.align_corners, self.interp_id, max_level)
        return out.view(prefix + [self.output_dim])

"""B200-native `scene_representation` -- host-side mirror of /root/reference/models/model.py:31.

Same constructor, attribute, method and state_dict surface as the reference class (SURVEY.md
8b: `forward/__call__`, `density`, `normal`, `warp`, `get_topo`, `get_sigma_albedo`,
`get_deform_code`, `background`, `pose_optimisation`, `get_RT`, `get_params_all`, `.max_level`,
`.sdf2density.get_beta()`, `.pose_array`; keys `encoder.embeddings`, `sdf_net.net.N.weight`,
`deform_net.net.N.weight_g/_v`, `deform_code.volumes.i`, `sdf2density.beta`, `pose_array.data`),
but every per-sample query runs in ONE fused CUDA launch (csrc/field_fwd_tc.cu, tcgen05; csrc/field_fwd.cu is the fp32 SIMT
parity engine) and its backward in one to three more (csrc/field_bwd_sdf_tc.cu, field_bwd_fd_tc.cu, field_bwd_tc.cu) instead of
~300 eager kernels with [M, .] intermediates in HBM; the real-view FD-normal regulariser is one forward+backward launch
(csrc/field_fd_reg_tc.cu).
The sub-modules below only own parameters; their torch `forward`s exist for the tiny side paths
(background colour, code regulariser) that the reference also runs outside the hot loop.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, packing
from ._lib import F_COLOR, F_FD, F_FD_WARPED, F_MAIN, F_TOPO_IN, F_WARP, SHADE, check, ptr, stream
from .gridencoder import GridEncoder

PROFILE = _lib.PROFILE
_TC_TABLES = {}


_TC_TABLES_T = {}
_TC_TABLES_S = {}
_TC_TABLES_ALL = {}


def _tc_tables_small(dev):
    key = str(dev)
    if key not in _TC_TABLES_S:
        _TC_TABLES_S[key] = packing.tc_tables_dgrad_small(dev)
    return _TC_TABLES_S[key]


def _tc_tables_dgrad(dev):
    key = str(dev)
    if key not in _TC_TABLES_T:
        _TC_TABLES_T[key] = packing.tc_tables_dgrad(dev)
    return _TC_TABLES_T[key]


def _tc_tables(dev):
    key = str(dev)
    if key not in _TC_TABLES:
        _TC_TABLES[key] = packing.tc_tables(dev)
    return _TC_TABLES[key]


def safe_normalize(x, eps=1e-20):
    """utils.py:70-71"""
    return x / torch.sqrt(torch.clamp(torch.sum(x * x, -1, keepdim=True), min=eps))


# ------------------------------------------------------------------------------------------------
# parameter containers (names/shapes of SURVEY.md Appendix B)
# ------------------------------------------------------------------------------------------------
class MLP(nn.Module):
    """models/decoders.py:9-64: Linear stack, ReLU between, optional geometric init, weight_norm default."""

    def __init__(self, dim_in, dim_out, dim_hidden, num_layers, bias=True, geo_init=False, inside_outside=False,
                 geo_bias=0.5, weight_norm=True, bias_init=None):
        super().__init__()
        self.dim_in, self.dim_out, self.dim_hidden, self.num_layers = dim_in, dim_out, dim_hidden, num_layers
        self.uses_weight_norm = weight_norm
        net = []
        for l in range(num_layers):
            d_in = dim_in if l == 0 else dim_hidden
            d_out = dim_out if l == num_layers - 1 else dim_hidden
            lin = nn.Linear(d_in, d_out, bias=bias)
            if geo_init:
                with torch.no_grad():
                    if l == num_layers - 1:
                        sign = -1.0 if inside_outside else 1.0
                        lin.weight.normal_(mean=sign * np.sqrt(np.pi) / np.sqrt(d_in), std=0.0001)
                        lin.bias.fill_(-sign * geo_bias)
                    elif l == 0:
                        lin.bias.zero_()
                        lin.weight[:, 3:].zero_()
                        lin.weight[:, :3].normal_(0.0, np.sqrt(2) / np.sqrt(d_out))
                    else:
                        lin.bias.zero_()
                        lin.weight.normal_(0.0, np.sqrt(2) / np.sqrt(d_out))
            if bias_init is not None and l == num_layers - 1:
                nn.init.constant_(lin.bias, -bias_init)
            if weight_norm:
                lin = nn.utils.weight_norm(lin)
            net.append(lin)
        self.net = nn.ModuleList(net)

    def effective(self):
        """[(W [out,in], b [out])] with weight_norm applied: W = g * v / ||v||_row."""
        out = []
        for lin in self.net:
            if self.uses_weight_norm:
                v, g = lin.weight_v, lin.weight_g
                W = v * (g / v.norm(dim=1, keepdim=True))
            else:
                W = lin.weight
            out.append((W, lin.bias))
        return out

    def forward(self, x):
        for l, (W, b) in enumerate(self.effective()):
            x = F.linear(x, W, b)
            if l != self.num_layers - 1:
                x = F.relu(x)
        return x


class MultiCode(nn.Module):
    """models/deform_code.py:5-42: three learnable 1-D feature lines sampled at t."""

    def __init__(self, sizes, c):
        super().__init__()
        self.volumes = nn.ParameterList([nn.Parameter(torch.randn((1, c, size, 1))) for size in sizes])

    def sample(self, t):
        t = t.clamp(0, 1).reshape(-1)
        feats = []
        for v in self.volumes:
            line = v[0, :, :, 0]
            S = line.shape[1]
            pos = ((t * 2 - 1) + 1) / 2 * (S - 1)
            i0 = torch.floor(pos).long().clamp(0, S - 1)
            i1 = (i0 + 1).clamp(0, S - 1)
            w1 = pos - i0.to(pos.dtype)
            valid = ((i0 + 1) <= S - 1).to(pos.dtype)
            feats.append((line[:, i0] * (1 - w1) + line[:, i1] * w1 * valid).t())
        return torch.cat(feats, dim=-1)

    def get_code(self, level=-1):
        return self.volumes[level].squeeze().permute(1, 0)


class LaplaceDensity(nn.Module):
    """models/density.py:17-31."""

    def __init__(self, params_init=None, beta_min=0.0001):
        super().__init__()
        for k, v in (params_init or {}).items():
            setattr(self, k, nn.Parameter(torch.tensor(v)))
        self.beta_min = beta_min

    def get_beta(self):
        return self.beta.abs() + self.beta_min

    def density_func(self, sdf, beta=None):
        if beta is None:
            beta = self.get_beta()
        return (1 / beta) * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() / beta))

    def forward(self, sdf, beta=None):
        return self.density_func(sdf, beta=beta)


class PoseArray(nn.Module):
    """models/pose.py:4-64."""

    def __init__(self, num_frames):
        super().__init__()
        self.num_frames = num_frames
        self.num_params = 6
        self.data = nn.Parameter(torch.zeros([num_frames, 6], dtype=torch.float32))

    def forward(self, ids):
        return self.data[ids]

    def get_translations(self, ids):
        tr = self.data[:, 3:6][ids]
        return tr[None, ...] if tr.dim() == 1 else tr

    def get_rotations(self, ids):
        return self.data[:, 0:3][ids]

    def get_rotation_matrices(self, ids):
        r = self.get_rotations(ids)
        if r.dim() == 1:
            r = r[None, ...]
        ca, cb, cg = torch.cos(r[:, 0]), torch.cos(r[:, 1]), torch.cos(r[:, 2])
        sa, sb, sg = torch.sin(r[:, 0]), torch.sin(r[:, 1]), torch.sin(r[:, 2])
        c1 = torch.stack([ca * cb, sa * cb, -sb], -1)
        c2 = torch.stack([ca * sb * sg - sa * cg, sa * sb * sg + ca * cg, cb * sg], -1)
        c3 = torch.stack([ca * sb * cg + sa * sg, sa * sb * cg - ca * sg, cb * cg], -1)
        return torch.stack([c1, c2, c3], -1)


class FreqEncoder_torch(nn.Module):
    """models/encodings.py:10-57 (used only off the hot path: background colour)."""

    def __init__(self, input_dim, max_freq_log2, N_freqs, **_):
        super().__init__()
        self.input_dim, self.N_freqs = input_dim, N_freqs
        self.output_dim = input_dim + input_dim * N_freqs * 2
        self.freq_bands = [float(2 ** k) for k in np.linspace(0, max_freq_log2, N_freqs)]

    def forward(self, x, max_level=None, **kwargs):
        n_on = self.N_freqs if max_level is None else int(max_level * self.N_freqs)
        out = [x]
        for k in range(n_on):
            out += [torch.sin(x * self.freq_bands[k]), torch.cos(x * self.freq_bands[k])]
        if self.N_freqs - n_on > 0:
            out.append(torch.zeros(*x.shape[:-1], (self.N_freqs - n_on) * 2 * x.shape[-1], device=x.device, dtype=x.dtype))
        return torch.cat(out, dim=-1)


def get_encoder(encoding, input_dim=3, multires=6, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                desired_resolution=2048, align_corners=False, interpolation='linear', **kwargs):
    """models/encodings.py:59-90 restricted to the encoders MorpheuS instantiates (models/model.py:101-167)."""
    if encoding == 'frequency_torch':
        enc = FreqEncoder_torch(input_dim=input_dim, max_freq_log2=multires - 1, N_freqs=multires)
    elif encoding in ('hashgrid', 'tiledgrid'):
        enc = GridEncoder(input_dim=input_dim, num_levels=num_levels, level_dim=level_dim, base_resolution=base_resolution,
                          log2_hashmap_size=log2_hashmap_size, desired_resolution=desired_resolution,
                          gridtype='hash' if encoding == 'hashgrid' else 'tiled', align_corners=align_corners, interpolation=interpolation)
    else:
        raise NotImplementedError(f'encoding {encoding!r}: MorpheuS only builds frequency_torch / hashgrid (SURVEY.md 2, rows 7 and 16)')
    return enc, enc.output_dim


class _PoseRays(torch.autograd.Function):
    """scene_representation.pose_optimisation (models/model.py:335-346) in one launch (and one backward launch) instead of
    ~35 eager ops + their autograd nodes; same operation order as the torch expression (bit-identical forward)."""

    @staticmethod
    def forward(ctx, pose, rays_o, rays_d, frame_ids):
        rays_o, rays_d = rays_o.contiguous().float(), rays_d.contiguous().float()
        frame_ids = frame_ids.contiguous().long()
        N = rays_o.shape[0]
        o2, d2 = torch.empty_like(rays_o), torch.empty_like(rays_d)
        with _lib.timed('pose_rays_fwd'):
            check(_lib.lib().mb_pose_rays_forward(ptr(pose.detach()), ptr(frame_ids), ptr(rays_o), ptr(rays_d), N, ptr(o2), ptr(d2), stream()),
                  'pose_rays_forward')
        ctx.save_for_backward(pose, rays_d, frame_ids)
        ctx.set_materialize_grads(False)
        return o2, d2

    @staticmethod
    def backward(ctx, g_o2, g_d2):
        pose, rays_d, frame_ids = ctx.saved_tensors
        N = rays_d.shape[0]
        g_pose = torch.zeros_like(pose)
        g_o = torch.empty_like(rays_d) if ctx.needs_input_grad[1] else None
        g_d = torch.empty_like(rays_d) if ctx.needs_input_grad[2] else None
        g_o2 = g_o2.contiguous().float() if g_o2 is not None else None
        g_d2 = g_d2.contiguous().float() if g_d2 is not None else None
        with _lib.timed('pose_rays_bwd'):
            check(_lib.lib().mb_pose_rays_backward(ptr(pose.detach()), ptr(frame_ids), ptr(rays_d), ptr(g_o2), ptr(g_d2), N, ptr(g_pose), ptr(g_o), ptr(g_d),
                                                   stream()), 'pose_rays_backward')
        return g_pose, g_o, g_d, None


class _CodeReg(torch.autograd.Function):
    """loss_code of render_rays (morpheus.py:762-771): mean (2 c(t) - c(t - 1/F) - c(t + 1/F))^2 over the 48 code channels in one
    launch (and one backward launch) instead of ~80 eager ops (three MultiCode.sample calls with index_put backward sorts)."""

    @staticmethod
    def forward(ctx, t_dev, inv_frames, sinks, v0, v1, v2):
        vols = [v.detach() for v in (v0, v1, v2)]
        out = torch.empty(1, device=v0.device, dtype=torch.float32)
        codes = (_lib.C.c_void_p * 3)(*[v.data_ptr() for v in vols])
        lens = (_lib.C.c_int * 3)(*[int(v.shape[2]) for v in vols])
        with _lib.timed('code_reg_fwd'):
            check(_lib.lib().mb_code_reg(codes, lens, ptr(t_dev), _lib.C.c_float(inv_frames), ptr(out), None, None, stream()), 'code_reg')
        ctx.save_for_backward(t_dev, v0, v1, v2)
        ctx.inv_frames, ctx.sinks = inv_frames, sinks
        return out[0]

    @staticmethod
    def backward(ctx, g):
        t_dev, v0, v1, v2 = ctx.saved_tensors
        vols = [v0, v1, v2]
        grads = ctx.sinks if ctx.sinks is not None else [torch.zeros_like(v) for v in vols]
        codes = (_lib.C.c_void_p * 3)(*[v.data_ptr() for v in vols])
        lens = (_lib.C.c_int * 3)(*[int(v.shape[2]) for v in vols])
        gptrs = (_lib.C.c_void_p * 3)(*[x.data_ptr() for x in grads])
        with _lib.timed('code_reg_bwd'):
            check(_lib.lib().mb_code_reg(codes, lens, ptr(t_dev), _lib.C.c_float(ctx.inv_frames), None, ptr(g.reshape(1).contiguous().float()), gptrs,
                                         stream()), 'code_reg(backward)')
        if ctx.sinks is not None:
            return None, None, None, None, None, None
        return None, None, None, grads[0], grads[1], grads[2]


# ------------------------------------------------------------------------------------------------
# parameter arena: weight_norm + transpose + pad of the 18 dense layers in ONE launch (and one backward launch)
# ------------------------------------------------------------------------------------------------
class _PackArena(torch.autograd.Function):
    """tensors (weight_v, weight_g, bias | weight, bias per layer) -> flat arena of packing.LAYOUT (mb_pack_arena_forward);
    backward maps the flat gradient arena back onto the parameters (mb_pack_arena_backward).  Replaces ~20 eager ops per
    layer and direction (norm, div, mul, pad, transpose, cat and their autograd nodes)."""

    @staticmethod
    def forward(ctx, owner, table, gtable, n_grad, *tensors):
        arena = torch.empty(packing.ARENA_FLOATS, device=tensors[0].device, dtype=torch.float32)
        with _lib.timed('pack_arena_fwd'):
            check(_lib.lib().mb_pack_arena_forward(ptr(table), table.shape[0], ptr(arena), stream()), 'pack_arena_forward')
        ctx.owner, ctx.table, ctx.gtable, ctx.n_grad = owner, table, gtable, n_grad
        ctx.save_for_backward(*tensors)
        return arena

    @staticmethod
    def backward(ctx, g_arena):
        tensors = ctx.saved_tensors
        ctx.owner._arena_cache = None      # this graph is consumed: the next query packs again
        sink = ctx.owner._direct_grad_table(tensors)
        if sink is not None:
            # gradient sink (train.FlatAdam): accumulate straight into the parameters' .grad views of the flat gradient buffer
            with _lib.timed('pack_arena_bwd'):
                check(_lib.lib().mb_pack_arena_backward(ptr(ctx.table), ptr(sink), ctx.table.shape[0], ptr(g_arena.contiguous()), None, stream()),
                      'pack_arena_backward')
            return (None, None, None, None) + (None,) * len(tensors)
        flat = torch.empty(ctx.n_grad, device=g_arena.device, dtype=torch.float32)
        with _lib.timed('pack_arena_bwd'):
            check(_lib.lib().mb_pack_arena_backward(ptr(ctx.table), ptr(ctx.gtable), ctx.table.shape[0], ptr(g_arena.contiguous()), ptr(flat),
                                                    stream()), 'pack_arena_backward')
        grads, off = [], 0
        for t in tensors:
            grads.append(flat[off:off + t.numel()].view(t.shape))
            off += t.numel()
        return (None, None, None, None) + tuple(grads)


# ------------------------------------------------------------------------------------------------
# the fused query as an autograd op
# ------------------------------------------------------------------------------------------------
class _FieldQuery(torch.autograd.Function):
    """One fused launch forward (mb_field_forward), one fused launch backward (mb_field_backward).
    Tensor inputs: x, t, light, topo_in, arena, emb_sdf, emb_col, code0..2, beta.
    Outputs: sdf, sigma, color, normal, normal_raw, deform, topo (None where not requested)."""

    @staticmethod
    def forward(ctx, cfg, x, t, light, topo_in, arena, emb_sdf, emb_col, code0, code1, code2, beta):
        flags, shading, ratio, n_levels, n_freq, offsets, bound, S, H, tcws, sinks = cfg
        M = x.shape[0]
        dev = x.device
        x = x.contiguous().float()
        t = t.contiguous().float().reshape(-1) if t is not None else None
        light = light.contiguous().float() if light is not None else None
        topo_in = topo_in.contiguous().float() if topo_in is not None else None
        codes = [c.detach().reshape(c.shape[1], c.shape[2]).contiguous() for c in (code0, code1, code2)]
        P = packing.fill_descs(_lib.FieldParams())
        P.arena = ptr(arena.detach())
        P.emb_sdf, P.emb_col, P.offsets = ptr(emb_sdf.detach()), ptr(emb_col.detach()), ptr(offsets)
        for i in range(3):
            P.code[i] = codes[i].data_ptr()
            P.code_len[i] = codes[i].shape[1]
        beta_d = beta.detach().reshape(1).contiguous().float()
        P.beta = ptr(beta_d)
        P.bound, P.two_bound, P.S, P.H = bound, float(2 * bound), S, H
        P.n_levels, P.n_freq = n_levels, n_freq
        io = _lib.FieldIO()
        io.M, io.flags, io.shading, io.ratio = M, flags, shading, float(ratio)
        io.x, io.t, io.light, io.topo_in = ptr(x), ptr(t), ptr(light), ptr(topo_in)

        def out(n, cond):
            return torch.empty((M, n) if n > 1 else (M,), device=dev, dtype=torch.float32) if cond else None
        sdf, sigma = out(1, flags & F_MAIN), out(1, flags & F_MAIN)
        color = out(3, (flags & F_COLOR) or shading in (2, 3))
        normal, normal_raw = out(3, flags & F_FD), out(3, flags & F_FD)
        deform, topo = out(3, flags & F_WARP), out(2, flags & F_WARP)
        io.sdf, io.sigma, io.color, io.normal, io.normal_raw = ptr(sdf), ptr(sigma), ptr(color), ptr(normal), ptr(normal_raw)
        io.deform, io.topo = ptr(deform), ptr(topo)
        stash = None
        if _lib.USE_TC:
            tabs = _tc_tables(dev)
            needs_grad = any(ctx.needs_input_grad)   # (grad mode itself is off inside Function.forward)
            if _lib.USE_TC_BWD and (flags & F_WARP) and needs_grad:
                stash = torch.empty(((M + 127) // 128) * 10 * 65536, dtype=torch.uint8, device=dev)
            tcw = tcws['f']
            with _lib.timed('field_fwd_main' if flags & F_MAIN else 'field_fwd_aux'):
                check(_lib.lib().mb_field_forward_tc(_lib.C.byref(P), _lib.C.byref(io), ptr(tcw), ptr(tabs[1]), ptr(stash), stream()), 'field_forward_tc')
        else:
            with _lib.timed('field_fwd_main' if flags & F_MAIN else 'field_fwd_aux'):
                check(_lib.lib().mb_field_forward(_lib.C.byref(P), _lib.C.byref(io), stream()), 'field_forward')
        ctx.cfg = cfg
        ctx.stash = stash
        ctx.shapes = (code0.shape, code1.shape, code2.shape)
        ctx.save_for_backward(x, t, light, topo_in, arena, emb_sdf, emb_col, codes[0], codes[1], codes[2], beta_d, deform, topo, normal_raw)
        ctx.set_materialize_grads(False)
        return sdf, sigma, color, normal, normal_raw, deform, topo

    @staticmethod
    def backward(ctx, g_sdf, g_sigma, g_color, g_normal, g_raw, g_deform, g_topo):
        flags, shading, ratio, n_levels, n_freq, offsets, bound, S, H, tcws, sinks = ctx.cfg
        x, t, light, topo_in, arena, emb_sdf, emb_col, c0, c1, c2, beta_d, deform, topo, normal_raw = ctx.saved_tensors
        M = x.shape[0]
        codes = [c0, c1, c2]
        P = packing.fill_descs(_lib.FieldParams())
        P.arena = ptr(arena.detach())
        P.emb_sdf, P.emb_col, P.offsets = ptr(emb_sdf.detach()), ptr(emb_col.detach()), ptr(offsets)
        for i in range(3):
            P.code[i] = codes[i].data_ptr()
            P.code_len[i] = codes[i].shape[1]
        P.beta = ptr(beta_d)
        P.bound, P.two_bound, P.S, P.H = bound, float(2 * bound), S, H
        P.n_levels, P.n_freq = n_levels, n_freq
        io = _lib.FieldIO()
        io.M, io.flags, io.shading, io.ratio = M, flags, shading, float(ratio)
        io.x, io.t, io.light, io.topo_in = ptr(x), ptr(t), ptr(light), ptr(topo_in)

        def cg(g):
            return g.contiguous().float() if g is not None else None
        g_sdf, g_sigma, g_color, g_normal, g_raw, g_deform, g_topo = map(cg, (g_sdf, g_sigma, g_color, g_normal, g_raw, g_deform, g_topo))
        G = _lib.FieldGrads()
        G.g_sdf, G.g_sigma, G.g_color, G.g_normal, G.g_normal_raw, G.g_deform, G.g_topo = map(ptr, (g_sdf, g_sigma, g_color, g_normal, g_raw, g_deform, g_topo))
        G.deform, G.topo, G.normal_raw = ptr(deform), ptr(topo), ptr(normal_raw)
        g_arena = torch.zeros_like(arena)
        if sinks is not None:       # accumulate (red.global.add) straight into the parameters' .grad: no zero-fill, no autograd adds
            g_es, g_ec, g_codes = sinks[0], sinks[1], [g.view(c.shape) for g, c in zip(sinks[2:], codes)]
        else:
            g_es = torch.zeros_like(emb_sdf)
            g_ec = torch.zeros_like(emb_col)
            g_codes = [torch.zeros_like(c) for c in codes]
        g_beta = torch.zeros(1, device=x.device, dtype=torch.float32)
        g_x = torch.empty_like(x)
        g_topo_in = torch.empty_like(topo_in) if topo_in is not None else None
        G.g_arena, G.g_emb_sdf, G.g_emb_col, G.g_beta, G.g_x, G.g_topo_in = map(ptr, (g_arena, g_es, g_ec, g_beta, g_x, g_topo_in))
        for i in range(3):
            G.g_code[i] = g_codes[i].data_ptr()
        stash = getattr(ctx, 'stash', None)
        g_def_out = g_topo_out = None
        if stash is not None:
            io.flags = flags | _lib.F_SKIP_WARP_BWD
            g_def_out = torch.empty(M, 3, device=x.device)
            g_topo_out = torch.empty(M, 2, device=x.device)
            G.g_def_out, G.g_topo_out = ptr(g_def_out), ptr(g_topo_out)
        use_sdf_tc = _lib.USE_TC and _lib.USE_TC_BWD_SDF and (flags & (F_MAIN | F_FD)) and (stash is not None or not (flags & F_WARP))
        if use_sdf_tc:
            tabs_f = _tc_tables(x.device)
            tabs_s = _tc_tables_small(x.device)
            tcw_f, tcw_s = tcws['f'], tcws['s']
            # FD-normal queries: specialised two-CTAs-per-SM kernel; with MAIN the general kernel first turns the upstream
            # gradients into d/d(sdf) of the six queries per sample (g_fd) and leaves the FD chains to it
            fd_kernel = _lib.USE_TC_BWD_FD and bool(flags & F_FD) and (bool(flags & F_MAIN) or shading == 0 or g_color is None) \
                and (g_normal is not None or g_raw is not None or (shading != 0 and g_color is not None))
            if flags & F_MAIN or not fd_kernel:
                if fd_kernel:
                    g_fd = torch.empty(M, 6, device=x.device, dtype=torch.float32)
                    G.g_fd = ptr(g_fd)
                    io.flags = io.flags | _lib.F_FD_DELEGATE
                with _lib.timed('field_bwd_sdf_tc_main' if flags & F_MAIN else 'field_bwd_sdf_tc_aux'):
                    check(_lib.lib().mb_field_backward_sdf_tc(_lib.C.byref(P), _lib.C.byref(io), _lib.C.byref(G), ptr(tcw_f), ptr(tabs_f[1]),
                                                              ptr(tcw_s), ptr(tabs_s[1]), stream()), 'field_backward_sdf_tc')
            if fd_kernel:
                with _lib.timed('field_bwd_fd_tc_main' if flags & F_MAIN else 'field_bwd_fd_tc_aux'):
                    check(_lib.lib().mb_field_backward_fd_tc(_lib.C.byref(P), _lib.C.byref(io), _lib.C.byref(G), ptr(tcw_f), ptr(tabs_f[1]),
                                                             ptr(tcw_s), ptr(tabs_s[1]), 1 if flags & F_MAIN else 0, stream()), 'field_backward_fd_tc')
        else:
            with _lib.timed('field_bwd_main' if flags & F_MAIN else 'field_bwd_aux'):
                check(_lib.lib().mb_field_backward(_lib.C.byref(P), _lib.C.byref(io), _lib.C.byref(G), stream()), 'field_backward')
        if stash is not None:
            tabs = _tc_tables_dgrad(x.device)
            tcw_t = tcws['t']
            gcode_ptrs = (_lib.C.c_void_p * 3)(*[g.data_ptr() for g in g_codes])
            with _lib.timed('field_bwd_warp_tc'):
                check(_lib.lib().mb_field_backward_warp_tc(_lib.C.byref(P), ptr(x), ptr(t), M, ptr(g_def_out), ptr(g_topo_out), ptr(stash),
                                                           ptr(tcw_t), ptr(tabs[1]), ptr(g_arena), gcode_ptrs, ptr(g_x), stream()),
                      'field_backward_warp_tc')
            ctx.stash = None
        if sinks is not None:
            return (None, g_x, None, None, g_topo_in, g_arena, None, None, None, None, None, g_beta.reshape(()))
        gc = [g.reshape(s) for g, s in zip(g_codes, ctx.shapes)]
        return (None, g_x, None, None, g_topo_in, g_arena, g_es, g_ec, gc[0], gc[1], gc[2], g_beta.reshape(()))


class _FDRegulariser(torch.autograd.Function):
    """loss_normal_perturb of MorpheuS.render_rays (morpheus.py:714-741) on a real view, forward AND backward in ONE launch
    (mb_fd_regulariser_tc, csrc/field_fd_reg_tc.cu): both 6-point FD normals of every sample (at x with the sample's topo, at
    x + noise * std with topo = 0), the L1 difference, and its gradients w.r.t. the SDF hash table, the SDF decoder, x and topo.
    The launch computes the gradients of the returned scalar; backward() only scales them by the upstream gradient."""

    @staticmethod
    def forward(ctx, cfg, x, topo, noise, arena, emb_sdf):
        n_levels, n_freq, offsets, bound, S, H, tcws, sink_emb, noise_std, gmul = cfg
        M = x.shape[0]
        dev = x.device
        x = x.contiguous().float()
        topo_c = topo.contiguous().float() if topo is not None else None
        noise_c = noise.contiguous().float() if noise is not None else None
        P = packing.fill_descs(_lib.FieldParams())
        P.arena = ptr(arena.detach())
        P.emb_sdf, P.offsets = ptr(emb_sdf.detach()), ptr(offsets)
        P.bound, P.two_bound, P.S, P.H = bound, float(2 * bound), S, H
        P.n_levels, P.n_freq = n_levels, n_freq
        normal = torch.empty(M, 3, device=dev, dtype=torch.float32)
        normal_raw = torch.empty(M, 3, device=dev, dtype=torch.float32)
        # one zero-fill launch for the three accumulation targets (loss | table gradient | arena gradient)
        n_e, n_a = emb_sdf.numel(), arena.numel()
        acc = torch.zeros(4 + n_e + n_a, device=dev, dtype=torch.float32)
        loss, g_emb, g_arena = acc[:1], acc[4:4 + n_e].view(emb_sdf.shape), acc[4 + n_e:].view(arena.shape)
        g_x = torch.empty(M, 3, device=dev, dtype=torch.float32)
        g_topo = torch.empty(M, 2, device=dev, dtype=torch.float32) if topo is not None else None
        tabs_f, tabs_s = _tc_tables(dev), _tc_tables_small(dev)
        with _lib.timed('fd_regulariser'):
            check(_lib.lib().mb_fd_regulariser_tc(_lib.C.byref(P), ptr(x), ptr(topo_c), ptr(noise_c), _lib.C.c_float(noise_std), M, _lib.C.c_float(gmul),
                                                  ptr(normal), ptr(normal_raw), ptr(loss), ptr(g_x), ptr(g_topo), ptr(g_emb), ptr(g_arena),
                                                  ptr(tcws['f']), ptr(tabs_f[1]), ptr(tcws['s']), ptr(tabs_s[1]), stream()), 'fd_regulariser_tc')
        ctx.sink_emb = sink_emb
        ctx.has_topo = topo is not None
        ctx.save_for_backward(g_x, g_topo, g_emb, g_arena)
        ctx.mark_non_differentiable(normal, normal_raw)
        return loss[0].clone(), normal, normal_raw

    @staticmethod
    def backward(ctx, g_loss, _gn, _gr):
        g_x, g_topo, g_emb, g_arena = ctx.saved_tensors
        g = g_loss.reshape(())
        if ctx.sink_emb is not None:       # train.FlatAdam: straight into the table's .grad view of the flat gradient buffer
            ctx.sink_emb.addcmul_(g_emb, g)
            g_emb_out = None
        else:
            g_emb_out = g_emb * g
        outs = torch._foreach_mul([g_x, g_topo, g_arena] if ctx.has_topo else [g_x, g_arena], g)      # one launch for the per-sample / arena scalings
        return None, outs[0], (outs[1] if ctx.has_topo else None), None, outs[-1], g_emb_out


# ------------------------------------------------------------------------------------------------
class scene_representation(nn.Module):
    def __init__(self, config, bound, max_level=None, num_layers=3, num_layers_t=6, hidden_dim=64, hidden_dim_t=128,
                 hidden_dim_tpo=128, num_layers_bg=2, geo_dim=32, deform_dim=16, hidden_dim_bg=32, amb_dim=2, num_frames=None,
                 use_app=False, use_t=False, color_grid=True, use_joint=False, encode_topo=False, encode_deform=True):
        super().__init__()
        shipped = dict(num_layers=3, num_layers_t=6, hidden_dim=64, hidden_dim_t=128, hidden_dim_tpo=128, geo_dim=32,
                       deform_dim=16, amb_dim=2, use_app=False, use_t=False, color_grid=True, use_joint=True,
                       encode_topo=False, encode_deform=True)
        given = dict(num_layers=num_layers, num_layers_t=num_layers_t, hidden_dim=hidden_dim, hidden_dim_t=hidden_dim_t,
                     hidden_dim_tpo=hidden_dim_tpo, geo_dim=geo_dim, deform_dim=deform_dim, amb_dim=amb_dim, use_app=use_app,
                     use_t=use_t, color_grid=color_grid, use_joint=use_joint, encode_topo=encode_topo, encode_deform=encode_deform)
        if given != shipped:
            diff = {k: v for k, v in given.items() if shipped[k] != v}
            raise NotImplementedError(f'morpheus_b200 kernels are specialised for the configuration every MorpheuS yaml ships '
                                      f'(morpheus.py:131-140, configs/*.yaml model section); unsupported overrides: {diff}')
        self.config, self.bound, self.max_level = config, bound, max_level
        self.num_layers, self.hidden_dim, self.geo_dim, self.num_frames = num_layers, hidden_dim, geo_dim, num_frames
        self.use_t, self.use_app, self.use_joint, self.encode_topo, self.encode_deform = use_t, use_app, use_joint, encode_topo, encode_deform
        self.pose_array = PoseArray(num_frames)
        self.in_dim_amb = amb_dim
        self.encoder_deform, self.in_dim_deform = get_encoder('frequency_torch', input_dim=3, multires=6)
        self.deform_code, self.deform_dim = MultiCode([num_frames // 8, num_frames // 4, num_frames], deform_dim), 3 * deform_dim
        self.deform_net = MLP(self.in_dim_deform + self.deform_dim, 3, hidden_dim_t, num_layers_t, bias=True)
        self.topo_net = MLP(self.in_dim_deform + self.deform_dim, amb_dim, hidden_dim_tpo, num_layers_t, bias=True)
        self.encoder, self.in_dim = get_encoder('hashgrid', input_dim=3, num_levels=16, log2_hashmap_size=15, desired_resolution=128)
        self.encoder_c, self.in_dim_c = get_encoder('hashgrid', input_dim=3, num_levels=16, log2_hashmap_size=15, desired_resolution=128)
        self.encoder_xyz, self.in_dim_xyz = get_encoder('frequency_torch', input_dim=3, multires=6)
        self.sdf_net = MLP(self.in_dim + self.in_dim_amb + self.in_dim_xyz, 1 + geo_dim, hidden_dim, num_layers, bias=True,
                           geo_init=True, geo_bias=0.4, weight_norm=False)
        self.color_net = MLP(self.in_dim_c + geo_dim, 3, hidden_dim, num_layers, bias=True)
        if self.config['model']['bg_radius'] > 0:
            self.encoder_bg, self.in_dim_bg = get_encoder('frequency_torch', input_dim=3, multires=6)
            self.encoder_bg_t, self.in_dim_bg_t = get_encoder('frequency_torch', input_dim=1, multires=6)
            self.bg_net = MLP(self.in_dim_bg + self.in_dim_bg_t, 3, hidden_dim_bg, num_layers_bg, bias=True)
        self.sdf2density = LaplaceDensity({'beta': 0.1})
        self._arena_cache = None
        self._pack_tab = None
        self._sink_tab = None
        self.grad_sink = False      # set by train.FlatAdam: kernels accumulate parameter gradients straight into .grad

    # -- packed parameters ----------------------------------------------------------------------------
    def _pack_inputs(self):
        tensors = []
        for mlp in (self.deform_net, self.topo_net, self.sdf_net, self.color_net):
            for lin in mlp.net:
                tensors += [lin.weight_v, lin.weight_g, lin.bias] if mlp.uses_weight_norm else [lin.weight, lin.bias]
        return tensors

    def _pack_tables(self, tensors):
        """device tables of mb_pack_arena_* (include/morpheus_b200.h), rebuilt only when a parameter's storage moves"""
        key = tuple(t.data_ptr() for t in tensors)
        if self._pack_tab is None or self._pack_tab[0] != key:
            rows, grows, it, off = [], [], iter(tensors), 0
            for name, mlp in (('deform', self.deform_net), ('topo', self.topo_net), ('sdf', self.sdf_net), ('color', self.color_net)):
                for (wt_off, w_off, b_off, K, N, Kp, Np) in packing.LAYOUT[name]:
                    if mlp.uses_weight_norm:
                        v, g, b = next(it), next(it), next(it)
                        rows.append([v.data_ptr(), g.data_ptr(), b.data_ptr(), K, N, Kp, Np, wt_off])
                        grows.append([off, off + v.numel(), off + v.numel() + g.numel(), 0])
                        off += v.numel() + g.numel() + b.numel()
                    else:
                        w, b = next(it), next(it)
                        rows.append([w.data_ptr(), 0, b.data_ptr(), K, N, Kp, Np, wt_off])
                        grows.append([off, -1, off + w.numel(), 0])
                        off += w.numel() + b.numel()
                    for t in (rows[-1][0], rows[-1][2]):
                        assert t % 4 == 0
            dev = tensors[0].device
            self._pack_tab = (key, torch.tensor(rows, dtype=torch.int64, device=dev), torch.tensor(grows, dtype=torch.int64, device=dev), off)
        return self._pack_tab[1:]

    def _direct_grad_table(self, tensors):
        """with `grad_sink` set (train.FlatAdam): device table of the parameters' .grad addresses for mb_pack_arena_backward's
        direct mode ({gv, gg, gb, 0} per layer), or None when any .grad is missing / the sink is off"""
        if not self.grad_sink or any(t.grad is None or not t.grad.is_contiguous() for t in tensors):
            return None
        key = tuple(t.grad.data_ptr() for t in tensors)
        if self._sink_tab is None or self._sink_tab[0] != key:
            rows, it = [], iter(tensors)
            for mlp in (self.deform_net, self.topo_net, self.sdf_net, self.color_net):
                for _ in mlp.net:
                    if mlp.uses_weight_norm:
                        v, g, b = next(it), next(it), next(it)
                        rows.append([v.grad.data_ptr(), g.grad.data_ptr(), b.grad.data_ptr(), 0])
                    else:
                        w, b = next(it), next(it)
                        rows.append([w.grad.data_ptr(), 0, b.grad.data_ptr(), 0])
            self._sink_tab = (key, torch.tensor(rows, dtype=torch.int64, device=tensors[0].device))
        return self._sink_tab[1]

    def _sinks(self):
        """.grad tensors the field kernels may accumulate into directly (hash tables, code lines) when `grad_sink` is set"""
        if not self.grad_sink or not torch.is_grad_enabled():
            return None
        ps = [self.encoder.embeddings, self.encoder_c.embeddings] + list(self.deform_code.volumes)
        if any(p.grad is None or not p.grad.is_contiguous() or not p.requires_grad for p in ps):
            return None
        return [p.grad for p in ps]

    def packed_arena(self):
        """(flat effective-weight arena, tensor-core operand tables).  Packed by ONE launch (differentiable through
        _PackArena) and cached until a parameter changes (version counters), the arena's backward has run, or `invalidate()`
        is called (optimisers that update parameters through raw pointers, e.g. train.FlatAdam, must call it)."""
        tensors = self._pack_inputs()
        for t in tensors:
            if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
                raise RuntimeError('morpheus_b200: MLP parameters must be contiguous fp32 CUDA tensors (there is no CPU path)')
        key = (torch.is_grad_enabled(), tuple((t.data_ptr(), t._version) for t in tensors))
        c = self._arena_cache
        if c is not None and c[0] == key:
            return c[1], c[2]
        table, gtable, n_grad = self._pack_tables(tensors)
        arena = _PackArena.apply(self, table, gtable, n_grad, *tensors)
        tcw = self._pack_tc(arena.detach()) if _lib.USE_TC else None
        self._arena_cache = (key, arena, tcw)
        return arena, tcw

    @staticmethod
    def _pack_tc(arena):
        """fp16 (hi, lo) tensor-core operand slabs of the arena: forward table, dgrad table of the deform/topology nets,
        dgrad table of the SDF/colour nets -- packed by ONE mb_pack_tc launch into one buffer (the three tables are views);
        shared by every query and backward until the arena changes"""
        dev = arena.device
        key = str(dev)
        if key not in _TC_TABLES_ALL:
            parts = [('f', _tc_tables(dev)), ('t', _tc_tables_dgrad(dev)), ('s', _tc_tables_small(dev))]
            descs, spans, base = [], {}, 0
            for name, tabs in parts:
                d = tabs[0].clone()
                d[:, 5] += base                      # dst_off of every layer, relative to the combined buffer
                descs.append(d)
                spans[name] = (base, base + int(tabs[2]))
                base += (int(tabs[2]) + 1023) // 1024 * 1024
            _TC_TABLES_ALL[key] = (torch.cat(descs, 0).contiguous(), spans, base)
        desc, spans, total = _TC_TABLES_ALL[key]
        w = torch.empty(total, dtype=torch.uint8, device=dev)
        with _lib.timed('pack_tc'):
            check(_lib.lib().mb_pack_tc(ptr(arena), ptr(desc), int(desc.shape[0]), ptr(w), stream()), 'pack_tc')
        return {name: w[a:b] for name, (a, b) in spans.items()}

    def invalidate(self):
        self._arena_cache = None

    def train(self, mode=True):
        self._arena_cache = None
        return super().train(mode)

    def _levels(self):
        L = 16
        n_levels = L if self.max_level is None else max(min(int(math.ceil(self.max_level * L)), L), 1)   # grid.py:42
        n_freq = 6 if self.max_level is None else int(self.max_level * 6)                                 # encodings.py:37-40
        return n_levels, n_freq

    def _query(self, x, t, flags, shading=0, ratio=1.0, light=None, topo_in=None, arena=None):
        if x.shape[0] == 0:
            raise RuntimeError('morpheus_b200: empty query (the reference would crash at morpheus.py:701 as well)')
        n_levels, n_freq = self._levels()
        enc = self.encoder
        cfg = (int(flags), int(shading), float(ratio), n_levels, n_freq, enc.offsets, float(self.bound),
               float(np.log2(enc.per_level_scale)), int(enc.base_resolution))
        tcw = None
        if arena is None:
            arena, tcw = self.packed_arena()
            if _lib.USE_TC and tcw is None:      # engine switched on after the arena was packed (tests toggle it)
                tcw = self._pack_tc(arena.detach())
                if self._arena_cache is not None and self._arena_cache[1] is arena:
                    self._arena_cache = (self._arena_cache[0], arena, tcw)
        elif _lib.USE_TC:
            tcw = self._pack_tc(arena.detach())
        cfg = cfg + (tcw, self._sinks())
        v = self.deform_code.volumes
        return _FieldQuery.apply(cfg, x, t, light, topo_in, arena, self.encoder.embeddings, self.encoder_c.embeddings,
                                 v[0], v[1], v[2], self.sdf2density.get_beta())

    # -- reference API --------------------------------------------------------------------------------
    def get_deform_code(self, t, app=False):
        return self.deform_code.sample(t)

    def code_regulariser(self, t, num_frames):
        """morpheus.py:766-771 on the three code lines: mean (2 c(t) - c(t - 1/F) - c(t + 1/F))^2, fused (one launch each way)"""
        t_dev = t.reshape(-1)[:1].contiguous().float()
        v = self.deform_code.volumes
        if not v[0].is_cuda:
            codes = self.get_deform_code(torch.cat([t_dev.view(1, 1), t_dev.view(1, 1) - 1 / num_frames, t_dev.view(1, 1) + 1 / num_frames], dim=0))
            return torch.square(2 * codes[0:1] - codes[1:2] - codes[2:3]).mean()
        sinks = self._sinks()
        return _CodeReg.apply(t_dev, 1.0 / num_frames, sinks[2:] if sinks is not None else None, v[0], v[1], v[2])

    def get_RT(self, frame_ids):
        frame_ids = frame_ids.squeeze()
        return self.pose_array.get_rotation_matrices(frame_ids), self.pose_array.get_translations(frame_ids)

    def warp(self, x, t):
        """model.py:412-437 -> (deform, topo, app_code=None)"""
        out = self._query(x, t, F_WARP)
        return out[5], out[6], None

    def get_topo(self, x, t):
        return self.warp(x, t)[1]

    def get_sigma_albedo(self, x, topo=None, app_code=None, return_color=True):
        """model.py:273-307"""
        flags = F_MAIN | (F_COLOR if return_color else 0) | (F_TOPO_IN if topo is not None else 0)
        out = self._query(x, None, flags, topo_in=topo)
        return out[0], out[1], out[2]

    def finite_difference_normal(self, x, epsilon=2e-3, topo=None):
        assert abs(epsilon - 2e-3) < 1e-12, 'the kernel is specialised for epsilon = 2e-3 (model.py:367)'
        out = self._query(x, None, F_FD | (F_TOPO_IN if topo is not None else 0), topo_in=topo)
        return out[4]

    def normal(self, x, t=None, cano=False, topo=None):
        """model.py:387-398 -> (normal, normal_raw)"""
        if t is not None and not cano:
            out = self._query(x, t, F_WARP | F_FD | F_FD_WARPED)
        else:
            out = self._query(x, None, F_FD | (F_TOPO_IN if topo is not None else 0), topo_in=topo)
        return out[3], out[4]

    def background(self, d, t):
        """model.py:400-410 (never reached with the shipped flags); plain torch."""
        h = torch.cat([self.encoder_bg(d), self.encoder_bg_t(t, max_level=self.max_level)], dim=-1)
        return torch.sigmoid(self.bg_net(h))

    def pose_optimisation(self, rays_o, rays_d, frame_ids):
        """model.py:335-346"""
        frame_ids = frame_ids.squeeze()
        if rays_o.is_cuda and rays_o.dim() == 2 and frame_ids.dim() == 1 and frame_ids.shape[0] == rays_o.shape[0]:
            return _PoseRays.apply(self.pose_array.data, rays_o, rays_d, frame_ids)
        R = self.pose_array.get_rotation_matrices(frame_ids)
        tr = self.pose_array.get_translations(frame_ids)
        return rays_o + tr, torch.sum(rays_d[..., None, :] * R, -1)

    def fd_regulariser(self, x, topo, noise, noise_std, gmul):
        """Real-view normal regulariser in one fused launch: -> (gmul * sum |n(x, topo) - n(x + noise * std, 0)|, n, n_raw) with
        n = normal(x, topo) of models/model.py:387-398 and the perturbed query of morpheus.py:724-736 (topo_none).  Tensor-core engine only."""
        if not _lib.USE_TC:
            raise RuntimeError('morpheus_b200: fd_regulariser needs the tensor-core engine (MORPHEUS_B200_TC=1)')
        if x.shape[0] == 0:
            raise RuntimeError('morpheus_b200: empty query')
        n_levels, n_freq = self._levels()
        enc = self.encoder
        arena, tcw = self.packed_arena()
        if tcw is None:
            tcw = self._pack_tc(arena.detach())
        sinks = self._sinks()
        cfg = (n_levels, n_freq, enc.offsets, float(self.bound), float(np.log2(enc.per_level_scale)), int(enc.base_resolution), tcw,
               sinks[0] if sinks is not None else None, float(noise_std), float(gmul))
        return _FDRegulariser.apply(cfg, x, topo, noise, arena, enc.embeddings)

    def forward_with_topo(self, x, t):
        """forward(x, t, shading='albedo') plus the topology coordinates of the warp (differentiable): -> (sdf, sigma, albedo, deform, topo)"""
        out = self._query(x, t, F_MAIN | F_COLOR | F_WARP)
        return out[0], out[1], out[2], out[5], out[6]

    def density(self, x, t=None, cano=False, allow_shape=False, return_color=True):
        """model.py:439-481"""
        if cano or t is None:
            out = self._query(x, None, F_MAIN | (F_COLOR if return_color else 0))
        else:
            if isinstance(t, float):
                t = t * torch.ones(x.shape[0], 1, device=x.device)
            if x.shape[0] != t.shape[0]:
                if not allow_shape:
                    raise Exception('Shape inconsistent!!!')
                t = t[0, 0] * torch.ones(x.shape[0], 1, device=x.device)
            out = self._query(x, t, F_WARP | F_MAIN | (F_COLOR if return_color else 0))
        return {'sdf': out[0], 'sigma': out[1], 'albedo': out[2]}

    def forward(self, x, t, light_dir=None, ratio=1, shading='albedo', cano=False, return_color=True):
        """model.py:483-533 -> (sdf, sigma, color, normal, deform, normal_raw)"""
        flags = F_MAIN | (F_COLOR if return_color else 0) | (0 if cano else F_WARP)
        if shading == 'albedo':
            out = self._query(x, None if cano else t, flags)
            return out[0], out[1], out[2], None, out[5], None
        out = self._query(x, None if cano else t, flags | F_FD, shading=SHADE[shading], ratio=ratio, light=light_dir)
        return out[0], out[1], out[2], out[3], out[5], out[4]

    def get_params_all(self, lr):
        """model.py:309-333 (same group names; morpheus.py:494-516 looks them up by name)"""
        params = [
            {'name': 'encoder_sdf', 'params': self.encoder.parameters(), 'lr': lr},
            {'name': 'encoder_color', 'params': self.encoder_c.parameters(), 'lr': lr},
            {'name': 'decoder_sdf', 'params': self.sdf_net.parameters(), 'lr': lr},
            {'name': 'decoder_topo', 'params': self.topo_net.parameters(), 'lr': lr},
            {'name': 'decoder_color', 'params': self.color_net.parameters(), 'lr': lr},
            {'name': 'density', 'params': self.sdf2density.parameters(), 'lr': lr / 2.},
            {'name': 'decoder_deform', 'params': self.deform_net.parameters(), 'lr': lr},
            {'name': 'code_deform', 'params': self.deform_code.parameters(), 'lr': lr},
            {'name': 'pose', 'params': self.pose_array.parameters(), 'lr': lr / 10.},
        ]
        if self.config['model']['bg_radius'] > 0:
            params.append({'name': 'decoder_bg', 'params': self.bg_net.parameters(), 'lr': lr})
        return params

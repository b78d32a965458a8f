"""Parameter arena layout shared by the host code and the field kernels (mb_layer_desc in
include/morpheus_b200.h).

Every dense layer of the four MLPs on the hot path (models/model.py:138-139,169-174) is stored as
  Wt [K_pad, N_pad]  k-major   (forward operand;  weight gradients are accumulated in this slot)
  W  [N_pad, K_pad]  n-major   (data-gradient operand)
  b  [N_pad]
with K_pad, N_pad rounded up to multiples of 16 and zero filled.  `pack` is built from
differentiable torch ops, so autograd maps the kernel's flat gradient arena back onto the
reference parameters (including weight_norm's weight_g / weight_v, models/decoders.py:51-52).
"""
import torch
import torch.nn.functional as F

from . import _lib

NET_DIMS = {
    'deform': [(87, 128), (128, 128), (128, 128), (128, 128), (128, 128), (128, 3)],
    'topo': [(87, 128), (128, 128), (128, 128), (128, 128), (128, 128), (128, 2)],
    'sdf': [(73, 64), (64, 64), (64, 33)],
    'color': [(64, 64), (64, 64), (64, 3)],
}
NET_ORDER = ('deform', 'topo', 'sdf', 'color')


def _pad16(n):
    return (n + 15) // 16 * 16


def layout():
    """-> (dict net -> list of (wt_off, w_off, b_off, K, N, K_pad, N_pad), total floats)"""
    off = 0
    table = {}
    for net in NET_ORDER:
        rows = []
        for (K, N) in NET_DIMS[net]:
            Kp, Np = _pad16(K), _pad16(N)
            wt_off = off
            w_off = wt_off + Kp * Np
            b_off = w_off + Kp * Np
            off = b_off + Np
            rows.append((wt_off, w_off, b_off, K, N, Kp, Np))
        table[net] = rows
    return table, off


LAYOUT, ARENA_FLOATS = layout()


def pack(layers):
    """layers: dict net -> list of (W [N,K], b [N]) effective (weight-normed) tensors -> flat arena."""
    parts = []
    for net in NET_ORDER:
        for (W, b), (_, _, _, K, N, Kp, Np) in zip(layers[net], LAYOUT[net]):
            assert tuple(W.shape) == (N, K), (net, tuple(W.shape), (N, K))
            Wp = F.pad(W, (0, Kp - K, 0, Np - N))            # [Np, Kp]
            parts += [Wp.t().reshape(-1), Wp.reshape(-1), F.pad(b, (0, Np - N))]
    return torch.cat(parts)


def fill_descs(params):
    """write the static layer table into a _lib.FieldParams"""
    for net in NET_ORDER:
        arr = getattr(params, net)
        for i, row in enumerate(LAYOUT[net]):
            d = arr[i]
            d.wt_off, d.w_off, d.b_off, d.K, d.N, d.K_pad, d.N_pad = row
    return params


def tc_tables(device):
    """Static tables for the tensor-core engine: (pack descriptors [18,8] i32, slab offsets [18,3] i32, total bytes).
    Layer order deform[6], topo[6], sdf[3], color[3]; first layers use the core-aligned K order of csrc/field_fwd_tc.cu."""
    desc, off, dst = [], [], 0
    for net in NET_ORDER:
        for li, (wt_off, w_off, b_off, K, N, Kp, Np) in enumerate(LAYOUT[net]):
            kind = 1 if (net in ('deform', 'topo') and li == 0) else (2 if (net == 'sdf' and li == 0) else 0)
            k_tc = 96 if kind == 1 else (80 if kind == 2 else Kp)
            desc.append([w_off, K, Kp, Np, kind, dst, k_tc, 0])
            off.append([dst, k_tc // 16, Np])
            dst += (k_tc // 16) * 64 * Np
    return (torch.tensor(desc, dtype=torch.int32, device=device), torch.tensor(off, dtype=torch.int32, device=device), dst)


def tc_tables_dgrad(device):
    """Tables for the tensor-core backward of the deform / topology nets: dgrad B operands B[k_tc][n] = Wt[korig(k_tc)][n]
    (pack mode 1).  -> (pack descriptors [12,8], slab table [12,3] = {offset, N_pad/16, rows}, total bytes)"""
    desc, off, dst = [], [], 0
    for net in ('deform', 'topo'):
        for li, (wt_off, w_off, b_off, K, N, Kp, Np) in enumerate(LAYOUT[net]):
            kind = 1 if li == 0 else 0
            rows = 96 if kind == 1 else Kp
            desc.append([wt_off, K, Np, rows, kind, dst, Np, 1])
            off.append([dst, Np // 16, rows])
            dst += (Np // 16) * 64 * rows
    return (torch.tensor(desc, dtype=torch.int32, device=device), torch.tensor(off, dtype=torch.int32, device=device), dst)


def tc_tables_dgrad_small(device):
    """dgrad B operands of the SDF and colour nets (pack mode 1): 6 layers sdf[3], color[3];
    sdf layer 0 uses the core-aligned tc feature order (kind 2, 80 rows)."""
    desc, off, dst = [], [], 0
    for net in ('sdf', 'color'):
        for li, (wt_off, w_off, b_off, K, N, Kp, Np) in enumerate(LAYOUT[net]):
            kind = 2 if (net == 'sdf' and li == 0) else 0
            rows = 80 if kind == 2 else Kp
            desc.append([wt_off, K, Np, rows, kind, dst, Np, 1])
            off.append([dst, Np // 16, rows])
            dst += (Np // 16) * 64 * rows
    return (torch.tensor(desc, dtype=torch.int32, device=device), torch.tensor(off, dtype=torch.int32, device=device), dst)

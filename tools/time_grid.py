"""Stand-alone hash-grid encode: ours (mb_grid_encode_forward / backward through morpheus_b200.gridencoder._backend) against the UNMODIFIED
reference kernel (oracle/_ref, external/encoders/gridencoder/src/gridencoder.cu:83-249, :253-349) on the same box, same inputs.
B = 524 288 points (BASELINE cfg-2's M), D = 3, C = 2, L = 16, fp32; forward with and without dy_dx, backward with and without
grad_inputs.  CUDA events on the launching stream, 5 warm-up + 20 timed launches, a 256 MiB L2 flush between launches.
Algorithmic HBM bytes: forward 12 B in + 128 B out (+ 384 B dy_dx) per point = 140 / 524 B; backward 128 B grad + 12 B in
(+ 384 B dy_dx read + 12 B grad_inputs) per point = 140 / 536 B (the 3.4 MB table stays in L2).
    python tools/time_grid.py [B]      (run under gpurun; prints one JSON document)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from morpheus_b200 import gridencoder as mg  # noqa: E402
from oracle.ref_gpu_step import load_ref_backend  # noqa: E402  (TEST INFRASTRUCTURE: the reference kernel as the yardstick)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 * 128
dev = torch.device('cuda:0')
peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
hbm = peaks.get('hbm_gbs', 6650.0)
enc = mg.GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=15, desired_resolution=128).to(dev)
g = torch.Generator().manual_seed(0)
x = torch.rand(B, 3, generator=g).to(dev)
emb = enc.embeddings.detach()
with torch.no_grad():
    emb.copy_(((torch.rand(emb.shape, generator=g) * 2 - 1) * 0.1).to(dev))
offsets = enc.offsets
L, C, D = 16, 2, 3
S, H = float(np.log2(enc.per_level_scale)), 16
grad = torch.randn(L, B, C, generator=g).to(dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
ref = load_ref_backend()


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    ts = []
    for i in range(n):
        flush.fill_(i & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


out = {'B': B, 'hbm_peak_gbs': hbm, 'cases': []}
for name, backend in (('ours', mg._backend), ('reference', ref)):
    if backend is None:
        out['cases'].append({'impl': name, 'unavailable': 'oracle/_ref not built'})
        continue
    for dydx_on in (False, True):
        outputs = torch.empty(L, B, C, device=dev)
        dy_dx = torch.empty(B, L * D * C, device=dev) if dydx_on else None
        ms_f = timeit(lambda: backend.grid_encode_forward(x, emb, offsets, outputs, B, D, C, L, L, S, H, dy_dx, 0, False, 0))
        bytes_f = B * (12 + 128 + (384 if dydx_on else 0))
        g_emb = torch.zeros_like(emb)
        g_in = torch.zeros_like(x) if dydx_on else None

        def bwd():
            g_emb.zero_()
            backend.grid_encode_backward(grad, x, emb, offsets, g_emb, B, D, C, L, L, S, H, dy_dx, g_in, 0, False, 0)
        ms_b = timeit(bwd)
        bytes_b = B * (128 + 12 + ((384 + 12) if dydx_on else 0))
        out['cases'].append({'impl': name, 'dy_dx': dydx_on, 'fwd_ms': ms_f, 'fwd_gbs': bytes_f / ms_f / 1e6, 'fwd_frac_of_hbm_peak': bytes_f / ms_f / 1e6 / hbm,
                             'bwd_ms': ms_b, 'bwd_gbs': bytes_b / ms_b / 1e6, 'bwd_frac_of_hbm_peak': bytes_b / ms_b / 1e6 / hbm,
                             'algorithmic_bytes_per_point': {'fwd': bytes_f // B, 'bwd': bytes_b // B}})
ours = {c['dy_dx']: c for c in out['cases'] if c['impl'] == 'ours' and 'fwd_ms' in c}
refc = {c['dy_dx']: c for c in out['cases'] if c['impl'] == 'reference' and 'fwd_ms' in c}
out['speedup_vs_reference'] = {('dy_dx' if k else 'no_dy_dx'): {'fwd': refc[k]['fwd_ms'] / ours[k]['fwd_ms'], 'bwd': refc[k]['bwd_ms'] / ours[k]['bwd_ms']}
                               for k in ours if k in refc}
print(json.dumps(out))

"""Reference-style GPU step (TEST INFRASTRUCTURE, never imported by the product path).

What the reference itself does on a GPU for the BASELINE cfg-2 real-view step, rebuilt from the pieces that can run
here: the eager-torch restatement of the step (oracle.train_step: ~300 launches per scene query, every [M, .]
intermediate in HBM, cuBLAS fp32 GEMMs) with the hash-grid encodes executed by the UNMODIFIED reference CUDA kernel
(oracle/_ref/_gridencoder_ref*.so, compiled by oracle/build_ref.sh from external/encoders/gridencoder/src, wrapped as
grid.py:25-96 `_grid_encode` does) and torch restatements of the nerfacc compositing calls (nerfacc is not installable
offline).  This is the "reference GPU path on the same B200" of SURVEY.md 8d -- the denominator of the north star's
">= 10x the reference's single-GPU rays/s".

    python -m oracle.ref_gpu_step [n_rays] [steps]        (on a GPU box; prints one JSON line)
"""
import glob
import importlib.util
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def load_ref_backend():
    paths = glob.glob(os.path.join(HERE, '_ref', '_gridencoder_ref*.so'))
    if not paths:
        return None
    spec = importlib.util.spec_from_file_location('_gridencoder_ref', paths[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_grid_fn(backend):
    """autograd wrapper of the reference kernel, as external/encoders/gridencoder/grid.py:25-96"""

    class _grid_encode(torch.autograd.Function):
        @staticmethod
        def forward(ctx, inputs, embeddings, offsets, S, H, calc_grad_inputs, max_level):
            inputs = inputs.contiguous()
            B, D = inputs.shape
            L = offsets.shape[0] - 1
            C = embeddings.shape[1]
            ml = L if max_level is None else max(min(int(np.ceil(max_level * L)), L), 1)
            outputs = torch.zeros(L, B, C, device=inputs.device, dtype=embeddings.dtype)      # grid.py:50,53
            dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=embeddings.dtype) if calc_grad_inputs else None
            backend.grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, ml, S, H, dy_dx, 0, False, 0)
            ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
            ctx.dims = [B, D, C, L, S, H, ml]
            return outputs.permute(1, 0, 2).reshape(B, L * C)

        @staticmethod
        def backward(ctx, grad):
            inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
            B, D, C, L, S, H, ml = ctx.dims
            grad = grad.view(B, L, C).permute(1, 0, 2).contiguous()                            # grid.py:82
            grad_embeddings = torch.zeros_like(embeddings)
            grad_inputs = torch.zeros_like(inputs) if dy_dx is not None else None
            backend.grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, ml, S, H, dy_dx, grad_inputs, 0, False, 0)
            return grad_inputs, grad_embeddings, None, None, None, None, None

    return _grid_encode


def main():
    n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    backend = load_ref_backend()
    if backend is None:
        print(json.dumps({'impl': 'reference_gpu', 'unavailable': 'oracle/_ref not built (needs /root/reference at build time)'}))
        return
    dev = torch.device('cuda:0')
    torch.set_default_device(dev)                 # the oracle creates its helper tensors with factory defaults
    import bench
    from morpheus_b200.rays import synthetic_real_view_batch
    from oracle import fields as of
    from oracle import train_step as ots
    grid_fn = make_grid_fn(backend)

    def grid(self, which, x):                      # SceneOracle.grid with the reference kernel instead of the numpy restatement
        u = (x + self.bound) / (2 * self.bound)
        return grid_fn.apply(u, self.sd[which + '.embeddings'], self.sd[which + '.offsets'], self.S, self.H, u.requires_grad, self.max_level)

    of.SceneOracle.grid = grid
    with torch.device('cpu'):
        m = bench.make_state()
    params = ots.make_params({k: v.detach().to(dev).clone() for k, v in m.state_dict().items()})
    opt = torch.optim.Adam([v for v in params.values() if v.requires_grad], lr=5e-4, betas=(0.9, 0.99), eps=1e-15)
    with torch.device('cpu'):
        batches = [synthetic_real_view_batch(n_rays, seed=100 + r) for r in range(steps + 2)]
    batches = [{k: v.to(dev) for k, v in b.items()} for b in batches]
    times = []
    for r, batch in enumerate(batches):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        opt.zero_grad()
        loss, _ = ots.step_loss(params, batch, bench.N_SAMPLES, bench.MAX_LEVEL)
        loss.backward()
        opt.step()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times[2:]))
    print(json.dumps({'impl': 'reference_gpu', 'metric': 'rays_per_sec_train_step', 'value': n_rays / (ms * 1e-3), 'unit': 'rays/s',
                      'ms_per_step': ms, 'rays': n_rays, 'samples_per_ray': bench.N_SAMPLES, 'steps': steps, 'final_loss': float(loss),
                      'peak_mem_gb': torch.cuda.max_memory_allocated() / 1e9,
                      'what': 'eager torch step (oracle.train_step) + UNMODIFIED reference gridencoder CUDA kernel (oracle/_ref), fp32, 1x B200'}))


if __name__ == '__main__':
    main()

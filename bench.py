#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the MorpheuS render-and-loss hot path.

Default workload = BASELINE.json configs[1] (`--config cfg2`): snoopy-shaped synthetic RGB-D, 4096 rays x 128 samples per step,
real-view TRAINING step = pose correction -> fixed-S sampling -> fused scene query (deform + topology MLPs, hash grids, SDF + colour
MLPs) -> Laplace sigma -> alpha compositing -> fused FD-normal regulariser (12 more SDF queries per sample: the normals at x and at the
perturbed point, morpheus.py:714-741) -> loss heads -> fused backward -> [NCCL all-reduce of the flat gradient arena] -> fused Adam.
13 SDF queries per sample, the reference's real-view mix.  The whole step is ONE CUDA-graph replay.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg3|cfg4|cfg5]
    cfg2  4096 rays x 128 samples, real-view step (default; the configuration BASELINE.json's metric is quoted on)
    cfg4  8192 rays x 128 samples (teddy shape), same step, ray-sharded over the GPUs
    cfg3  72 x 72 novel view + Zero-1-to-3 SDS (seeded random weights: the checkpoint is not downloadable offline), occupancy sampling
    cfg5  the trainer's iteration mix: 1 virtual (SDS) step + 10 real steps of 2048 rays with every shipped loss term, occupancy
          sampling and the occupancy refresh every 16 steps (morpheus.py:1377-1424)
Under torchrun (N > 1) the ray batch / the novel view's rays are split over ranks (strong scaling), one all-reduce per optimiser step.
Prints ONE JSON line (rank 0).  `value` = rays/s with inputs resident in HBM; `e2e` = same step with the batch copied from pinned
host memory and the loss read back every step.  Timing: CUDA events around every step on the launching stream (sum over the K steps,
max over ranks); a 256 MiB L2 flush runs BETWEEN the timed steps, outside the event pairs.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RAYS, N_SAMPLES, NUM_FRAMES, MAX_LEVEL = 4096, 128, 200, 1.0
CONFIG = {'model': {'bg_radius': 1.4, 'activation': 'exp'}, 'render': {'step_size': 0.01}}
# algorithmic MACs per sample (SURVEY.md 8d / BASELINE.md section 2)
MAC_DEFORM, MAC_TOPO, MAC_COLOR, MAC_SDF = 77056, 76928, 8384, 10880
SDF_FD = 73 * 64 + 64 * 64 + 64           # an FD query only needs output row 0 of the last SDF layer


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('bf16_tflops_sustained', 1400.0), d.get('hbm_gbs', 6650.0), 'measured (MEASURED_PEAKS.json, sustained)'
    return 1400.0, 6650.0, 'fallback (B200_PROFILING.md)'


def csrc_sha():
    """fingerprint of the field-kernel sources (the kernels the roofline block reports on): profiles/r02_traffic.json records the one its
    ncu capture was taken at"""
    h = hashlib.sha1()
    d = os.path.join(ROOT, 'morpheus_b200', 'csrc')
    for f in sorted(os.listdir(d)):
        if f.startswith(('field_', 'tc_', 'common')):
            h.update(open(os.path.join(d, f), 'rb').read())
    return h.hexdigest()[:12]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={q}', '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace('.', '').isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith('active')})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons, 'samples': len(sm)}


def cpu_threads():
    """host threads for the CPU arm: all cores up to 32 (beyond that the oracle's many small torch ops get SLOWER from
    oversubscription: measured 1.2 rays/s at 128 threads vs ~20 rays/s at 8 on the same step)"""
    return min(os.cpu_count() or 1, 32)


def make_state():
    """seeded reference-style initial state (geometric-init SDF ~ sphere of radius 0.4, weight_norm g=|v|, U(-1e-4,1e-4) tables)"""
    torch.manual_seed(2024)   # morpheus.py:45 seed_everything(2024)
    from morpheus_b200.model import scene_representation
    m = scene_representation(CONFIG, 1.01, num_frames=NUM_FRAMES, deform_dim=16, use_app=False, use_t=False, amb_dim=2,
                             color_grid=True, use_joint=True, encode_topo=False)
    m.max_level = MAX_LEVEL
    return m


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the same step on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_step_seconds(state_dict, n_rays, repeats):
    """the ONLY place bench.py executes oracle/: the CPU checker timed as the CPU baseline."""
    from oracle import train_step as ots
    from morpheus_b200.rays import synthetic_real_view_batch
    torch.set_num_threads(cpu_threads())
    params = ots.make_params({k: v.detach().cpu().clone() for k, v in state_dict.items()})
    opt = torch.optim.Adam([v for v in params.values() if v.requires_grad], lr=5e-4, betas=(0.9, 0.99), eps=1e-15)
    times = []
    for r in range(repeats):
        batch = synthetic_real_view_batch(n_rays, seed=100 + r)
        t0 = time.perf_counter()
        opt.zero_grad()
        loss, _ = ots.step_loss(params, batch, N_SAMPLES, MAX_LEVEL)
        loss.backward()
        opt.step()
        times.append(time.perf_counter() - t0)
    return times


def run_reference(args, rank):
    """`--impl reference`: the reference's CPU arm = the oracle port of the same step (the reference's own hash-grid kernel is
    CUDA-only, so it has no CPU path; kind = "port") on all host cores.  Each step is a BOUNDED SAMPLE of the workload (64 rays x
    128 samples x 13 SDF queries); `ms_per_step` is the MEASURED time of that sample step, the extrapolation to the full ray batch is
    reported separately.  A batch-size scan shows the rays/s does not depend on the sample size."""
    if rank != 0:
        return
    if args.config in ('cfg3', 'cfg5'):
        print(json.dumps({'impl': 'reference', 'unavailable': f'{args.config}: the SDS leg needs the Zero-1-to-3 stack on the CPU (UNet 353 GFLOP + VAE 818 GFLOP per step); '
                          'only the real-view configs (cfg2, cfg4) have a CPU port arm'}))
        return
    m = make_state()
    n_rays = 64
    times = cpu_step_seconds(m.state_dict(), n_rays, args.warmup + args.steps)[args.warmup:]
    sec = sum(times) / len(times)
    val = n_rays / sec
    cores = cpu_threads()
    scan = {}
    for nr in (256, 1024):
        ts = cpu_step_seconds(m.state_dict(), nr, 1)
        scan[str(nr)] = {'rays_per_s': nr / ts[0], 'ms_per_step': ts[0] * 1e3}
    scan[str(n_rays)] = {'rays_per_s': val, 'ms_per_step': sec * 1e3}
    line = {'impl': 'reference', 'metric': 'rays_per_sec_train_step', 'value': val, 'unit': 'rays/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.config, args.gpus), 'run': {'sample_rays_per_step': n_rays},
            'extrapolated_ms_per_full_step': sec * 1e3 * (N_RAYS / n_rays),
            'cpu_baseline': {'value': val, 'unit': 'rays/s', 'cores': cores, 'kind': 'port',
                             'sample': f'{n_rays} rays x {N_SAMPLES} samples per step (1/{N_RAYS // n_rays} of the workload), fwd+bwd+Adam, torch CPU oracle port; '
                                       'the reference hash-grid kernel is CUDA-only so its own CPU path does not exist',
                             'batch_scan': scan},
            'e2e': {'value': val, 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def workload_config(cfg, n_gpus):
    flush = '256 MiB write between timed steps (outside the per-step CUDA-event pairs)'
    if cfg in ('cfg2', 'cfg4'):
        return {'workload': f'{cfg}: snoopy-shape synthetic RGB-D, {N_RAYS} rays x {N_SAMPLES} samples, real-view train step (albedo_normal + perturbed-normal '
                            'reg: 13 SDF queries/sample), fwd+bwd+allreduce+Adam', 'rays': N_RAYS, 'samples_per_ray': N_SAMPLES, 'frames': NUM_FRAMES,
                'parallelism': f'ray-sharded dp{n_gpus}', 'l2_flush': flush}
    if cfg == 'cfg3':
        return {'workload': 'cfg3: 72x72 novel view, occupancy-grid sampling (step 0.01), lambertian shading, Zero-1-to-3 SDS (seeded random weights), '
                            'fwd+bwd+allreduce+Adam', 'rays': 72 * 72, 'frames': NUM_FRAMES, 'parallelism': f'view rays sharded dp{n_gpus} + all-gather of pred_rgb',
                'l2_flush': flush}
    return {'workload': 'cfg5: trainer iteration = 1 virtual (SDS, 72x72) + 10 real steps of 2048 rays, every shipped loss term, occupancy sampling, '
                        'occupancy refresh every 16 steps', 'rays': 72 * 72 + 10 * 2048, 'frames': NUM_FRAMES, 'parallelism': f'rays sharded dp{n_gpus}',
            'l2_flush': flush}


def reference_gpu_leg():
    """the reference's GPU path on THIS box (oracle/ref_gpu_step.py: eager torch step + the unmodified reference gridencoder kernel from
    oracle/_ref), run as a separate process after our measurement -- the denominator of the north star's '>= 10x the reference's
    single-GPU rays/s'.  Skipped when oracle/_ref was not built."""
    import glob
    if not glob.glob(os.path.join(ROOT, 'oracle', '_ref', '_gridencoder_ref*.so')):
        return {'unavailable': 'oracle/_ref not built (needs /root/reference at build time)'}
    try:
        out = subprocess.run([sys.executable, '-m', 'oracle.ref_gpu_step', str(N_RAYS), '4'], cwd=ROOT, capture_output=True, text=True, timeout=600)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith('{'):
                d = json.loads(ln)
                return {k: d.get(k) for k in ('value', 'unit', 'ms_per_step', 'rays', 'samples_per_ray', 'steps', 'peak_mem_gb', 'what', 'unavailable') if k in d}
        return {'unavailable': 'no JSON from oracle.ref_gpu_step: ' + out.stderr[-300:]}
    except Exception as e:      # noqa: BLE001
        return {'unavailable': f'{type(e).__name__}: {e}'}


# ------------------------------------------------------------------------------------------------
# SDS pieces of cfg3 / cfg5
# ------------------------------------------------------------------------------------------------
def make_guidance(dev, precision):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from ldm_util import load_key_table, seeded_state
    from morpheus_b200 import guidance
    table = load_key_table()
    sd = {}
    sd.update(seeded_state(table['unet'], 1, 'model.diffusion_model.'))
    sd.update(seeded_state(table['encoder'], 2, 'first_stage_model.encoder.'))
    sd.update(seeded_state(table['quant_conv'], 3, 'first_stage_model.quant_conv.'))
    sd.update(seeded_state(table['cc_projection'], 4, 'cc_projection.'))
    z = guidance.Zero123(dev, state_dict=sd, t_range=[0.02, 0.5], precision=precision, graph=True)
    g = torch.Generator().manual_seed(2)
    emb = {'c_crossattn': [torch.randn(1, 1, 768, generator=g).to(dev)], 'c_concat': [torch.randn(1, 4, 32, 32, generator=g).to(dev)],
           'ref_radii': [2.5], 'ref_polars': [90.0], 'ref_azimuths': [0.0], 'zero123_ws': [1]}
    return z, emb


# ------------------------------------------------------------------------------------------------
def main():
    global N_RAYS
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='cfg2', choices=['cfg2', 'cfg3', 'cfg4', 'cfg5'])
    ap.add_argument('--cpu-baseline-steps', type=int, default=2)
    ap.add_argument('--no-graph', action='store_true', help='run the step eagerly instead of replaying the captured CUDA graph')
    ap.add_argument('--rays', type=int, default=None, help='global rays per step (overrides the config: cfg2 4096, cfg4 8192)')
    ap.add_argument('--full-step', action='store_true', help='cfg2/cfg4: add the SURVEY 8f rank-1 terms (normal smoothness on 11 band points/ray, '
                    'surface-point SDF/colour) to the step: the complete real-view iteration of morpheus.py:1147-1236')
    ap.add_argument('--sds-precision', default='fp32', choices=['fp32', 'reference', 'tf32'],
                    help="cfg3/cfg5: arithmetic of the frozen diffusion nets; 'reference' = what stock PyTorch gives the reference on this GPU (TF32 convolutions, fp32 matmuls)")
    ap.add_argument('--no-reference-gpu', action='store_true', help='skip the reference-GPU-path leg (oracle/ref_gpu_step.py, ~15 s)')
    args = ap.parse_args()
    if args.config == 'cfg4':
        N_RAYS = 8192
    if args.rays:
        N_RAYS = args.rays
    args.warmup = max(args.warmup, 3 if args.impl == 'ours' else 1)
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # NCCL prints its 'NCCL version ...' banner (and any NCCL_DEBUG output) on STDOUT when the first communicator is created:
        # route fd 1 to stderr while that happens, so that rank 0's stdout carries exactly ONE JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    if args.config in ('cfg3', 'cfg5'):
        line = run_virtual_configs(args, rank, world, dev)
    else:
        line = run_real_view(args, rank, world, local_rank, dev)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        # leave without destroy_process_group: tearing the communicator down while CUDA graphs that captured its collectives are still
        # alive can hang; the barrier makes sure every rank has finished its timed region first
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def timed_steps(step_fn, n_warm, n_steps, flush, world, dev, rank, local_rank, after_step=None):
    """W warm-up steps, then K steps each bracketed by a CUDA-event pair on the launching stream; the L2 flush sits between the pairs.
    -> (sum of the K step times in ms, max over ranks; clocks; last return value of step_fn)"""
    import torch.distributed as dist
    for s in range(n_warm):
        out = step_fn(s)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    evs = []
    for s in range(n_warm, n_warm + n_steps):
        flush.fill_(s & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = step_fn(s)
        if after_step is not None:
            after_step(out)
        e1.record()
        evs.append((e0, e1))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), clocks, out


def run_real_view(args, rank, world, local_rank, dev):
    import torch.distributed as dist  # noqa: F401
    from morpheus_b200 import model as mmodel
    from morpheus_b200 import train as mtrain
    from morpheus_b200.nerfacc_compat import OccGridEstimator
    from morpheus_b200.rays import synthetic_real_view_batch
    from morpheus_b200.render import Renderer
    model = make_state().to(dev).train()
    state_for_cpu = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()} if rank == 0 else None
    tr = dict(mtrain.FULL_TRAIN_CFG if args.full_step else mtrain.DEFAULT_TRAIN_CFG)
    cfg = dict(CONFIG, train=tr)
    est = OccGridEstimator(torch.tensor([-1.01] * 3 + [1.01] * 3), 128).to(dev).train()
    R = Renderer(model, est, cfg, NUM_FRAMES, uniform_samples=N_SAMPLES)
    R.world_size = world
    opt = mtrain.FlatAdam(model, tr['lr'])
    n_local = N_RAYS // world
    total_steps = args.warmup + args.steps
    # pinned host batches (one per step, distinct pixels/frames); each rank keeps its contiguous shard
    host = []
    for s in range(total_steps):
        b = synthetic_real_view_batch(N_RAYS, seed=1000 + s, frame=(37 * s) % NUM_FRAMES)
        shard = {k: v[rank * n_local:(rank + 1) * n_local].contiguous().pin_memory() for k, v in b.items()}
        if world > 1:      # global normaliser of the SDF band loss (utils.py:107), known to whoever shards the batch: mean over ranks of
            # count_nonzero(depth[ray_indices]) = (#rays with depth) * S / world  (see render.global_count)
            shard['n_depth'] = (torch.count_nonzero(b['depth']).float() * N_SAMPLES / world).reshape(1).pin_memory()
        host.append(shard)
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def to_dev(b):
        return {k: v.to(dev, non_blocking=True) for k, v in b.items()}

    graphed = None if args.no_graph else mtrain.GraphedStep(R, opt, to_dev(host[0]), tr, world)
    prof = mmodel.PROFILE

    def one_step(b):
        if graphed is not None:
            return graphed.step(b)
        return mtrain.train_step(R, opt, b if b['rays_o'].is_cuda else to_dev(b), tr, world)

    resident = [to_dev(b) for b in host]
    torch.cuda.synchronize()
    ms_total, clocks, last_loss = timed_steps(lambda s: one_step(resident[s]), args.warmup, args.steps, flush, world, dev, rank, local_rank)
    last_loss = float(last_loss)
    del resident
    loss_host = torch.zeros(1).pin_memory()

    def read_back(loss):
        loss_host.copy_(loss.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the user reads the loss every step (morpheus.py:1426 loss.item())
    ms_e2e, _, _ = timed_steps(lambda s: one_step(host[s]), args.warmup, args.steps, flush, world, dev, rank, local_rank, after_step=read_back)

    # per-kernel CUDA-event timing of OUR launches: the same steps run eagerly (events cannot be read back from a replayed graph),
    # on the launching stream, after warm-up
    res2 = [to_dev(b) for b in host[:args.warmup + min(args.steps, 5)]]
    for b in res2[:args.warmup]:
        mtrain.train_step(R, opt, b, tr, world)
    torch.cuda.synchronize()
    prof.enabled = True
    prof.reset()
    for b in res2[args.warmup:]:
        flush.fill_(1)
        mtrain.train_step(R, opt, b, tr, world)
    kern = prof.summary()
    prof.enabled = False
    kern_steps = len(res2) - args.warmup

    # occupancy refresh (morpheus.py:905-913): the full 128^3 = 2 097 152-cell refresh of the first 256 steps, every 16th step
    occ_ms = None
    if rank == 0:
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            R.update_occ_grid(res2[0]['rays_t'].reshape(-1, 1), step=0)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        occ_ms = sorted(ts)[1]
    if rank != 0:
        return None
    ms_step = ms_total / args.steps
    value = N_RAYS / (ms_step * 1e-3)
    e2e_val = N_RAYS / (ms_e2e / args.steps * 1e-3)
    tf_peak, hbm_peak, which = peaks()
    # algorithmic FLOPs per launch (2 per MAC; backward = dgrad + wgrad = 2x forward; recompute and the 3x fp16 split are NOT counted)
    M_local = n_local * N_SAMPLES
    fused_fd = 'fd_regulariser' in kern
    flops = {
        'field_fwd_main': 2 * (MAC_DEFORM + MAC_TOPO + MAC_COLOR + MAC_SDF + (0 if fused_fd else 6 * SDF_FD)) * M_local,
        'fd_regulariser': 6 * 12 * SDF_FD * M_local,          # forward + backward of the 12 FD queries of a sample
        'field_fwd_aux': 2 * 6 * SDF_FD * M_local,
        'field_bwd_sdf_tc_main': 4 * (MAC_COLOR + MAC_SDF) * M_local,
        'field_bwd_fd_tc_main': 4 * 6 * SDF_FD * M_local,
        'field_bwd_fd_tc_aux': 4 * 6 * SDF_FD * M_local,
        'field_bwd_warp_tc': 4 * (MAC_DEFORM + MAC_TOPO) * M_local,
    }
    # per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) from the `ncu --set full` captures of the SAME kernel sources
    # (profiles/r02_traffic.json records the csrc fingerprint it was taken at; a stale file reads as null)
    traffic, traffic_note = {}, 'no profiles/r02_traffic.json'
    tp = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get('csrc_sha') == csrc_sha() and tj.get('M') == M_local:
            traffic, traffic_note = tj.get('kernels', {}), f"ncu --set full at csrc {tj.get('csrc_sha')} ({tj.get('command', '')})"
        else:
            traffic_note = f"stale: captured at csrc {tj.get('csrc_sha')} / M {tj.get('M')}, sources are now {csrc_sha()} / M {M_local}"
    rooflines = []
    for name, fl in flops.items():
        if name in kern and not args.full_step:      # (--full-step adds launches of other sizes under the same names)
            ach = fl / (kern[name]['avg_ms'] * 1e-3) / 1e12
            rooflines.append({'kernel': name, 'engine': 'tcgen05 (3x fp16 split)', 'bound': 'tensor', 'achieved': ach, 'peak': tf_peak, 'unit': 'TFLOP/s',
                              'frac': ach / tf_peak, 'traffic': traffic.get(name), 'avg_launch_ms': kern[name]['avg_ms'], 'launches_timed': kern[name]['n'],
                              'algorithmic_flops_per_launch': fl})
    rooflines.sort(key=lambda r: -r['avg_launch_ms'])
    roof = None
    if rooflines:
        roof = dict(rooflines[0], peak_source=which, traffic_source=traffic_note,
                    note='dominant kernel by measured time; achieved = ALGORITHMIC flops (2/MAC, backward = 2x forward, recompute and the 3x fp16 hi/lo '
                         'split of every product not counted) / CUDA-event launch time (eager pass of the same steps); the kernel is bound by the '
                         'table scatter (LSU atomic issue), L2 gather latency and CUDA-core epilogues, not by the tensor pipe (profiles/); all kernels in "rooflines"')
    launches = sum(v['n'] for v in kern.values()) // max(kern_steps, 1) if kern else None
    cpu_val = None
    if args.cpu_baseline_steps > 0 and world == 1:
        times = cpu_step_seconds(state_for_cpu, 64, 1 + args.cpu_baseline_steps)[1:]
        cpu_val = 64 / (sum(times) / len(times))
    ref_gpu = None
    nccl_flag = bool(getattr(graphed, 'nccl_in_graph', False)) if (graphed is not None and world > 1) else None
    if world == 1 and not args.no_reference_gpu and not args.full_step:
        graphed = None
        torch.cuda.empty_cache()
        ref_gpu = reference_gpu_leg()
        if ref_gpu.get('value'):
            ref_gpu['ours_over_reference_gpu'] = value / ref_gpu['value']
    line = {'metric': 'rays_per_sec_train_step', 'value': value, 'unit': 'rays/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.config, world),
            'run': {'cuda_graph': (not args.no_graph), 'full_step': bool(args.full_step), 'nccl_in_graph': nccl_flag},
            'clocks': clocks,
            'e2e': {'value': e2e_val, 'unit': 'rays/s', 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': (launches * args.steps if launches else launches), 'gpu_launches_per_step': launches, 'kernels': kern, 'final_loss': last_loss,
            'roofline': roof, 'rooflines': rooflines,
            'occ_refresh_ms': occ_ms, 'occ_refresh_ms_amortised_per_step': (occ_ms / 16.0 if occ_ms else None),
            'cpu_baseline': {'value': cpu_val, 'unit': 'rays/s', 'cores': cpu_threads(), 'kind': 'port',
                             'sample': f'64 rays x {N_SAMPLES} samples (1/{N_RAYS // 64} of the step), fwd+bwd+Adam, {args.cpu_baseline_steps} timed steps after 1 warm-up, oracle port on torch CPU'},
            'reference_gpu': ref_gpu}
    return line


def run_virtual_configs(args, rank, world, dev):
    """cfg3: the virtual-view (SDS) step; cfg5: the trainer's iteration mix (1 virtual + 10 real steps, morpheus.py:1377-1424)."""
    import torch.distributed as dist  # noqa: F401
    from morpheus_b200 import rays as mrays
    from morpheus_b200 import train as mtrain
    from morpheus_b200.nerfacc_compat import OccGridEstimator
    from morpheus_b200.rays import synthetic_real_view_batch
    from morpheus_b200.render import Renderer
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    model = make_state().to(dev).train()
    model.max_level = 0.75
    tr = dict(mtrain.FULL_TRAIN_CFG)
    cfg = dict(CONFIG, train=tr)
    est = OccGridEstimator(torch.tensor([-1.01] * 3 + [1.01] * 3), 128).to(dev).train()
    R = Renderer(model, est, cfg, NUM_FRAMES)
    R.world_size = world
    opt = mtrain.FlatAdam(model, tr['lr'])
    z123, emb = make_guidance(dev, args.sds_precision)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    real_rays = 2048
    n_local = real_rays // world
    state = {'step': 0}

    def virtual(s):
        view = mrays.virtual_view_rays(frame=(7 * s) % NUM_FRAMES, num_frames=NUM_FRAMES, H=360, W=360, focal=517.0, scale=0.2,
                                       generator=torch.Generator().manual_seed(s), device=dev)
        R.update_occ_grid(view['rays_t'].reshape(-1, 1)[:1], step=state['step'])
        state['step'] += 1
        loss, _ = mtrain.virtual_view_step(R, z123, opt, view, emb, tr, shading='lambertian', ambient_ratio=0.4, bg_color=torch.rand(3, device=dev),
                                           world_size=world)
        return loss

    def real(s, i):
        b = synthetic_real_view_batch(real_rays, seed=5000 + 16 * s + i, frame=(37 * s + i) % NUM_FRAMES)
        b = {k: v[rank * n_local:(rank + 1) * n_local].contiguous().to(dev, non_blocking=True) for k, v in b.items()}
        R.update_occ_grid(b['rays_t'].reshape(-1, 1)[:1], step=state['step'])
        state['step'] += 1
        return mtrain.train_step(R, opt, b, tr, world)

    def step_fn(s):
        loss = virtual(s)
        if args.config == 'cfg5':
            for i in range(10):
                loss = real(s, i)
        return loss

    ms_total, clocks, last = timed_steps(step_fn, args.warmup, args.steps, flush, world, dev, rank, local_rank)
    from morpheus_b200 import _lib
    _lib.PROFILE.enabled = True
    _lib.PROFILE.reset()
    step_fn(args.warmup + args.steps)
    kern = _lib.PROFILE.summary()
    _lib.PROFILE.enabled = False
    own_launches = sum(v['n'] for v in kern.values())
    if rank != 0:
        return None
    rays = 72 * 72 + (10 * real_rays if args.config == 'cfg5' else 0)
    ms_step = ms_total / args.steps
    tf_peak, hbm_peak, which = peaks()
    # roofline of the SDS leg: the UNet streams its 3.44 GB of fp32 weights once per call (batch 2): an HBM floor of 3.44 GB / peak
    unet_bytes = sum(v.numel() * v.element_size() for v in z123.unet.values())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pred = torch.rand(1, 3, 72, 72, device=dev, requires_grad=True)
    pol, az, rad = torch.tensor([10.0]), torch.tensor([30.0]), torch.tensor([0.1])
    for _ in range(3):
        z123.train_step(emb, pred, pol, az, rad, guidance_scale=5, grad_scale=0.01)[0].backward()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        z123.train_step(emb, pred, pol, az, rad, guidance_scale=5, grad_scale=0.01)[0].backward()
    e1.record()
    torch.cuda.synchronize()
    sds_ms = e0.elapsed_time(e1) / 5
    sds_flops = 352.7e9 + 272.7e9 * 2          # UNet forward (batch 2) + VAE encoder forward + input-gradient backward
    line = {'metric': 'rays_per_sec_train_step', 'value': rays / (ms_step * 1e-3), 'unit': 'rays/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.config, world), 'run': {'sds_precision': args.sds_precision, 'sds_chain_cuda_graph': True, 'max_level': 0.75},
            'clocks': clocks, 'final_loss': float(last),
            'e2e': {'value': rays / (ms_step * 1e-3), 'unit': 'rays/s', 'h2d_bytes_per_step': (10 * n_local * 68 if args.config == 'cfg5' else 0), 'd2h_bytes_per_step': 0,
                    'note': 'rays are generated on the device (virtual views) / copied from host batches inside the timed step (real views of cfg5)'},
            'sds_chain': {'ms': sds_ms, 'what': 'Zero123.train_step forward + backward to pred_rgb, one CUDA-graph replay (resize, VAE encoder, add-noise, UNet x2 CFG, '
                          'SDS gradient, VAE input-gradient)', 'algorithmic_gflop': sds_flops / 1e9, 'achieved_tflops': sds_flops / (sds_ms * 1e-3) / 1e12,
                          'kernels': 'fp32 mode: 3x3 / 1x1 stride-1 convolutions of the UNet and the VAE (forward and input-gradient) on the own tcgen05 kernel (csrc/conv_tc.cu); '
                                     'linear layers, attention, normalisations, stride-2 / first / last convolutions on cuBLAS / SDPA / cuDNN through torch'},
            'roofline': {'kernel': 'Zero-1-to-3 UNet forward (batch 2, inside the SDS chain graph)', 'bound': 'hbm', 'achieved': unet_bytes / (sds_ms * 1e-3) / 1e9,
                         'peak': hbm_peak, 'unit': 'GB/s', 'frac': unet_bytes / (sds_ms * 1e-3) / 1e9 / hbm_peak, 'traffic': None, 'peak_source': which,
                         'note': f'lower bound: {unet_bytes / 1e9:.2f} GB of UNet weights streamed once per step over the WHOLE chain time (the chain also runs the '
                                 'VAE encoder forward + backward, 818 GFLOP of convolutions)'},
            'gpu_launches': own_launches * args.steps, 'gpu_launches_per_step': own_launches, 'kernels': kern,
            'cpu_baseline': None}
    return line


if __name__ == '__main__':
    main()

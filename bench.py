#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the MorpheuS render-and-loss hot path.

Workload (BASELINE.json configs[1]): snoopy-shaped synthetic RGB-D, 4096 rays x 128 samples per step,
real-view TRAINING step = pose correction -> fixed-S sampling -> fused scene query (deform + topology MLPs,
hash grids, SDF + colour MLPs, 6-point FD normals: 'albedo_normal') -> Laplace sigma -> alpha compositing
-> perturbed-normal query (6 more SDF queries/sample) -> losses -> fused backward -> [NCCL all-reduce of the
flat gradient arena] -> fused Adam.  13 SDF queries per sample, the reference's real-view mix.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
Under torchrun (N > 1) the global ray batch is split over ranks (strong scaling), one all-reduce per step.
Prints ONE JSON line (rank 0).  `value` = rays/s with inputs resident in HBM; `e2e` = same step with the
batch copied from pinned host memory and the loss read back every step.
"""
import argparse
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


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('bf16_tflops_sustained', 1400.0), d.get('hbm_gbs', 6650.0), 'measured (MEASURED_PEAKS.json, sustained)'
    return 1400.0, 6650.0, 'fallback (B200_PROFILING.md)'


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
    if rank != 0:
        return
    m = make_state()
    n_rays = 64          # bounded sample of the 4096-ray step: 64 rays x 128 samples x 13 SDF queries
    times = cpu_step_seconds(m.state_dict(), n_rays, args.warmup + args.steps)[args.warmup:]
    sec = sum(times) / len(times)
    val = n_rays / sec
    cores = cpu_threads()
    line = {'impl': 'reference', 'metric': 'rays_per_sec_train_step', 'value': val, 'unit': 'rays/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': sec * 1e3 * (N_RAYS / n_rays), 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.gpus),
            'cpu_baseline': {'value': val, 'unit': 'rays/s', 'cores': cores, 'kind': 'port',
                             'sample': f'{n_rays} rays x {N_SAMPLES} samples per step (1/{N_RAYS // n_rays} of the workload), fwd+bwd+Adam, torch CPU oracle port; '
                                       'the reference hash-grid kernel is CUDA-only so its own CPU path does not exist'},
            'e2e': {'value': val, 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def workload_config(n_gpus):
    return {'workload': f'snoopy-shape synthetic RGB-D, {N_RAYS} rays x {N_SAMPLES} samples, real-view train step (albedo_normal + perturbed-normal reg: '
                        '13 SDF queries/sample), fwd+bwd+allreduce+Adam', 'rays': N_RAYS, 'samples_per_ray': N_SAMPLES, 'frames': NUM_FRAMES,
            'parallelism': f'ray-sharded dp{n_gpus}', 'l2_flush': '256 MiB write between timed steps (inside the timed region)'}


# ------------------------------------------------------------------------------------------------
def main():
    global N_RAYS
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--cpu-baseline-steps', type=int, default=2)
    ap.add_argument('--no-graph', action='store_true', help='run the step eagerly instead of replaying the captured CUDA graph')
    ap.add_argument('--rays', type=int, default=N_RAYS, help='global rays per step (BASELINE cfg-2: 4096; cfg-4 teddy shape: 8192)')
    ap.add_argument('--full-step', action='store_true', help='add the SURVEY 8f rank-1 terms (normal smoothness on 11 band points/ray, '
                    'surface-point SDF/colour) to the step: the complete real-view iteration of morpheus.py:1147-1236')
    args = ap.parse_args()
    N_RAYS = args.rays
    args.warmup = max(args.warmup, 3 if args.impl == 'ours' else 1)
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    import torch.distributed as dist
    from morpheus_b200 import train as mtrain
    from morpheus_b200.nerfacc_compat import OccGridEstimator
    from morpheus_b200.rays import synthetic_real_view_batch
    from morpheus_b200.render import Renderer
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
    model = make_state().to(dev).train()
    state_for_cpu = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()} if rank == 0 else None
    tr = dict(mtrain.FULL_TRAIN_CFG if args.full_step else mtrain.DEFAULT_TRAIN_CFG)
    cfg = dict(CONFIG, train=tr)
    R = Renderer(model, OccGridEstimator(torch.tensor([-1.01] * 3 + [1.01] * 3), 128).to(dev), cfg, NUM_FRAMES, uniform_samples=N_SAMPLES)
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graphed = None if args.no_graph else mtrain.GraphedStep(R, opt, to_dev(host[0]), tr, world)

    def one_step(b):
        if graphed is not None:
            return graphed.step(b)
        return mtrain.train_step(R, opt, b if b['rays_o'].is_cuda else to_dev(b), tr, world)

    def run(e2e, profile):
        """K timed steps.  e2e=False: batches already resident in HBM (value); e2e=True: every step copies its batch from
        pinned host memory and reads the loss back (what a user of train_step pays)."""
        resident = [to_dev(b) for b in host] if not e2e else None
        torch.cuda.synchronize()
        loss_host = torch.zeros(1).pin_memory()
        for s in range(args.warmup):
            loss = one_step(host[s] if e2e else resident[s])
        barrier()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        if sampler:
            sampler.start()
        prof.enabled = profile
        prof.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(args.warmup, total_steps):
            flush.fill_(s & 0xFF)
            loss = one_step(host[s] if e2e else resident[s])
            if e2e:
                loss_host.copy_(loss.reshape(1), non_blocking=True)
                torch.cuda.current_stream().synchronize()   # the user reads the loss every step (morpheus.py:1426 loss.item())
        e1.record()
        barrier()
        prof.enabled = False
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), clocks, float(loss)

    def kernel_pass():
        """per-kernel CUDA-event timing of OUR launches: the same steps run eagerly (events cannot be read back from a
        replayed graph), on the launching stream, after warm-up"""
        resident = [to_dev(b) for b in host[:args.warmup + min(args.steps, 5)]]
        for b in resident[:args.warmup]:
            mtrain.train_step(R, opt, b, tr, world)
        torch.cuda.synchronize()
        prof.enabled = True
        prof.reset()
        for b in resident[args.warmup:]:
            flush.fill_(1)
            mtrain.train_step(R, opt, b, tr, world)
        out = prof.summary()
        prof.enabled = False
        return out, len(resident) - args.warmup

    from morpheus_b200 import model as mmodel
    prof = mmodel.PROFILE
    ms_total, clocks, last_loss = run(e2e=False, profile=False)
    ms_e2e, _, _ = run(e2e=True, profile=False)
    kern, kern_steps = kernel_pass()
    if rank == 0:
        ms_step = ms_total / args.steps
        value = N_RAYS / (ms_step * 1e-3)
        e2e_val = N_RAYS / (ms_e2e / args.steps * 1e-3)
        tf_peak, hbm_peak, which = peaks()
        # algorithmic FLOPs per launch (2 per MAC; backward = dgrad + wgrad = 2x forward; recompute is NOT counted)
        M_local = n_local * N_SAMPLES
        sdf_fd = 73 * 64 + 64 * 64 + 64           # an FD query only needs output row 0 of the last SDF layer
        tc_bwd = 'field_bwd_warp_tc' in kern
        sdf_bwd_main = 4 * (MAC_COLOR + MAC_SDF + 6 * sdf_fd) * M_local
        flops = {
            'field_fwd_main': 2 * (MAC_DEFORM + MAC_TOPO + MAC_COLOR + MAC_SDF + 6 * sdf_fd) * M_local,
            'field_fwd_aux': 2 * 6 * sdf_fd * M_local,
            'field_bwd_main': (0 if tc_bwd else 4 * (MAC_DEFORM + MAC_TOPO) * M_local) + sdf_bwd_main,
            'field_bwd_aux': 4 * 6 * sdf_fd * M_local,
            'field_bwd_sdf_tc_main': sdf_bwd_main if 'field_bwd_fd_tc_main' not in kern else 4 * (MAC_COLOR + MAC_SDF) * M_local,
            'field_bwd_sdf_tc_aux': 4 * 6 * sdf_fd * M_local,
            'field_bwd_fd_tc_main': 4 * 6 * sdf_fd * M_local,
            'field_bwd_fd_tc_aux': 4 * 6 * sdf_fd * M_local,
            'field_bwd_warp_tc': 4 * (MAC_DEFORM + MAC_TOPO) * M_local,
        }
        engine = {k: 'tcgen05 (3x fp16 split)' for k in flops}
        engine['field_bwd_main'] = engine['field_bwd_aux'] = 'fp32 SIMT'
        # per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the `ncu --set full` captures summarised under
        # profiles/ (tools/summarize_profiles.py traffic): same M and flags as the bench's main launches
        traffic = {}
        tp = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tp):
            traffic = json.load(open(tp))
        rooflines = []
        for name, fl in flops.items():
            if name in kern and not args.full_step:      # (--full-step adds launches of other sizes under the same names)
                ach = fl / (kern[name]['avg_ms'] * 1e-3) / 1e12
                rooflines.append({'kernel': name, 'engine': engine[name], 'bound': 'tensor', 'achieved': ach, 'peak': tf_peak, 'unit': 'TFLOP/s',
                                  'frac': ach / tf_peak, 'traffic': traffic.get(name), 'avg_launch_ms': kern[name]['avg_ms'], 'launches_timed': kern[name]['n'],
                                  'algorithmic_flops_per_launch': fl})
        rooflines.sort(key=lambda r: -r['avg_launch_ms'])
        roof = None
        if rooflines:
            roof = dict(rooflines[0], peak_source=which,
                        note='dominant kernel by measured time; achieved = ALGORITHMIC flops (2/MAC, backward = 2x forward, recompute and the 3x fp16 hi/lo '
                             'split of every product not counted) / CUDA-event launch time; the kernel is bound by L2 gather/scatter latency and CUDA-core '
                             'epilogues, not by the tensor pipe (5 % active, profiles/); traffic = DRAM bytes per launch from ncu; all kernels in "rooflines"')
        launches = sum(v['n'] for v in kern.values()) // max(kern_steps, 1) if kern else None
        cpu_val = None
        if args.cpu_baseline_steps > 0:
            times = cpu_step_seconds(state_for_cpu, 64, 1 + args.cpu_baseline_steps)[1:]
            cpu_val = 64 / (sum(times) / len(times))
        line = {'metric': 'rays_per_sec_train_step', 'value': value, 'unit': 'rays/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': dict(workload_config(world), cuda_graph=(not args.no_graph), full_step=bool(args.full_step)), 'clocks': clocks,
                'e2e': {'value': e2e_val, 'unit': 'rays/s', 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e / args.steps},
                'gpu_launches': (launches * args.steps if launches else launches), 'gpu_launches_per_step': launches, 'kernels': kern, 'final_loss': last_loss,
                'roofline': roof, 'rooflines': rooflines,
                'cpu_baseline': {'value': cpu_val, 'unit': 'rays/s', 'cores': cpu_threads(), 'kind': 'port',
                                 'sample': f'64 rays x {N_SAMPLES} samples (1/64 of the step), fwd+bwd+Adam, {args.cpu_baseline_steps} timed steps after 1 warm-up, oracle port on torch CPU'}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

"""Turn gpurun_out/ ncu artefacts into the tracked markdown summaries under profiles/.

  python tools/summarize_profiles.py launches gpurun_out/launches_nograph.csv profiles/rNN_launches_summary.md "<command>"
  python tools/summarize_profiles.py kernel   gpurun_out/x.ncu-rep            profiles/rNN_<kernel>_summary.md "<command>"
  python tools/summarize_profiles.py traffic  gpurun_out/x.ncu-rep                       # -> JSON {kernel, dram_bytes, ms}
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

KEEP = re.compile(r'^(dram__bytes_read\.sum|dram__bytes_write\.sum|gpu__time_duration\.sum|launch__grid_size|launch__block_size|'
                  r'launch__registers_per_thread|launch__occupancy_limit_(registers|shared_mem)|smsp__inst_executed\.sum|'
                  r'sm__issue_active\.avg\.pct_of_peak_sustained_elapsed|sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_elapsed|'
                  r'sm__warps_active\.avg\.pct_of_peak_sustained_active|l1tex__t_sector_hit_rate\.pct|lts__t_sector_hit_rate\.pct|'
                  r'l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|'
                  r'dram__throughput\.avg\.pct_of_peak_sustained_elapsed|smsp__inst_executed_op_global_red\.sum|sm__cycles_elapsed\.max|'
                  r'smsp__average_warps_issue_stalled_(long_scoreboard|barrier|wait|short_scoreboard|no_instruction|branch_resolving|mio_throttle|lg_throttle)_per_issue_active\.ratio)$')


def raw_metrics(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        res.append(d)
    return res


def to_bytes(v, u):
    x = float(v.replace(',', ''))
    return x * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)


def to_ms(v, u):
    x = float(v.replace(',', ''))
    return x * {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1, 'msecond': 1, 's': 1e3, 'second': 1e3}.get(u, 1)


def cmd_kernel(rep, dst, command):
    ms = raw_metrics(rep)
    lines = [f'# ncu --set full capture: `{rep.split("/")[-1]}`', '', f'Command (under gpurun, 1x B200): `{command}`', '']
    for d in ms:
        name = d.get('Kernel Name', ('?', ''))[0]
        lines += [f'## `{name[:110]}`', '', '| metric | value | unit |', '|---|---:|---|']
        for h, (v, u) in d.items():
            if KEEP.search(h):
                lines.append(f'| {h} | {v} | {u} |')
        lines.append('')
    # per-line stall samples
    out = subprocess.run([sys.executable, __file__.replace('summarize_profiles.py', 'ncu_lines.py'), rep, '14'], capture_output=True, text=True).stdout
    lines += ['Source lines with the most warp-stall samples (`tools/ncu_lines.py`, -lineinfo):', '', '```', out.strip(), '```', '']
    open(dst, 'w').write('\n'.join(lines))
    print('wrote', dst)


def cmd_traffic(rep):
    for d in raw_metrics(rep):
        b = to_bytes(*d['dram__bytes_read.sum']) + to_bytes(*d['dram__bytes_write.sum'])
        print(json.dumps({'kernel': d['Kernel Name'][0][:80], 'dram_bytes': b, 'ms': to_ms(*d['gpu__time_duration.sum'])}))


BENCH_NAMES = {'fd_reg_tc_kernel': 'fd_regulariser', 'field_fwd_tc_kernel': 'field_fwd_main', 'field_bwd_sdf_tc_kernel': 'field_bwd_sdf_tc_main',
               'field_bwd_warp_tc_kernel': 'field_bwd_warp_tc', 'field_bwd_fd_tc_kernel': 'field_bwd_fd_tc_main'}


def cmd_traffic_json(rep, dst, M, command):
    """profiles/rNN_traffic.json for bench.py's roofline.traffic: {csrc_sha, M, command, kernels: {bench kernel name: DRAM bytes per launch}}"""
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    kernels, detail = {}, {}
    for d in raw_metrics(rep):
        name = d['Kernel Name'][0]
        key = next((v for k, v in BENCH_NAMES.items() if k in name), None)
        if key is None or key in kernels:
            continue
        kernels[key] = to_bytes(*d['dram__bytes_read.sum']) + to_bytes(*d['dram__bytes_write.sum'])
        detail[key] = {'dram_read': to_bytes(*d['dram__bytes_read.sum']), 'dram_write': to_bytes(*d['dram__bytes_write.sum']),
                       'ncu_ms': to_ms(*d['gpu__time_duration.sum'])}
    json.dump({'csrc_sha': bench.csrc_sha(), 'M': int(M), 'command': command, 'kernels': kernels, 'detail': detail,
               'what': 'dram__bytes_read.sum + dram__bytes_write.sum per launch from one ncu --set full capture'}, open(dst, 'w'), indent=1)
    print('wrote', dst, kernels)


def cmd_launches(path, dst, command):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        rows.append((row['Kernel Name'], to_ms(row['Metric Value'], row['Metric Unit'])))
    adam = [i for i, (k, _) in enumerate(rows) if 'adam_kernel' in k or 'adam_groups_kernel' in k]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for k, ms in rows:
        tot[k.split('(')[0][-70:]][0] += 1
        tot[k.split('(')[0][-70:]][1] += ms
    T = sum(v[1] for v in tot.values())
    n_steps = len(adam)
    per_step = (adam[-1] - adam[-2]) if n_steps >= 2 else len(rows)
    ours = sum(v[1] for k, v in tot.items() if 'mb::' in k or k.startswith(('tc', 'mb', 'tcs', 'tcb', 'tcr', 'tcf')))
    out = [f'# ncu launch list: `{path.split("/")[-1]}`', '', f'Command (under gpurun, 1x B200): `{command}`', '',
           f'{len(rows)} launches, {n_steps} optimiser steps captured, {per_step} launches per steady-state step; total device time {T:.1f} ms '
           f'(cold-cache, serialised: compare SHARES); our kernels {100 * ours / T:.1f}% of device time.', '',
           '| kernel | launches | total ms | share |', '|---|---:|---:|---:|']
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:18]:
        out.append(f'| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / T:.1f}% |')
    out.append('')
    open(dst, 'w').write('\n'.join(out))
    print('wrote', dst)


if __name__ == '__main__':
    mode = sys.argv[1]
    if mode == 'kernel':
        cmd_kernel(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else '')
    elif mode == 'traffic':
        cmd_traffic(sys.argv[2])
    elif mode == 'traffic-json':
        cmd_traffic_json(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else '')
    else:
        cmd_launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else '')

"""Aggregate an ncu report's per-line stall samples:  python tools/ncu_lines.py report.ncu-rep [topN]
(reads `ncu --page source --csv --print-source cuda,sass`; prints the source lines with most samples + stall mix)."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; hdr = None
agg = collections.OrderedDict()
tot = 0
for r in rows:
    if not r: continue
    if r[0] == 'File Name': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] == '' : continue       # SASS row
    d = dict(zip(hdr, r))
    try: s = int(d['# Samples'])
    except Exception: continue
    inst = int(d['Instructions Executed']) if d['Instructions Executed'].isdigit() else 0
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k and v.isdigit() and int(v) > 0}
    agg[(cur_file, r[0])] = (s, inst, stalls, r[1].strip()[:110])
    tot += s
print('total samples', tot)
for (f, ln), (s, inst, st, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    mix = ' '.join(f'{k}:{v}' for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
    print(f'{100*s/tot:5.1f}% {inst:>10d} {f}:{ln:<5s} {src}\n        [{mix}]')

#!/usr/bin/env python
"""Turn the ncu exports of tools/ncu_zgemm.sh into the tracked evidence under profiles/:
  profiles/<tag>_ncu_zgemm_metrics.csv   one row per profiled launch, the metrics DESIGN.md / bench.py quote
  profiles/<tag>_launch_shares.csv       device time per kernel over one timed forward (share of the step)
  profiles/traffic.json                  per-kernel DRAM bytes per launch + tensor-pipe % that bench.py attaches to `roofline`
usage: python tools/ncu_summarize.py <tag>      (reads gpurun_out/<tag>_prof_zgemm_raw.csv and gpurun_out/<tag>_launches_cfg2_depth1.csv)"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else 'r02'
WANT = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active']


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


raw = os.path.join(ROOT, 'gpurun_out', f'{tag}_prof_zgemm_raw.csv')
traffic_path = os.path.join(ROOT, 'profiles', 'traffic.json')
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
if os.path.exists(raw):
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = os.path.join(ROOT, 'profiles', f'{tag}_ncu_zgemm_metrics.csv')
    per_kernel = collections.defaultdict(list)
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['kernel'] + [f'{m} [{units[idx[m]]}]' for m in WANT if m in idx])
        for r in data:
            name = re.sub(r'\(.*', '', r[idx['Kernel Name']]).replace('void ', '').strip()
            w.writerow([name] + [r[idx[m]] for m in WANT if m in idx])
            per_kernel[name].append(r)
    z = {}
    for name, rs in per_kernel.items():
        dram = [to_bytes(r[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']]) +
                to_bytes(r[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']]) for r in rs]
        tp = [float(r[idx['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']]) for r in rs]
        z[name] = dict(launches_profiled=len(rs), dram_bytes_per_launch=sum(dram) / len(dram), tensor_pipe_pct=sum(tp) / len(tp))
    conv = {k: v for k, v in z.items() if not k.startswith('zgemm_kernel<4')}       # MODE 4 is LinearSE3, timed as `linear`
    z_all, z = z, (conv or z)
    traffic['zgemm'] = dict(source=f'profiles/{tag}_ncu_zgemm_metrics.csv (ncu --set full, cfg2 depth-1 slice, same launches as the headline)',
                            per_variant=z_all,
                            dram_bytes_per_launch=sum(v['dram_bytes_per_launch'] * v['launches_profiled'] for v in z.values()) / max(1, sum(v['launches_profiled'] for v in z.values())),
                            sm__pipe_tensor_cycles_active_pct=sum(v['tensor_pipe_pct'] * v['launches_profiled'] for v in z.values()) / max(1, sum(v['launches_profiled'] for v in z.values())))
    print('wrote', out)
ll = os.path.join(ROOT, 'gpurun_out', f'{tag}_launches_cfg2_depth1.csv')
if os.path.exists(ll):
    rows = list(csv.reader(open(ll)))
    start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[start]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start + 1:]:
        if len(r) < len(hdr) or r[idx['Metric Name']] != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'[<(].*', '', r[idx['Kernel Name']]).replace('void ', '').strip()
        v = float(r[idx['Metric Value']].replace(',', ''))
        u = r[idx['Metric Unit']]
        v = v / 1e6 if u.startswith('ns') else v / 1e3 if u.startswith('us') else v
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    out = os.path.join(ROOT, 'profiles', f'{tag}_launch_shares.csv')
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['kernel', 'launches', 'device_ms (ncu, cold cache, serialised)', 'share_of_forward'])
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, v[0], f'{v[1]:.3f}', f'{v[1] / tot:.4f}'])
        w.writerow(['TOTAL', sum(v[0] for v in agg.values()), f'{tot:.3f}', '1.0'])
    print('wrote', out)
    traffic['launch_shares'] = dict(source=f'profiles/{tag}_launch_shares.csv', shares={k: v[1] / tot for k, v in agg.items()})
json.dump(traffic, open(traffic_path, 'w'), indent=1)

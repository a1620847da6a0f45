"""Per-kernel table from `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active` over one training step (north_star: "ncu captures reporting
achieved HBM GB/s and tensor-pipe % against Blackwell peak").
usage: python scripts/ncu_kernel_table.py gpurun_out/r02_cfg4_kernels.csv [hbm_peak_GBs] > profiles/r02_cfg4_kernel_table.txt"""
import collections
import csv
import json
import os
import re
import sys

path = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
if peak is None:
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        peak = 6650.0
lines = [l for l in open(path, errors='replace') if not l.startswith('==')]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r'^void ', '', r['Kernel Name'])
    name = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', name)
    name = re.sub(r'\(.*$', '', name)[:90]
    a = agg.setdefault(name, dict(n=set(), t=0.0, rd=0.0, wr=0.0, tp=0.0, tpn=0))
    m, v, u = r['Metric Name'], float(r['Metric Value'].replace(',', '')), r['Metric Unit']
    a['n'].add(r['ID'])
    scale = {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3, 'second': 1e6}.get(u, 1.0)
    bscale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1.0)
    if m == 'gpu__time_duration.sum': a['t'] += v * scale
    elif m == 'dram__bytes_read.sum': a['rd'] += v * bscale
    elif m == 'dram__bytes_write.sum': a['wr'] += v * bscale
    elif m.startswith('sm__pipe_tensor_cycles_active'): a['tp'] += v; a['tpn'] += 1
tot = sum(a['t'] for a in agg.values())
print(f'{len(rows)} metric rows, {sum(len(a["n"]) for a in agg.values())} launches, {tot / 1e3:.3f} ms of kernel time (cold-cache, serialised: compare SHARES)')
print(f'HBM peak used for the fraction: {peak:.0f} GB/s (measured copy bandwidth)\n')
print(f'{"kernel":92s}{"n":>5s}{"ms":>9s}{"share":>7s}{"avg us":>9s}{"DRAM GB":>9s}{"GB/s":>8s}{"of HBM":>8s}{"tensor%":>9s}')
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]['t']):
    n = len(a['n'])
    gb = (a['rd'] + a['wr']) / 1e9
    gbs = gb / (a['t'] * 1e-6) if a['t'] else 0.0
    tp = a['tp'] / a['tpn'] if a['tpn'] else 0.0
    print(f'{name:92s}{n:5d}{a["t"] / 1e3:9.3f}{a["t"] / tot * 100:6.1f}%{a["t"] / n:9.1f}{gb:9.2f}{gbs:8.0f}{gbs / peak * 100:7.1f}%{tp:9.1f}')

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try:
        v = float(row['Metric Value'].replace(',', ''))
    except Exception:
        continue
    u = row['Metric Unit']
    v *= {'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6}.get(u, 1)
    name = re.sub(r'^void |<unnamed>::|\(.*', '', row['Kernel Name'])[:80]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f'total {tot / 1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f'{v[1] / 1e6:9.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  avg {v[1] / v[0] / 1e3:8.1f} us  {k}')
